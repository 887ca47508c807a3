export VK_SANITIZE_SEL='test_northstar_filter_aggregate_vs_oracle and (partitioned or general) or test_group_by_hostile_key_distributions and (partitioned or general) and (late_groups or sentinel or one_hot_90) or test_group_by_more_groups_than_the_first_table_holds and uniform_2e6 and 1-partitioned'
export VK_SANITIZE_TOOLS='memcheck racecheck synccheck'
bash scripts/sanitize.sh
grep -c "Race reported\|Error:" gpurun_out/sanitize_racecheck.log
grep -o "Race reported between.*" gpurun_out/sanitize_racecheck.log | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -20
grep -B1 -A6 "========= Error\|========= Warning" gpurun_out/sanitize_racecheck.log | grep -o "in .*kernel[^(]*\|vk_[a-z_]*\.cuh\?:[0-9]*" | sort | uniq -c | sort -rn | head -30
