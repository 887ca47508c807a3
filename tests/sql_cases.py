"""SQL parity cases shared by the golden generator (oracle/gen_sql_golden.py, which runs
them through the REFERENCE's planner + executor) and the GPU tests (tests/test_sql_gpu.py).

Queries follow the categories of the reference's own suite (vinum/tests/test_query_results.py:
select / expressions :50-147, WHERE :131-306, GROUP BY + aggregates :315-613, ORDER BY :627-746,
LIMIT :324-346, NULL semantics :1185-1301) on a deterministic synthetic table."""
import numpy as np
import pyarrow as pa


def main_table(n: int = 4000) -> pa.Table:
    rng = np.random.default_rng(20240917)
    ids = np.arange(n, dtype=np.int64)
    k_i32 = (rng.integers(0, 17, n)).astype(np.int32)
    k_i64 = rng.integers(0, 300, n).astype(np.int64)
    grp_choices = np.array(["air", "bus", "car", "ship", "train", "walk"], dtype=object)
    grp = grp_choices[rng.integers(0, 6, n)]
    grp_mask = rng.random(n) < 0.03
    v_f64 = np.round(rng.normal(0, 100, n), 3)
    v_i64 = rng.integers(-1_000_000, 1_000_000, n).astype(np.int64)
    v_f32 = rng.random(n).astype(np.float32)
    small = rng.integers(1, 9, n).astype(np.int8)
    flag = rng.random(n) < 0.4
    nf = np.round(rng.normal(10, 5, n), 2)
    nf_mask = rng.random(n) < 0.1
    ni = rng.integers(0, 50, n).astype(np.int64)
    ni_mask = rng.random(n) < 0.1
    names = np.array(["alpha", "beta", "gamma", "delta", "epsilon", "zeta", "eta", "theta", "iota", "kappa"], dtype=object)
    name = names[rng.integers(0, 10, n)]
    ts = (np.int64(1_600_000_000) + rng.integers(0, 10_000_000, n)).astype("datetime64[s]")
    return pa.table({
        "id": pa.array(ids),
        "k_i32": pa.array(k_i32),
        "k_i64": pa.array(k_i64),
        "grp": pa.array(grp, type=pa.string(), mask=grp_mask),
        "v_f64": pa.array(v_f64),
        "v_i64": pa.array(v_i64),
        "v_f32": pa.array(v_f32),
        "small": pa.array(small),
        "flag": pa.array(flag),
        "nf": pa.array(nf, mask=nf_mask),
        "ni": pa.array(ni, mask=ni_mask),
        "ts": pa.array(ts),
        "name": pa.array(name, type=pa.string()),
    })


def readme_table() -> pa.Table:
    """BASELINE.json configs[0] (README.rst:90-96)."""
    return pa.table({"value": [300.1, 2.8, 880.0], "mode": ["air", "bus", "air"]})


TABLES = {"main": main_table, "readme": readme_table}

# (table, sql, ordered) -- ordered: row order is defined by the query and must match exactly
CASES = [
    ("readme", "SELECT value FROM t WHERE mode='air'", True),
    ("readme", "SELECT * FROM t", True),
    ("readme", "SELECT mode, sum(value), count(*) FROM t GROUP BY mode", False),
    # ---- projection / expressions
    ("main", "select * from t", True),
    ("main", "select id, v_f64 as val, grp from t", True),
    ("main", "select id + 1, v_i64 - id, v_f64 * 2, v_i64 / 7, v_i64 % 7, -v_i64 from t", True),
    ("main", "select id, v_i64 + v_f64 as s, v_f32 * small as p, small / 2 as h from t", True),
    ("main", "select k_i64 & 12, k_i64 | 3, k_i64 # 5, ~k_i64 from t", True),
    ("main", "select (v_i64 + 3) * (id - 2) / (small + 1) as e, 7 as seven, id from t", True),
    ("main", "select v_f64 > 0 as pos, k_i32 = 3 as is3, id from t", True),
    ("main", "select nf + 1 as a, ni * 2 as b, nf / ni as c from t", True),
    ("main", "select abs(v_f64) as a, sqrt(v_f32) as r, power(small, 2) as sq, log(small) as lg, cos(v_f32) as c from t", True),
    ("main", "select np.sin(v_f64) as s, np.floor(v_f64) as f, to_int(v_f64) as i, to_float(k_i32) as fl from t", True),
    ("main", "select upper(grp) as u, lower(grp) as l from t where grp is not null", True),
    # ---- WHERE
    ("main", "select id, v_f64 from t where v_f64 > 50.5", True),
    ("main", "select * from t where v_f64 > 0", True),
    ("main", "select id from t where k_i32 = 3", True),
    ("main", "select id from t where k_i32 != 3 and v_i64 >= 0", True),
    ("main", "select id from t where 100 > v_f64 and 5 <= k_i32", True),
    ("main", "select id from t where v_f64 < -50 or v_f64 > 50 or k_i32 = 0", True),
    ("main", "select id from t where not (v_f64 < 0) and (k_i32 < 5 or k_i32 > 12)", True),
    ("main", "select id from t where v_i64 between -1000 and 250000", True),
    ("main", "select id from t where v_f64 not between -10.5 and 80", True),
    ("main", "select id from t where k_i32 in (1, 5, 9)", True),
    ("main", "select id from t where k_i64 not in (1, 2, 3, 4, 5, 6, 7, 8, 9, 10)", True),
    ("main", "select id from t where v_i64 + id > v_f64 * 1000", True),
    ("main", "select id from t where v_i64 > k_i64", True),
    ("main", "select id, grp from t where grp = 'air'", True),
    ("main", "select id, grp from t where grp != 'bus' and v_f64 > 0", True),
    ("main", "select id, grp from t where grp in ('car', 'walk')", True),
    ("main", "select id, name from t where name like '%et%'", True),
    ("main", "select id, name from t where name not like '_eta' and v_f64 > 0", True),
    ("main", "select id, nf from t where nf is null", True),
    ("main", "select id, ni from t where ni is not null and ni > 25", True),
    ("main", "select id, nf from t where nf > 12", True),
    ("main", "select id, grp from t where grp is null", True),
    ("main", "select id, name || '_' || grp as tag, concat(name, '-', id) as c2 from t where grp is not null", True),
    ("main", "select id from t where v_f64 > 1000000", True),
    # ---- GROUP BY + aggregates
    ("main", "select k_i32, count(*) from t group by k_i32", False),
    ("main", "select k_i32, count(*), sum(v_f64), min(v_f64), max(v_f64), avg(v_f64) from t group by k_i32", False),
    ("main", "select k_i64, sum(v_i64), avg(v_i64), min(v_i64), max(v_i64), count(v_i64) from t group by k_i64", False),
    ("main", "select k_i32, sum(small), avg(small), sum(v_f32), avg(v_f32) from t group by k_i32", False),
    ("main", "select k_i64, count(*) as cnt, sum(v_f64) as total from t where v_f64 > 0 group by k_i64", False),
    ("main", "select k_i32, k_i64, count(*), sum(v_f64) from t group by k_i32, k_i64", False),
    ("main", "select grp, count(*), sum(v_f64), avg(v_i64) from t group by grp", False),
    ("main", "select grp, k_i32, count(*) from t where v_f64 < 10 group by grp, k_i32", False),
    ("main", "select count(*) from t", True),
    ("main", "select count(*), sum(v_f64), min(v_i64), max(v_i64), avg(v_f32) from t where k_i32 > 8", True),
    ("main", "select count(nf), count(ni), sum(nf), sum(ni), avg(nf), avg(ni), min(nf), max(ni) from t", True),
    ("main", "select ni, count(*), sum(nf) from t group by ni", False),
    ("main", "select k_i32, sum(v_i64 + id) as s, avg(v_f64 * 2) as a from t group by k_i32", False),
    ("main", "select k_i64 % 10 as bucket, count(*), sum(v_f64) from t group by bucket", False),
    ("main", "select k_i32, sum(v_f64) / count(*) as mean, max(v_f64) - min(v_f64) as spread from t group by k_i32", False),
    ("main", "select k_i32, count(*) from t group by k_i32 having count(*) > 230", False),
    ("main", "select k_i64, sum(v_f64) as total from t group by k_i64 having total > 0 and k_i64 < 150", False),
    ("main", "select flag, count(*), sum(v_i64) from t group by flag", False),
    ("main", "select ts, count(*) from t where id < 50 group by ts", False),
    ("main", "select k_i32, min(ts), max(ts) from t group by k_i32", False),
    ("main", "select distinct k_i32 from t", False),
    ("main", "select distinct grp, flag from t", False),
    ("main", "select np.min(v_f64), np.max(v_f64), np.sum(v_i64) from t", True),
    # ---- ORDER BY / LIMIT
    ("main", "select id, v_f64 from t order by v_f64", True),
    ("main", "select id, v_f64 from t order by v_f64 desc", True),
    ("main", "select id, k_i32, v_i64 from t order by k_i32, v_i64 desc", True),
    ("main", "select id, k_i32 from t order by k_i32", True),
    ("main", "select id, k_i32, small from t order by k_i32 desc, small asc, id desc", True),
    ("main", "select id, nf from t order by nf, id", True),
    ("main", "select id, ni from t order by ni desc, id", True),
    ("main", "select id, v_f64 from t where k_i32 = 2 order by v_f64 desc limit 10", True),
    ("main", "select id, v_i64 from t order by v_i64 limit 25 offset 100", True),
    ("main", "select id from t limit 7", True),
    ("main", "select id, v_i64 * 2 as dbl from t order by dbl desc limit 5", True),
    ("main", "select id, v_f64 from t order by abs(v_f64) limit 12", True),
    ("main", "select id, grp from t order by grp, id limit 40", True),
    ("main", "select id, ts from t order by ts desc, id limit 30", True),
    ("main", "select k_i32, count(*) as cnt from t group by k_i32 order by k_i32", True),
    ("main", "select k_i64, sum(v_f64) as total from t group by k_i64 order by total desc limit 10", True),
    ("main", "select grp, count(*) as cnt, avg(v_f64) as m from t where grp is not null group by grp order by grp", True),
    ("main", "select k_i32, k_i64, count(*) as c from t group by k_i32, k_i64 order by k_i32 desc, k_i64 limit 50", True),
    ("main", "select k_i64, count(*) as c, sum(v_f64) as s from t where v_f64 > 0.5 group by k_i64 order by k_i64", True),
]
