"""The reference's own golden vectors for the hash-aggregate operators, restated from
vinum_cpp/test/hash_agg_test.cpp (fixture tables :155-283, expected batches :340-778).

The gtest target cannot be built here (it downloads googletest and needs
ArrowTesting), so its fixtures are restated as data.  `cases()` are the numeric-key
vectors; `generic_cases()` the string / bool key and string MIN/MAX vectors of
GenericHashAggregate.

Each case: (name, table, groupby_cols, agg_cols, funcs[(type, column, out_name)],
expected pyarrow.RecordBatch, sort_cols).  As in the gtest (:108-133) the table is fed
in two half-size batches and the result is sorted by `sort_cols` before comparing.
"""
import decimal

import pyarrow as pa

TS_MS = pa.timestamp("ms")
T32_MS = pa.time32("ms")


def _arr(values, valid, t):
    return pa.array([v if ok else None for v, ok in zip(values, valid)], type=t)


T, F = True, False


def test_table() -> pa.Table:
    """CreateTestTable, hash_agg_test.cpp:155-249 (numeric columns)."""
    return pa.table({
        "id": _arr([1, 2, 3, 4, 5, 6, 7, 8], [T] * 8, pa.int64()),
        "timestamp_int64": _arr([1602127614, 1602217613, 1602304012, 1602390411, 0, 1602563209, 0, 1602736007],
                                [T, T, T, T, F, T, F, T], pa.int64()),
        "lat": _arr([52.51, 48.51, 44.89, 42.89, 44.89, 48.51, 44.89, 52.51], [T] * 8, pa.float64()),
        "lng": _arr([13.66, 12.3, 14.23, 15.89, 14.23, 12.3, 14.23, 13.66], [T] * 8, pa.float64()),
        "total": _arr([0, 143.15, 33.4, 53.1, 0, 0, 33.4, 0], [F, T, T, T, F, F, T, F], pa.float64()),
        "grp_int8": _arr([0, 2, 7, 3, 1, 2, 1, 1], [F, T, F, T, T, T, T, T], pa.int8()),
        "grp_neg_int8": _arr([0, -1, -1, 3, 1, -1, 1, 1], [F, T, F, T, T, T, T, T], pa.int8()),
        "date64": pa.array([None, 1611664426386, 1611664426519, 1611664416382, None, 1611664426519, 1611664416382,
                            1611664426386], type=pa.int64()).cast(pa.date64()),
        "time32": pa.array([None, 7, None, 7, 41, 130, None, 130], type=pa.int32()).cast(T32_MS),
        "timestamp": pa.array([1611664420588, 1611663913570, None, 1611664414385, 1611664420588, None, None,
                               1611664414385], type=pa.int64()).cast(TS_MS),
        "grp_neg_int64": _arr([-9223372036854775807, -9223372036854775806, 9223372036854775807,
                               -9223372036854775807, 9223372036854775806, 9223372036854775806,
                               9223372036854775807, -9223372036854775806], [T] * 8, pa.int64()),
    })


def overflow_table() -> pa.Table:
    """CreateOverflowTestTable, hash_agg_test.cpp:251-273."""
    return pa.table({
        "id": _arr([1, 2, 1, 1, 2, 2, 1, 1], [T] * 8, pa.int64()),
        "int_64": _arr([9223372036854775807, 9223372036854775806, 9223372036854775805, 9223372036854775804,
                        9223372036854775803, 9223372036854775802, 9223372036854775801, 9223372036854775799],
                       [T, T, T, T, F, T, F, T], pa.int64()),
        "uint_64": _arr([18446744073709551615, 18446744073709551614, 18446744073709551613, 18446744073709551612,
                         18446744073709551611, 18446744073709551610, 18446744073709551609, 18446744073709551608],
                        [T, T, T, T, F, T, F, T], pa.uint64()),
    })


def empty_batch_table() -> pa.Table:
    """CreateEmptyTestRecordBatch, hash_agg_test.cpp:275-283."""
    return pa.table({"id": pa.array([], type=pa.int64())})


def _dec(values):
    return pa.array([decimal.Decimal(v) for v in values], type=pa.decimal128(38, 0))


def cases():
    out = []
    # CreateDoubleGrp_IntArgFuncs, :340-386
    out.append(("double_grp__int_arg_funcs", test_table(), ["lat"], ["lat"],
                [("COUNT_STAR", "", "count"), ("MIN", "id", "min_0"), ("MAX", "id", "max_0"),
                 ("SUM", "id", "sum_0"), ("AVG", "id", "avg_0")],
                pa.RecordBatch.from_arrays([
                    pa.array([42.89, 44.89, 48.51, 52.51], type=pa.float64()),
                    pa.array([1, 3, 2, 2], type=pa.uint64()),
                    pa.array([4, 3, 2, 1], type=pa.int64()),
                    pa.array([4, 7, 6, 8], type=pa.int64()),
                    pa.array([4, 15, 8, 9], type=pa.int64()),
                    pa.array([4.0, 5.0, 4.0, 4.5], type=pa.float64()),
                ], names=["lat", "count", "min_0", "max_0", "sum_0", "avg_0"]), [0]))
    # CreateInt64Grp_IntOverflowArgFuncs, :388-437
    out.append(("int64_grp__int_overflow_arg_funcs", overflow_table(), ["id"], ["id"],
                [("SUM", "int_64", "sum_1"), ("SUM", "uint_64", "sum_2"), ("AVG", "int_64", "avg_1"),
                 ("AVG", "uint_64", "avg_2")],
                pa.RecordBatch.from_arrays([
                    pa.array([1, 2], type=pa.int64()),
                    _dec(["36893488147419103215", "18446744073709551608"]),
                    _dec(["73786976294838206448", "36893488147419103224"]),
                    pa.array([9.223372036854776e+18, 9.223372036854776e+18], type=pa.float64()),
                    pa.array([1.8446744073709552e+19, 1.8446744073709552e+19], type=pa.float64()),
                ], names=["id", "sum_1", "sum_2", "avg_1", "avg_2"]), [0]))
    # CreateInt8Grp_DoubleArgFuncs, :479-531 (NULL key group sorts last)
    out.append(("int8_grp__double_arg_funcs", test_table(), ["grp_int8"], ["grp_int8"],
                [("COUNT_STAR", "", "count"), ("COUNT", "total", "count_9"), ("MIN", "lat", "min_6"),
                 ("MAX", "lat", "max_6"), ("SUM", "lat", "sum_6"), ("AVG", "lat", "avg_6")],
                pa.RecordBatch.from_arrays([
                    pa.array([1, 2, 3, None], type=pa.int8()),
                    pa.array([3, 2, 1, 2], type=pa.uint64()),
                    pa.array([1, 1, 1, 1], type=pa.uint64()),
                    pa.array([44.89, 48.51, 42.89, 44.89], type=pa.float64()),
                    pa.array([52.51, 48.51, 42.89, 52.51], type=pa.float64()),
                    pa.array([142.29, 97.02, 42.89, 97.4], type=pa.float64()),
                    pa.array([47.43, 48.51, 42.89, 48.7], type=pa.float64()),
                ], names=["grp_int8", "count", "count_9", "min_6", "max_6", "sum_6", "avg_6"]), [0]))
    # CreateMultiIntGrp_DateArgFuncs, :533-599
    out.append(("multi_int_grp__date_arg_funcs", test_table(),
                ["grp_neg_int8", "date64", "time32", "timestamp"], ["grp_neg_int8", "date64", "time32", "timestamp"],
                [("COUNT_STAR", "", "count"), ("MIN", "date64", "min_12"), ("MAX", "timestamp", "max_14"),
                 ("SUM", "time32", "sum_13")],
                pa.RecordBatch.from_arrays([
                    pa.array([-1, -1, 1, 1, 1, 3, None, None], type=pa.int8()),
                    pa.array([1611664426386, 1611664426519, 1611664416382, 1611664426386, None, 1611664416382,
                              1611664426519, None], type=pa.int64()).cast(pa.date64()),
                    pa.array([7, 130, None, 130, 41, 7, None, None], type=pa.int32()).cast(T32_MS),
                    pa.array([1611663913570, None, None, 1611664414385, 1611664420588, 1611664414385, None,
                              1611664420588], type=pa.int64()).cast(TS_MS),
                    pa.array([1] * 8, type=pa.uint64()),
                    pa.array([1611664426386, 1611664426519, 1611664416382, 1611664426386, None, 1611664416382,
                              1611664426519, None], type=pa.int64()).cast(pa.date64()),
                    pa.array([1611663913570, None, None, 1611664414385, 1611664420588, 1611664414385, None,
                              1611664420588], type=pa.int64()).cast(TS_MS),
                    pa.array([7, 130, None, 130, 41, 7, None, None], type=pa.int32()).cast(T32_MS),
                ], names=["grp_neg_int8", "date64", "time32", "timestamp", "count", "min_12", "max_14", "sum_13"]),
                [0, 1, 2, 3]))
    # CreateNegInt64Grp_TimestampArgFuncs, :652-708
    out.append(("neg_int64_grp__timestamp_arg_funcs", test_table(), ["grp_neg_int64"], ["grp_neg_int64"],
                [("COUNT_STAR", "", "count"), ("COUNT", "timestamp", "count_ts"), ("MIN", "timestamp", "min_14"),
                 ("MAX", "timestamp", "max_14"), ("AVG", "grp_int8", "avg_10"), ("AVG", "grp_neg_int8", "avg_11")],
                pa.RecordBatch.from_arrays([
                    pa.array([-9223372036854775807, -9223372036854775806, 9223372036854775806, 9223372036854775807],
                             type=pa.int64()),
                    pa.array([2, 2, 2, 2], type=pa.uint64()),
                    pa.array([2, 2, 1, 0], type=pa.uint64()),
                    pa.array([1611664414385, 1611663913570, 1611664420588, None], type=pa.int64()).cast(TS_MS),
                    pa.array([1611664420588, 1611664414385, 1611664420588, None], type=pa.int64()).cast(TS_MS),
                    pa.array([3.0, 1.5, 1.5, 1.0], type=pa.float32()),
                    pa.array([3.0, 0, 0, 1.0], type=pa.float32()),
                ], names=["grp_neg_int64", "count", "count_ts", "min_14", "max_14", "avg_10", "avg_11"]), [0]))
    # CreateNoGrp_AggFuncs, :710-759
    out.append(("no_grp__agg_funcs", test_table(), [], [],
                [("COUNT_STAR", "", "count_star"), ("COUNT", "timestamp_int64", "count_int64"),
                 ("MIN", "timestamp_int64", "min_int64"), ("MAX", "timestamp_int64", "max_int64"),
                 ("SUM", "timestamp_int64", "sum_int64"), ("AVG", "timestamp_int64", "avg_int64")],
                pa.RecordBatch.from_arrays([
                    pa.array([8], type=pa.uint64()), pa.array([6], type=pa.uint64()),
                    pa.array([1602127614], type=pa.int64()), pa.array([1602736007], type=pa.int64()),
                    pa.array([9614338866], type=pa.int64()), pa.array([1602389811.0], type=pa.float64()),
                ], names=["count_star", "count_int64", "min_int64", "max_int64", "sum_int64", "avg_int64"]), []))
    # CreateEmptyTable_AggFuncs, :761-778
    out.append(("empty_table__agg_funcs", empty_batch_table(), [], [],
                [("COUNT_STAR", "", "count_star")],
                pa.RecordBatch.from_arrays([pa.array([0], type=pa.uint64())], names=["count_star"]), []))
    return out


def generic_table() -> pa.Table:
    """The string / bool columns of CreateTestTable (hash_agg_test.cpp:163-183) next to the numeric ones."""
    t = test_table()
    extra = {
        "date": _arr(["", "2020-10-09T04:26:53", "2020-10-10T04:26:52", "2020-10-11T04:26:51", "2020-10-12T04:26:50",
                      "2020-10-13T04:26:49", "0", "2020-10-15T04:26:47"], [F, T, T, T, T, T, F, T], pa.string()),
        "is_vendor": _arr([True, True, False, False, True, False, False, False], [T, T, T, F, T, F, F, F], pa.bool_()),
        "city_from": _arr(["", "Munich", "", "San Francisco", "Berlin", "Munich", "Berlin", "Berlin"],
                          [F, T, F, T, T, T, T, T], pa.string()),
        "city_to": _arr(["Munich", "Riva", "Naples", "Naples", "Riva", "Riva", "Munich", "Munich"], [T] * 8, pa.string()),
        "name": _arr(["Joe", "", "Joseph", "Joseph", "", "Jonas", "Joseph", "Joe"], [T, F, T, T, F, T, T, T], pa.string()),
    }
    for k, v in extra.items():
        t = t.append_column(k, v)
    return t


def generic_cases():
    out = []
    # CreateStringGrp_DoubleArgFuncs, :286-338 (NULL key group sorts last)
    out.append(("string_grp__double_arg_funcs", generic_table(), ["city_from"], ["city_from"],
                [("COUNT_STAR", "", "count"), ("COUNT", "total", "count_9"), ("MIN", "lat", "min_6"),
                 ("MAX", "lat", "max_6"), ("SUM", "lat", "sum_6"), ("AVG", "lat", "avg_6")],
                pa.RecordBatch.from_arrays([
                    _arr(["Berlin", "Munich", "San Francisco", ""], [T, T, T, F], pa.string()),
                    pa.array([3, 2, 1, 2], type=pa.uint64()),
                    pa.array([1, 1, 1, 1], type=pa.uint64()),
                    pa.array([44.89, 48.51, 42.89, 44.89], type=pa.float64()),
                    pa.array([52.51, 48.51, 42.89, 52.51], type=pa.float64()),
                    pa.array([142.29, 97.02, 42.89, 97.4], type=pa.float64()),
                    pa.array([47.43, 48.51, 42.89, 48.7], type=pa.float64()),
                ], names=["city_from", "count", "count_9", "min_6", "max_6", "sum_6", "avg_6"]), [0]))
    # CreateInt64Grp_StringArgFuncs, :439-477
    dates = _arr(["", "2020-10-09T04:26:53", "2020-10-10T04:26:52", "2020-10-11T04:26:51", "2020-10-12T04:26:50",
                  "2020-10-13T04:26:49", "", "2020-10-15T04:26:47"], [F, T, T, T, T, T, F, T], pa.string())
    out.append(("int64_grp__string_arg_funcs", generic_table(), ["id"], ["id"],
                [("COUNT", "date", "count_2"), ("MIN", "date", "min_2"), ("MAX", "date", "max_2")],
                pa.RecordBatch.from_arrays([
                    pa.array([1, 2, 3, 4, 5, 6, 7, 8], type=pa.int64()),
                    pa.array([0, 1, 1, 1, 1, 1, 0, 1], type=pa.uint64()),
                    dates, dates,
                ], names=["id", "count_2", "min_2", "max_2"]), [0]))
    # CreateBooleanGrp_DateArgFuncs, :601-650
    t32 = lambda vals, valid: pa.array([v if ok else None for v, ok in zip(vals, valid)], type=pa.int32()).cast(T32_MS)  # noqa: E731
    out.append(("boolean_grp__date_arg_funcs", generic_table(), ["is_vendor"], ["is_vendor"],
                [("COUNT_STAR", "", "count"), ("MIN", "time32", "min_12"), ("MAX", "time32", "max_14"),
                 ("SUM", "time32", "sum_13"), ("AVG", "time32", "avg_13")],
                pa.RecordBatch.from_arrays([
                    _arr([False, True, False], [T, T, F], pa.bool_()),
                    pa.array([1, 3, 4], type=pa.uint64()),
                    t32([0, 7, 7], [F, T, T]), t32([0, 41, 130], [F, T, T]), t32([0, 48, 267], [F, T, T]),
                    _arr([0.0, 24.0, 89.0], [F, T, T], pa.float64()),
                ], names=["is_vendor", "count", "min_12", "max_14", "sum_13", "avg_13"]), [0]))
    return out


def split_in_two(table: pa.Table):
    """aggregate_and_sort feeds the table in chunks of num_rows >> 1 (:111-121)."""
    mid = table.num_rows >> 1
    if mid > 0:
        return table.to_batches(max_chunksize=mid)
    batches = table.to_batches()
    if not batches:
        return [pa.RecordBatch.from_arrays([pa.array([], type=f.type) for f in table.schema], schema=table.schema)]
    return batches


def sort_result(batch: pa.RecordBatch, sort_cols):
    """sort_table, hash_agg_test.cpp:75-106 (ascending, NULLs last)."""
    if batch.num_rows == 0 or not sort_cols:
        return batch
    names = batch.schema.names
    idx = pa.compute.sort_indices(batch, sort_keys=[(names[i], "ascending") for i in sort_cols])
    return batch.take(idx)
