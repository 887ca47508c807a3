"""Generates tests/golden/*.arrow from the reference's OWN operators (oracle/_ref, the
unmodified C++ under /root/reference compiled by oracle/build_ref.sh) and, for the
Python-side operators (comparison / filter / arithmetic), from the reference's own
callables imported from /root/reference/vinum/core/expressions.py.

Run in the build container only (needs /root/reference):
    python oracle/gen_golden.py
The fixtures are small Arrow IPC files: `<case>.in.arrow` (input table),
`<case>.out.arrow` (reference result), plus cases.json describing each case.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np
import pyarrow as pa

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import ref  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"
REFERENCE = Path("/root/reference")


def _write(name: str, table: pa.Table) -> None:
    with pa.OSFile(str(GOLDEN / name), "wb") as f:
        with pa.ipc.new_file(f, table.schema) as w:
            w.write_table(table)


def random_table(rng: np.random.Generator, n: int, null_frac: float) -> pa.Table:
    def nulls(a, t):
        mask = rng.random(n) < null_frac
        return pa.array(a, type=t, mask=mask)
    f64 = rng.normal(0, 100, n)
    f64[rng.random(n) < 0.02] = np.nan
    f64[rng.random(n) < 0.02] = -0.0
    f64[rng.random(n) < 0.01] = np.inf
    cols = {
        "k_i8": nulls(rng.integers(-3, 4, n).astype(np.int8), pa.int8()),
        "k_i32": nulls(rng.integers(-50, 50, n).astype(np.int32), pa.int32()),
        "k_i64": nulls(rng.choice(np.array([-(2**63) + 1, -7, 0, 5, 2**62, 2**63 - 1], dtype=np.int64), n), pa.int64()),
        "k_u16": nulls(rng.integers(0, 9, n).astype(np.uint16), pa.uint16()),
        "k_f64": nulls(rng.choice(np.array([0.0, -0.0, 1.5, np.nan, -2.25, np.inf]), n), pa.float64()),
        "k_f32": nulls(rng.choice(np.array([0.5, -1.0, 3.25], dtype=np.float32), n), pa.float32()),
        "k_ts": nulls(rng.integers(1_600_000_000_000, 1_600_000_000_005, n), pa.timestamp("ms")),
        "k_d32": nulls(rng.integers(18000, 18004, n).astype(np.int32), pa.date32()),
        "v_i8": nulls(rng.integers(-128, 128, n).astype(np.int8), pa.int8()),
        "v_i16": nulls(rng.integers(-30000, 30000, n).astype(np.int16), pa.int16()),
        "v_i32": nulls(rng.integers(-2**31, 2**31, n).astype(np.int32), pa.int32()),
        "v_i64": nulls(rng.integers(-2**40, 2**40, n), pa.int64()),
        "v_i64_big": nulls(rng.integers(-2**62, 2**62, n), pa.int64()),
        "v_u64_big": nulls(rng.integers(2**62, 2**64 - 1, n, dtype=np.uint64), pa.uint64()),
        "v_u8": nulls(rng.integers(0, 256, n).astype(np.uint8), pa.uint8()),
        "v_u32": nulls(rng.integers(0, 2**32, n).astype(np.uint32), pa.uint32()),
        "v_u64": nulls(rng.integers(0, 2**40, n).astype(np.uint64), pa.uint64()),
        "v_f32": nulls(rng.normal(0, 10, n).astype(np.float32), pa.float32()),
        "v_f64": nulls(f64, pa.float64()),
        "v_f64_clean": nulls(rng.normal(0, 100, n), pa.float64()),
        "v_t32": nulls(rng.integers(0, 86_400_000, n).astype(np.int32), pa.time32("ms")),
        "v_ts": nulls(rng.integers(1_600_000_000_000, 1_700_000_000_000, n), pa.timestamp("ms")),
        "row": pa.array(np.arange(n, dtype=np.int64)),
    }
    return pa.table(cols)


def agg_cases():
    F = lambda t, c, o=None: (t, c, o or f"{t.lower()}_{c}")  # noqa: E731
    allf = lambda c: [F("COUNT", c), F("MIN", c), F("MAX", c), F("SUM", c), F("AVG", c)]  # noqa: E731
    cs = ("COUNT_STAR", "", "count_star")
    return [
        ("agg_i32_key_floats", ["k_i32"], ["k_i32"], [cs] + allf("v_f64_clean") + allf("v_f32")),
        ("agg_i8_key_ints", ["k_i8"], ["k_i8"], [cs] + allf("v_i8") + allf("v_i16") + allf("v_i32")),
        ("agg_i64_extreme_keys", ["k_i64"], ["k_i64"], [cs] + allf("v_i64") + allf("v_u64")),
        ("agg_f64_key", ["k_f64"], ["k_f64"], [cs] + allf("v_u8") + allf("v_u32")),
        ("agg_f32_key", ["k_f32"], ["k_f32"], [cs, F("SUM", "v_f64_clean"), F("COUNT", "v_ts"), F("MIN", "v_ts"), F("MAX", "v_ts")]),
        ("agg_ts_key_time", ["k_ts"], ["k_ts"], [cs, F("SUM", "v_t32"), F("AVG", "v_t32"), F("MIN", "v_t32"), F("MAX", "v_t32")]),
        ("agg_multi_2", ["k_i8", "k_u16"], ["k_i8", "k_u16"], [cs] + allf("v_f64_clean") + [F("SUM", "v_i64")]),
        ("agg_multi_4", ["k_i8", "k_d32", "k_f32", "k_ts"], ["k_i8", "k_ts"], [cs, F("AVG", "v_i64"), F("MAX", "v_u64")]),
        ("agg_distinct_only", ["k_i32", "k_i8"], ["k_i32", "k_i8"], []),
        ("agg_nogroup", [], [], [cs] + allf("v_f64_clean") + allf("v_i64") + allf("v_u8")),
        # 128-bit sums that overflow int64/uint64 -> decimal128(38,0).  Dense table only: with a
        # validity bitmap the reference's CopyBuilder (agg_funcs.h:425-434) takes the NULL-ness of
        # already summarised groups from unrelated INPUT rows (`array_iter->IsNull(i)`), which is
        # data-dependent garbage and deliberately not reproduced (deviations).
        ("agg_overflow_big", ["k_i8"], ["k_i8"], [cs] + allf("v_i64_big") + allf("v_u64_big")),
    ]


def sort_cases():
    return [
        ("sort_f64_desc", ["v_f64"], ["DESC"]),
        ("sort_f64_asc", ["v_f64"], ["ASC"]),
        ("sort_i64_desc", ["v_i64"], ["DESC"]),
        ("sort_i8_asc_ties", ["k_i8"], ["ASC"]),
        ("sort_multi_mixed", ["k_i8", "v_f32", "k_u16"], ["DESC", "ASC", "DESC"]),
        ("sort_f64key_then_ts", ["k_f64", "v_ts"], ["ASC", "DESC"]),
        ("sort_u64_asc", ["v_u64"], ["ASC"]),
        ("sort_t32_desc", ["v_t32"], ["DESC"]),
    ]


def expression_cases(table: pa.Table):
    """Masks / arithmetic produced by the reference's own EXPRESSION_FUNCTIONS
    (vinum/core/expressions.py) on the reference's NumPy views (record_batch.py)."""
    sys.path.insert(0, str(ROOT / "oracle" / "stubs"))
    sys.path.insert(0, str(REFERENCE))
    sys.modules.setdefault("vinum_lib", ref.ref_lib())
    from vinum.core.expressions import EXPRESSION_FUNCTIONS  # type: ignore
    from vinum.parser.query import SQLExpression  # type: ignore
    from vinum.arrow.record_batch import RecordBatch  # type: ignore

    rb = RecordBatch(table.combine_chunks().to_batches()[0])

    class _C:
        def __init__(self, n):
            self.n = n

        def get_column_name(self):
            return self.n

    def col(name):
        return rb.get_np_column(_C(name))

    E = SQLExpression
    fn = lambda e: EXPRESSION_FUNCTIONS[e][0]  # noqa: E731
    out = {}
    cmp_ops = {"==": E.EQUALS, "!=": E.NOT_EQUALS, ">": E.GREATER_THAN, ">=": E.GREATER_THAN_OR_EQUAL,
               "<": E.LESS_THAN, "<=": E.LESS_THAN_OR_EQUAL}
    specs = []
    for cname, scalar in [("v_f64", 10.5), ("v_i64", 12345), ("v_i32", -7), ("v_f32", 0.1), ("v_u8", 100),
                          ("v_i16", 2.5), ("k_i8", 0), ("v_u64", 2**62)]:
        for op, e in cmp_ops.items():
            key = f"cmp|{cname}|{op}|{scalar!r}"
            with np.errstate(all="ignore"):
                out[key] = np.asarray(fn(e)(col(cname), scalar))
            specs.append({"kind": "cmp", "col": cname, "op": op, "scalar": scalar, "key": key})
    for a, b in [("v_i32", "v_i64"), ("v_f32", "v_f64"), ("v_i8", "v_i16"), ("v_f64", "v_i64")]:
        for op, e in cmp_ops.items():
            key = f"cmpcol|{a}|{op}|{b}"
            with np.errstate(all="ignore"):
                out[key] = np.asarray(fn(e)(col(a), col(b)))
            specs.append({"kind": "cmpcol", "a": a, "op": op, "b": b, "key": key})
    for cname, lo, hi in [("v_i32", -100000, 100000), ("v_f64", -50.0, 50.5), ("v_u8", 10, 20)]:
        for neg, e in ((False, E.BETWEEN), (True, E.NOT_BETWEEN)):
            key = f"between|{cname}|{lo}|{hi}|{neg}"
            with np.errstate(all="ignore"):
                out[key] = np.asarray(fn(e)(col(cname), lo, hi))
            specs.append({"kind": "between", "col": cname, "lo": lo, "hi": hi, "negate": neg, "key": key})
    for cname, vals in [("k_i8", [1, -2, 3]), ("k_u16", [0, 5]), ("k_f32", [0.5, 3.25])]:
        for neg, e in ((False, E.IN), (True, E.NOT_IN)):
            key = f"isin|{cname}|{vals}|{neg}"
            out[key] = np.asarray(fn(e)(col(cname), vals))
            specs.append({"kind": "isin", "col": cname, "values": vals, "negate": neg, "key": key})
    ar = {"+": E.ADDITION, "-": E.SUBTRACTION, "*": E.MULTIPLICATION, "/": E.DIVISION, "%": E.MODULUS}
    for a, b in [("v_i64", 3), ("v_i32", "v_i8"), ("v_f64_clean", 2.5), ("v_f32", 0.1), ("v_i16", -7), ("v_u8", "v_u32"),
                 ("v_i32", "v_f64_clean"), ("v_i64", "v_i64"), ("v_f64", "v_i8"), ("v_u64", 5), (7, "v_i32"), (2.0, "v_f32")]:
        for op, e in ar.items():
            x = col(a) if isinstance(a, str) else a
            y = col(b) if isinstance(b, str) else b
            key = f"arith|{a!r}|{op}|{b!r}"
            with np.errstate(all="ignore"):
                out[key] = np.asarray(fn(e)(x, y))
            specs.append({"kind": "arith", "a": a, "op": op, "b": b, "key": key,
                          "a_is_col": isinstance(a, str), "b_is_col": isinstance(b, str)})
    for a, b in [("v_i64", 255), ("v_i32", "v_i8"), ("v_u8", "v_u32")]:
        for op, e in {"&": E.BINARY_AND, "|": E.BINARY_OR, "#": E.BINARY_XOR}.items():
            x, y = col(a), (col(b) if isinstance(b, str) else b)
            key = f"arith|{a!r}|{op}|{b!r}"
            try:
                out[key] = np.asarray(fn(e)(x, y))
            except TypeError:
                continue  # NULLs turned the view into floats: the reference raises here too
            specs.append({"kind": "arith", "a": a, "op": op, "b": b, "key": key, "a_is_col": True,
                          "b_is_col": isinstance(b, str)})
    for a in ["v_i64", "v_f64_clean", "v_i8"]:
        key = f"arith|{a!r}|neg|None"
        out[key] = np.asarray(fn(E.NEGATION)(col(a)))
        specs.append({"kind": "arith", "a": a, "op": "neg", "b": None, "key": key, "a_is_col": True, "b_is_col": False})
    for a in ["v_i64", "v_u8"]:
        key = f"arith|{a!r}|~|None"
        try:
            out[key] = np.asarray(fn(E.BINARY_NOT)(col(a)))
        except TypeError:
            continue
        specs.append({"kind": "arith", "a": a, "op": "~", "b": None, "key": key, "a_is_col": True, "b_is_col": False})
    return specs, out


def main() -> None:
    assert ref.ref_lib() is not None, "build oracle/_ref first (oracle/build_ref.sh)"
    GOLDEN.mkdir(parents=True, exist_ok=True)
    rng = np.random.default_rng(20261017)
    manifest = {"agg": [], "sort": [], "expr": []}

    for tname, n, nf in (("t_nulls", 257, 0.15), ("t_dense", 190, 0.0)):
        table = random_table(rng, n, nf)
        _write(f"{tname}.in.arrow", table)
        batches = table.to_batches(max_chunksize=64)  # streaming across batches is pinned too
        for name, gb, ac, funcs in agg_cases():
            if name == "agg_overflow_big" and nf > 0:
                continue
            res = ref.ref_aggregate(batches, gb, ac, funcs)
            _write(f"{tname}.{name}.out.arrow", pa.Table.from_batches([res]))
            manifest["agg"].append({"table": tname, "name": name, "groupby": gb, "agg_cols": ac,
                                    "funcs": [list(f) for f in funcs]})
        for name, cols, orders in sort_cases():
            res = ref.ref_sort(batches, cols, orders)
            # the permutation is the golden value (row ids are unique)
            _write(f"{tname}.{name}.out.arrow", pa.table({"row": res.column(res.schema.get_field_index("row"))}))
            manifest["sort"].append({"table": tname, "name": name, "cols": cols, "orders": orders})
        specs, outs = expression_cases(table)
        arrays, names = [], []
        for s in specs:
            v = outs[s["key"]]
            arrays.append(pa.array(v))
            names.append(s["key"])
            s["table"] = tname
            s["dtype"] = str(v.dtype)
        _write(f"{tname}.expr.out.arrow", pa.table(arrays, names=names))
        manifest["expr"].extend(specs)

    with open(GOLDEN / "cases.json", "w") as f:
        json.dump(manifest, f, indent=1, default=lambda o: o if not isinstance(o, np.generic) else o.item())
    print(f"wrote {len(manifest['agg'])} aggregate, {len(manifest['sort'])} sort, {len(manifest['expr'])} expression cases")


if __name__ == "__main__":
    main()
