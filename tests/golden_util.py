"""Helpers shared by the oracle (CPU) and GPU parity tests."""
import json
from pathlib import Path

import numpy as np
import pyarrow as pa
import pyarrow.compute  # noqa: F401

GOLDEN = Path(__file__).resolve().parent / "golden"


def read_arrow(name: str) -> pa.Table:
    with pa.OSFile(str(GOLDEN / name), "rb") as f:
        return pa.ipc.open_file(f).read_all()


def manifest():
    with open(GOLDEN / "cases.json") as f:
        return json.load(f)


def canonical(batch, key_cols):
    """Group-by output order is unspecified: sort rows by the key columns
    (NULLs last, NaN before NULL, -0.0 distinguished from 0.0 through its bits).

    The emitted key columns may be a strict subset of the group-by columns
    (planner.py:452-469), so they need not be unique: ties are broken by every exact
    (non-floating) column, then by the NULL-ness and finally the value of the floating
    aggregate columns (two rows whose floats are that close compare equal within the
    tolerance whichever way they are ordered)."""
    if isinstance(batch, pa.RecordBatch):
        batch = pa.Table.from_batches([batch])
    if batch.num_rows == 0 or not key_cols:
        return batch
    exact = [f.name for f in batch.schema if f.name not in key_cols and not pa.types.is_floating(f.type)
             and not pa.types.is_decimal(f.type)]
    floats = [f.name for f in batch.schema if f.name not in key_cols and pa.types.is_floating(f.type)]
    sort_keys = []  # highest priority first

    def add(name, by_bits):
        col = batch.column(name).combine_chunks()
        valid = np.asarray(col.is_valid().to_numpy(zero_copy_only=False), dtype=bool)
        t = col.type
        if pa.types.is_floating(t):
            v = np.asarray(col.fill_null(0).to_numpy(zero_copy_only=False), dtype=np.float64)
            if by_bits:
                bits = v.view(np.uint64)
                # total order on the bit patterns (distinct groups for -0.0 / 0.0 / NaN payloads)
                code = np.where(bits >> np.uint64(63), ~bits, bits | np.uint64(1 << 63))
            else:
                code = np.where(np.isnan(v), np.inf, v)
        elif pa.types.is_string(t) or pa.types.is_large_string(t):
            uniq = pa.compute.unique(col).drop_null()
            ranked = uniq.take(pa.compute.sort_indices(uniq))
            code = np.asarray(pa.compute.index_in(col, value_set=ranked).fill_null(0).to_numpy(zero_copy_only=False)).astype(np.uint64)
        elif pa.types.is_boolean(t):
            code = np.asarray(col.cast(pa.uint8()).fill_null(0).to_numpy(zero_copy_only=False)).astype(np.uint64)
        else:
            phys = pa.int32() if (pa.types.is_date32(t) or pa.types.is_time32(t)) else (
                pa.int64() if pa.types.is_temporal(t) else t)
            v = np.asarray(col.view(phys).fill_null(0).to_numpy(zero_copy_only=False))
            code = v.astype(np.int64).view(np.uint64) ^ np.uint64(1 << 63) if v.dtype.kind == "i" else v.astype(np.uint64)
        sort_keys.append(~valid)
        sort_keys.append(code)

    for name in list(key_cols) + exact:
        add(name, True)
    for name in floats:
        add(name, False)
    order = np.lexsort(tuple(reversed(sort_keys)))
    return batch.take(pa.array(order, type=pa.int64()))


def assert_tables_match(got, want, key_cols=(), rtol=1e-6, float_exact_cols=()):
    """Schema (names + types) equal; integer / temporal / decimal / key columns
    bit-exact; floating aggregate columns within `rtol` relative (north_star: 1e-6)."""
    got = canonical(got, list(key_cols))
    want = canonical(want, list(key_cols))
    assert got.schema.names == want.schema.names, (got.schema.names, want.schema.names)
    for name in want.schema.names:
        assert got.schema.field(name).type == want.schema.field(name).type, (
            name, got.schema.field(name).type, want.schema.field(name).type)
    assert got.num_rows == want.num_rows, (got.num_rows, want.num_rows)
    for name in want.schema.names:
        g = got.column(name).combine_chunks()
        w = want.column(name).combine_chunks()
        gv = np.asarray(g.is_valid().to_numpy(zero_copy_only=False), dtype=bool)
        wv = np.asarray(w.is_valid().to_numpy(zero_copy_only=False), dtype=bool)
        assert np.array_equal(gv, wv), f"validity differs in column {name}"
        if pa.types.is_floating(w.type) and name not in key_cols and name not in float_exact_cols:
            ga = np.asarray(g.fill_null(0).to_numpy(zero_copy_only=False), dtype=np.float64)[wv]
            wa = np.asarray(w.fill_null(0).to_numpy(zero_copy_only=False), dtype=np.float64)[wv]
            assert np.allclose(ga, wa, rtol=rtol, atol=0, equal_nan=True), f"column {name}: {ga} vs {wa}"
        elif pa.types.is_floating(w.type):
            ga = np.asarray(g.fill_null(0).to_numpy(zero_copy_only=False))[wv]
            wa = np.asarray(w.fill_null(0).to_numpy(zero_copy_only=False))[wv]
            assert np.array_equal(ga.view(np.uint8), wa.view(np.uint8)), f"column {name} differs bitwise"
        else:
            assert g.filter(pa.array(wv)).equals(w.filter(pa.array(wv))), f"column {name} differs: {g} vs {w}"
