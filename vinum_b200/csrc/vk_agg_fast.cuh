// agg_fast_kernel -- fused filter -> hash aggregate for low-cardinality single-key
// group-bys (the north-star pipeline).  See DESIGN.md 3.1.
//
// One persistent CTA per SM.  Rows stream through registers (16-byte loads, next tile
// prefetched while the current one is processed); the WHERE predicate is evaluated in
// registers; the key is resolved in a CTA-shared open-addressing table to a DENSE group
// id; COUNT and up to three 64-bit accumulator cells per group live in WARP-PRIVATE
// shared-memory arrays that are updated with plain read-modify-write (shared-memory
// 64-bit atomics are CAS loops on sm_100a and far too slow for a per-row path).  Lanes
// of a warp that hit the same group in the same step are serialised in
// __match_any_sync rank order.  The whole per-row path is warp-convergent: the probe
// loop runs until __any_sync says no lane is searching, so the warp never splits into
// sub-warps that would replay the loads and the table code.
#pragma once
#include "vk_hashagg.cuh"

namespace vk {

constexpr int FA_ROWS = 4;        // rows per thread per tile (two lane-contiguous pairs)
constexpr int FA_MAX_COLS = 3;    // distinct value columns
constexpr int FA_MAX_CELLS = 3;   // 64-bit accumulator cells per group
constexpr int FA_MAX_WARPS = 16;
constexpr uint32_t GID_PENDING = 0xFFFFu;  // key claimed, dense id not published yet
constexpr uint32_t GID_SPILL = 0xFFFEu;    // more groups than the CTA holds: rows go to the global table
constexpr uint64_t LK_EMPTY = 0xFFFFFFFFFFFFFFFFULL;

enum CellOp { CELL_ADD_F64 = 0, CELL_ADD_I64 = 1, CELL_ADD_I128 = 2 /* this cell = lo, next = hi */,
              CELL_I128_HI = 3, CELL_MAXORD = 4 };

struct FastCell {
    int32_t op;          // CellOp
    int32_t col;         // index into FastParams::col
    uint32_t func_mask;  // aggregate functions this cell is flushed into
    int32_t ord;         // MAXORD: OrdKind
    int32_t is_min;      // MAXORD
    int32_t in_unsigned; // ADD_I128: zero- instead of sign-extend
};

struct FastParams {
    Pred pred;
    Col key;
    int key_mode;                 // 0: 8-byte raw bits, 1: int32 sign-extend, 2: 4-byte zero-extend
    int n_cols;
    Col col[FA_MAX_COLS];
    int col_mode[FA_MAX_COLS];    // 0: 8-byte raw, 1: int32 -> int64, 2: uint32 -> uint64, 3: float32 -> float64
    int n_cells;
    FastCell cell[FA_MAX_CELLS];
    int64_t n;
    int64_t num_tiles;
    int log2s;                    // shared key table slots = 1 << log2s
    int gmax;                     // dense group ids per CTA
    int64_t row_limit;            // row-level global inserts stop here (flush reserve above it)
    GTable table;
    ReplayList replay;
};

__host__ __device__ inline size_t fast_smem_bytes(int log2s, int gmax, int n_cells, int warps) {
    size_t S = (size_t) 1 << log2s;
    size_t table = S * 8 + S * 2;
    table = (table + 15) & ~(size_t) 15;
    size_t per_warp = (size_t) gmax * 4 + (size_t) n_cells * gmax * 8;
    return table + per_warp * warps;
}

template <int PK, int NV>
struct TileRegs {
    uint64_t key[FA_ROWS];
    uint64_t pred[(PK == PK_F64_VEC || PK == PK_I64_VEC) ? FA_ROWS : 1];
    uint64_t val[NV > 0 ? NV : 1][FA_ROWS];
    uint32_t flags;  // bit r: row r in range (and, for non-vector predicates, already selected)
};

__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) { return ((uint64_t) hi << 32) | lo; }

__device__ __forceinline__ uint64_t widen4(uint32_t raw, int mode) {
    if (mode == 1) return (uint64_t) (int64_t) (int32_t) raw;
    if (mode == 3) return (uint64_t) __double_as_longlong((double) __uint_as_float(raw));
    return raw;
}

template <int PK, int NV>
__device__ __forceinline__ void load_tile(const FastParams& p, int64_t tile, int tid, int nthreads,
                                          TileRegs<PK, NV>& t) {
    t.flags = 0;
    const int64_t base = tile * (int64_t) (nthreads * FA_ROWS);
#pragma unroll
    for (int j = 0; j < FA_ROWS / 2; ++j) {
        const int64_t r0 = base + (int64_t) j * (nthreads * 2) + tid * 2;
        const int a = 2 * j, b = 2 * j + 1;
        if (r0 + 1 < p.n) {
            if (p.key_mode == 0) {
                uint4 q = ldg_stream16(p.key.data + r0 * 8);
                t.key[a] = u64_of(q.x, q.y);
                t.key[b] = u64_of(q.z, q.w);
            } else {
                uint2 q = ldg_stream8(p.key.data + r0 * 4);
                t.key[a] = p.key_mode == 1 ? (uint64_t) (int64_t) (int32_t) q.x : (uint64_t) q.x;
                t.key[b] = p.key_mode == 1 ? (uint64_t) (int64_t) (int32_t) q.y : (uint64_t) q.y;
            }
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                if (p.col_mode[v] == 0) {
                    uint4 q = ldg_stream16(p.col[v].data + r0 * 8);
                    t.val[v][a] = u64_of(q.x, q.y);
                    t.val[v][b] = u64_of(q.z, q.w);
                } else {
                    uint2 q = ldg_stream8(p.col[v].data + r0 * 4);
                    t.val[v][a] = widen4(q.x, p.col_mode[v]);
                    t.val[v][b] = widen4(q.y, p.col_mode[v]);
                }
            }
            if constexpr (PK == PK_F64_VEC || PK == PK_I64_VEC) {
                uint4 q = ldg_stream16(p.pred.col.data + r0 * 8);
                t.pred[a] = u64_of(q.x, q.y);
                t.pred[b] = u64_of(q.z, q.w);
                t.flags |= 3u << a;
            } else {
                bool f0, f1;
                pred_pair<PK>(p.pred, r0, p.n, f0, f1);
                t.flags |= ((uint32_t) f0 << a) | ((uint32_t) f1 << b);
            }
        } else if (r0 < p.n) {
            // last odd row of the chunk
            t.key[a] = p.key_mode == 0 ? reinterpret_cast<const uint64_t*>(p.key.data)[r0]
                     : p.key_mode == 1 ? (uint64_t) (int64_t) reinterpret_cast<const int32_t*>(p.key.data)[r0]
                                       : (uint64_t) reinterpret_cast<const uint32_t*>(p.key.data)[r0];
            t.key[b] = 0;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                t.val[v][a] = p.col_mode[v] == 0 ? reinterpret_cast<const uint64_t*>(p.col[v].data)[r0]
                                                 : widen4(reinterpret_cast<const uint32_t*>(p.col[v].data)[r0], p.col_mode[v]);
                t.val[v][b] = 0;
            }
            if constexpr (PK == PK_F64_VEC || PK == PK_I64_VEC) {
                t.pred[a] = reinterpret_cast<const uint64_t*>(p.pred.col.data)[r0];
                t.pred[b] = 0;
                t.flags |= 1u << a;
            } else {
                bool f0, f1;
                pred_pair<PK>(p.pred, r0, p.n, f0, f1);
                t.flags |= (uint32_t) f0 << a;
            }
        }
    }
}

// Branch-free comparison: `sel_mask` has one bit per outcome {less, equal, greater, unordered}.
__host__ __device__ inline uint32_t cmp_outcome_mask(int op) {
    switch (op) {
        case VK_EQ: return 0b0010u;
        case VK_NE: return 0b1101u;
        case VK_GT: return 0b0100u;
        case VK_GE: return 0b0110u;
        case VK_LT: return 0b0001u;
        default: return 0b0011u;  // LE
    }
}
template <typename T>
__device__ __forceinline__ bool cmp_by_mask(uint32_t mask, T x, T c) {
    const uint32_t code = x < c ? 0u : (x == c ? 1u : (x > c ? 2u : 3u));
    return (mask >> code) & 1u;
}

__device__ __forceinline__ uint64_t cell_apply(const FastCell& c, uint64_t cur, uint64_t v) {
    switch (c.op) {
        case CELL_ADD_F64:
            return (uint64_t) __double_as_longlong(__longlong_as_double((long long) cur) +
                                                   __longlong_as_double((long long) v));
        case CELL_MAXORD: {
            uint64_t o = ord_transform(c.ord, c.is_min, v);
            return o > cur ? o : cur;
        }
        default: return cur + v;  // ADD_I64 and the low limb of ADD_I128
    }
}

// A row that does not go through the CTA-local table: straight to the global one.
template <int NV>
__device__ __forceinline__ void fast_global_row(const FastParams& p, uint64_t key, const uint64_t* vals, int64_t row) {
    int64_t g = gt1_find_or_insert(p.table, key, false, hash_key1(key), p.row_limit);
    if (g < 0) {
        replay_append(p.replay, row);
        return;
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(p.table.count_star + g), 1ULL);
    for (int c = 0; c < p.n_cells; ++c) {
        const FastCell cell = p.cell[c];
        if (cell.op == CELL_I128_HI) continue;
        uint64_t v = 0;
#pragma unroll
        for (int k = 0; k < NV; ++k)
            if (cell.col == k) v = vals[k];
        uint32_t fm = cell.func_mask;
        while (fm) {
            const int fi = __ffs(fm) - 1;
            fm &= fm - 1;
            switch (cell.op) {
                case CELL_ADD_F64:
                    atomicAdd(reinterpret_cast<double*>(p.table.acc_lo[fi] + g), __longlong_as_double((long long) v));
                    break;
                case CELL_ADD_I64:
                    atomicAdd(reinterpret_cast<unsigned long long*>(p.table.acc_lo[fi] + g), (unsigned long long) v);
                    break;
                case CELL_ADD_I128:
                    acc_add_i128(p.table.acc_lo[fi] + g, p.table.acc_hi[fi] + g, v,
                                 (!cell.in_unsigned && (int64_t) v < 0) ? ~0ULL : 0ULL);
                    break;
                default:
                    atomicMax(reinterpret_cast<unsigned long long*>(p.table.acc_lo[fi] + g),
                              (unsigned long long) ord_transform(cell.ord, cell.is_min, v));
                    break;
            }
        }
    }
}

template <int PK, int NV>
__global__ void __launch_bounds__(FA_MAX_WARPS * 32, 1) agg_fast_kernel(const __grid_constant__ FastParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_ngroups;
    const int S = 1 << p.log2s;
    const uint32_t smask = S - 1;
    const int G = p.gmax;
    const int nthreads = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const int ncells = p.n_cells;

    uint64_t* s_keys = reinterpret_cast<uint64_t*>(smem);
    uint16_t* s_gid = reinterpret_cast<uint16_t*>(smem + (size_t) S * 8);
    const size_t table_bytes = ((size_t) S * 10 + 15) & ~(size_t) 15;
    const size_t per_warp = (size_t) G * 4 + (size_t) ncells * G * 8;
    uint8_t* acc_base = smem + table_bytes;
    uint64_t* my_cells = reinterpret_cast<uint64_t*>(acc_base + (size_t) warp * per_warp);  // [ncells][G]
    uint32_t* my_cnt = reinterpret_cast<uint32_t*>(acc_base + (size_t) warp * per_warp + (size_t) ncells * G * 8);

    for (int i = tid; i < S; i += nthreads) {
        s_keys[i] = LK_EMPTY;
        s_gid[i] = (uint16_t) GID_PENDING;
    }
    {
        uint32_t* z = reinterpret_cast<uint32_t*>(acc_base);
        const size_t words = per_warp * nwarps / 4;
        for (size_t i = tid; i < words; i += nthreads) z[i] = 0;
    }
    if (tid == 0) s_ngroups = 0;
    __syncthreads();

    const unsigned lt = lanemask_lt();
    const uint32_t opmask = cmp_outcome_mask(p.pred.op);
    const uint64_t pscalar = p.pred.scalar.bits;
    uint32_t spilled = 0;

    TileRegs<PK, NV> cur, nxt;
    int64_t tile = blockIdx.x;
    if (tile < p.num_tiles) load_tile<PK, NV>(p, tile, tid, nthreads, cur);
    for (; tile < p.num_tiles; tile += gridDim.x) {
        __syncwarp();
        const int64_t tnext = tile + gridDim.x;
        if (tnext < p.num_tiles) load_tile<PK, NV>(p, tnext, tid, nthreads, nxt);

#pragma unroll
        for (int r = 0; r < FA_ROWS; ++r) {
            // ---- predicate (registers only) ----
            bool act = (cur.flags >> r) & 1;
            if constexpr (PK == PK_F64_VEC)
                act = act && cmp_by_mask(opmask, __longlong_as_double((long long) cur.pred[r]),
                                         __longlong_as_double((long long) pscalar));
            else if constexpr (PK == PK_I64_VEC)
                act = act && cmp_by_mask(opmask, (int64_t) cur.pred[r], (int64_t) pscalar);
            const uint64_t key = cur.key[r];

            // ---- key -> dense group id: warp-convergent probe loop ----
            uint64_t x = key ^ (key >> 29);
            x *= 0x9E3779B97F4A7C15ULL;
            const uint32_t hh = (uint32_t) (x >> 32);
            uint32_t h = hh >> (32 - p.log2s);
            const uint32_t step = ((hh << 1) | 1u) & smask;
            uint32_t gid = GID_SPILL;
            bool searching = act && key != LK_EMPTY;
            for (int it = 0; it < 4 * 64; ++it) {
                if (!__any_sync(0xffffffffu, searching)) break;
                if (searching) {
                    const uint64_t k = reinterpret_cast<volatile uint64_t*>(s_keys)[h];
                    if (k == key) {
                        const uint32_t g = reinterpret_cast<volatile uint16_t*>(s_gid)[h];
                        if (g != GID_PENDING) {
                            gid = g;
                            searching = false;
                        }
                    } else if (k == LK_EMPTY) {
                        const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(s_keys + h),
                                                                 (unsigned long long) LK_EMPTY, (unsigned long long) key);
                        if (old == LK_EMPTY) {
                            uint32_t g = atomicAdd(&s_ngroups, 1u);
                            if (g >= (uint32_t) G) g = GID_SPILL;
                            reinterpret_cast<volatile uint16_t*>(s_gid)[h] = (uint16_t) g;
                            gid = g;
                            searching = false;
                        } else if (old != key) {
                            h = (h + step) & smask;
                        }  // old == key: another lane just inserted it; re-read its id next round
                    } else {
                        h = (h + step) & smask;
                    }
                }
                __syncwarp();
            }
            const bool upd = act && !searching && gid < (uint32_t) G;
            const bool spill = act && !upd;

            // ---- accumulate into the warp-private arrays; same-group lanes take turns ----
            const unsigned peers = __match_any_sync(0xffffffffu, upd ? gid : (0x10000u | (unsigned) lane));
            const int mult = upd ? __popc(peers) : 0;
            const int rank = __popc(peers & lt);
            const int maxm = __reduce_max_sync(0xffffffffu, mult);
            if (upd && rank == 0) my_cnt[gid] += (uint32_t) mult;
            for (int round = 0; round < maxm; ++round) {
                if (upd && rank == round) {
                    for (int c = 0; c < ncells; ++c) {
                        const FastCell cell = p.cell[c];
                        if (cell.op == CELL_I128_HI) continue;
                        uint64_t v = 0;
#pragma unroll
                        for (int k = 0; k < NV; ++k)
                            if (cell.col == k) v = cur.val[k][r];
                        uint64_t* slot = my_cells + (size_t) c * G + gid;
                        const uint64_t old = *slot;
                        const uint64_t nw = cell_apply(cell, old, v);
                        *slot = nw;
                        if (cell.op == CELL_ADD_I128) {
                            const uint64_t ext = (!cell.in_unsigned && (int64_t) v < 0) ? ~0ULL : 0ULL;
                            slot[G] += ext + (nw < old ? 1ULL : 0ULL);
                        }
                    }
                }
                if (maxm > 1) __syncwarp();
            }

            // ---- rows the CTA table could not take: global table, off the hot path ----
            if (__any_sync(0xffffffffu, spill)) {
                if (spill) {
                    uint64_t vals[NV > 0 ? NV : 1];
#pragma unroll
                    for (int k = 0; k < NV; ++k) vals[k] = cur.val[k][r];
                    const int64_t row = tile * (int64_t) (nthreads * FA_ROWS) + (int64_t) (r >> 1) * (nthreads * 2) +
                                        tid * 2 + (r & 1);
                    ++spilled;
                    fast_global_row<NV>(p, key, vals, row);
                }
                __syncwarp();
            }
        }
        cur = nxt;
    }

    // ---- flush: combine the warps' private accumulators, one global update per group ----
    __syncthreads();
    for (int d = 16; d > 0; d >>= 1) spilled += __shfl_xor_sync(0xffffffffu, spilled, d);
    if (lane == 0 && spilled) atomicAdd(p.replay.spilled, (unsigned long long) spilled);
    for (int s = tid; s < S; s += nthreads) {
        const uint64_t key = s_keys[s];
        if (key == LK_EMPTY) continue;
        const uint32_t gid = s_gid[s];
        if (gid >= (uint32_t) G) continue;
        uint64_t cnt = 0;
        for (int w = 0; w < nwarps; ++w)
            cnt += reinterpret_cast<const uint32_t*>(acc_base + (size_t) w * per_warp + (size_t) ncells * G * 8)[gid];
        if (cnt == 0) continue;
        // the host reserves capacity for every CTA's groups: this insert cannot fail
        const int64_t g = gt1_find_or_insert(p.table, key, false, hash_key1(key), INT64_MAX);
        if (g < 0) {
            atomicAdd(p.replay.lost, (unsigned long long) cnt);
            continue;
        }
        atomicAdd(reinterpret_cast<unsigned long long*>(p.table.count_star + g), (unsigned long long) cnt);
        for (int c = 0; c < ncells; ++c) {
            const FastCell cell = p.cell[c];
            if (cell.op == CELL_I128_HI) continue;
            uint64_t lo = 0, hi = 0;
            double fsum = 0.0;
            for (int w = 0; w < nwarps; ++w) {
                const uint64_t* cells = reinterpret_cast<const uint64_t*>(acc_base + (size_t) w * per_warp);
                const uint64_t v = cells[(size_t) c * G + gid];
                if (cell.op == CELL_ADD_F64) fsum += __longlong_as_double((long long) v);
                else if (cell.op == CELL_MAXORD) lo = v > lo ? v : lo;
                else {
                    const uint64_t nl = lo + v;
                    if (cell.op == CELL_ADD_I128) hi += cells[(size_t) (c + 1) * G + gid] + (nl < lo ? 1ULL : 0ULL);
                    lo = nl;
                }
            }
            uint32_t fm = cell.func_mask;
            while (fm) {
                const int fi = __ffs(fm) - 1;
                fm &= fm - 1;
                switch (cell.op) {
                    case CELL_ADD_F64: atomicAdd(reinterpret_cast<double*>(p.table.acc_lo[fi] + g), fsum); break;
                    case CELL_ADD_I64:
                        atomicAdd(reinterpret_cast<unsigned long long*>(p.table.acc_lo[fi] + g), (unsigned long long) lo);
                        break;
                    case CELL_ADD_I128: acc_add_i128(p.table.acc_lo[fi] + g, p.table.acc_hi[fi] + g, lo, hi); break;
                    default:
                        atomicMax(reinterpret_cast<unsigned long long*>(p.table.acc_lo[fi] + g), (unsigned long long) lo);
                        break;
                }
            }
        }
    }
}

}  // namespace vk
