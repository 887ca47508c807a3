// agg_fast_kernel -- fused filter -> hash aggregate for low-cardinality single-key
// group-bys (the north-star pipeline).
//
// Measured on B200 (scripts/ubench/smem_prims.cu, profiles/): MATCH.ANY costs 64 issue
// cycles per warp, a shared 64-bit CAS ~11, a shared f64 atomicAdd (CAS loop) ~20, while a
// random LDS.128 + STS.128 pair costs ~12-19 and a one-byte STS + LDS pair ~6.  The per-row
// path below therefore contains no atomic and no MATCH:
//
//  * one persistent CTA per SM; rows stream through registers with 16-byte loads that stay
//    RAW (no instruction touches a loaded register until the tile is processed, so the loads
//    of tile k+1 are in flight during all of tile k); the WHERE predicate is evaluated in
//    registers and never materialised;
//  * the key is mapped to a dense group id either DIRECTLY (key - base < G: dense integer
//    domains such as dictionary codes, learned by the host from the first chunk) or through
//    a CTA-shared open-addressing table that is only READ on the per-row path (an 8-byte CAS
//    runs once per distinct key per CTA);
//  * COUNT and the accumulator cells of a group live in WARP-PRIVATE 16/32-byte entries
//    updated with plain LDS.128 / STS.128.  Lanes of a warp that hit the same group in the
//    same step are serialised by tag arbitration: every lane stores its lane id in a byte
//    array indexed by group, the warp syncs, the lane that reads its own id back owns the
//    group for this round, the others retry;
//  * every loop on the path is warp-convergent (runs until __any_sync says no lane is
//    pending), so a warp never splits into sub-warps.
//
// At the end every CTA folds its warps' entries and flushes one update per group into the
// global table.  Keys that do not fit the CTA's table go to the global table row by row.
#pragma once
#include "vk_hashagg.cuh"

namespace vk {

#ifndef VK_FA_MAXT
#define VK_FA_MAXT 384
#endif
constexpr int FA_MAX_THREADS = VK_FA_MAXT;
constexpr int FA_R = 8;            // rows per thread per tile (four lane-contiguous pairs)
#ifndef VK_FA_K
#define VK_FA_K 2
#endif
constexpr int FA_K = VK_FA_K;      // rows per arbitration round in phase 2 (2, 4 or 8)
constexpr int FA_MAX_COLS = 3;     // distinct value columns
constexpr int FA_MAX_CELLS = 3;    // 64-bit accumulator cells per group (besides COUNT)
constexpr int FA_MAXPROBE = 64;
constexpr uint32_t GID_PENDING = 0xFFFFu;  // key claimed, dense id not published yet
constexpr uint32_t GID_SPILL = 0xFFFEu;    // more groups than the CTA holds: rows go to the global table
constexpr uint64_t LK_EMPTY = 0xFFFFFFFFFFFFFFFFULL;

enum CellOp { CELL_ADD_F64 = 0, CELL_ADD_I64 = 1, CELL_ADD_I128 = 2 /* this cell = lo, next = hi */,
              CELL_I128_HI = 3, CELL_MAXORD = 4 };

// Column layout specialisations (compile time): what a loaded register pair means.
enum FastMode {
    FM_ALL8 = 0,     // key and every value column are 8-byte raw
    FM_KEY4 = 1,     // 4-byte key (int32 sign- or uint32/float32 zero-extended), 8-byte values
    FM_RUNTIME = 2   // per-column modes decided at run time (key_mode / col_mode)
};

struct FastCell {
    int32_t op;           // CellOp
    int32_t col;          // index into FastParams::col
    uint32_t func_mask;   // aggregate functions this cell is flushed into
    int32_t ord;          // MAXORD: OrdKind
    int32_t is_min;       // MAXORD
    int32_t in_unsigned;  // ADD_I128: zero- instead of sign-extend
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
    int log2s;                    // shared key table slots = 1 << log2s (hash mode)
    int gmax;                     // dense group ids per CTA
    uint64_t direct_base;         // direct mode: gid = key - direct_base
    int64_t row_limit;            // row-level global inserts stop here (flush reserve above it)
    int pf_dist;                  // L2 bulk prefetch distance in tiles (0 = off)
    // hash mode with a host-built dictionary (read-only in the kernel): cuckoo placement of the keys the
    // learning launch found, two hash functions, S = 1 << log2s slots; nullptr = insert-as-you-go table
    const uint64_t* dict_keys;    // [S], LK_EMPTY = free
    const uint16_t* dict_gids;    // [S] dense group id of the slot's key
    uint64_t seed_a;              // seed of the key fold (dict_fold)
    long long* key_range;         // optional {min, max} (signed order) of the keys this launch flushes
    GTable table;
    ReplayList replay;
};

// Dynamic shared memory: [hash mode: keys[S] u64 | gid[S] u16] | per warp entries[G][NW] u64.
__host__ __device__ inline size_t fast_table_bytes(int log2s, bool direct) { return direct ? 0 : ((size_t) 10 << log2s); }
__host__ __device__ inline size_t fast_warp_bytes(int gmax, int nw) { return (size_t) gmax * nw * 8; }
__host__ __device__ inline size_t fast_smem_bytes(int log2s, bool direct, int gmax, int nw, int warps) {
    return fast_table_bytes(log2s, direct) + fast_warp_bytes(gmax, nw) * warps;
}

// ---- shared-memory accessors (explicit so that nothing is cached in registers) ---------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t lds64(uint32_t a) {
    uint64_t v;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts16(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts64(uint32_t a, uint64_t v) {
    asm volatile("st.shared.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
__device__ __forceinline__ void lds128(uint32_t a, uint32_t (&e)[4]) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]) : "r"(a) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t a, const uint32_t (&e)[4]) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(e[0]), "r"(e[1]), "r"(e[2]), "r"(e[3]) : "memory");
}
__device__ __forceinline__ uint32_t ldg_stream2(const void* p) {
    uint16_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}

__device__ __forceinline__ uint64_t u64_of(uint32_t lo, uint32_t hi) { return ((uint64_t) hi << 32) | lo; }

// Asynchronous bulk prefetch of [p, p + bytes) into L2 (no register, no shared memory): it
// keeps more bytes in flight than the register tiles alone can.
__device__ __forceinline__ void l2_prefetch_bulk(const void* ptr, uint32_t bytes) {
    const uint64_t a = reinterpret_cast<uint64_t>(ptr) & ~(uint64_t) 15;
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"(bytes) : "memory");
}

// ---- tile registers: RAW load results, decoded only when the tile is processed ----------
template <int PK, int NV>
struct RawTile {
    uint4 kq[FA_R / 2];
    uint4 pq[(PK == PK_F64_VEC || PK == PK_I64_VEC || PK == PK_MASK) ? FA_R / 2 : 1];
    uint4 vq[NV > 0 ? NV : 1][FA_R / 2];
    uint32_t flags;  // bit r: row r exists (and, for PK_GENERIC, is already selected)
};

// Row pairs are lane-contiguous: pair j of thread t covers rows base + j*2*threads + 2t, +1,
// so every 8-byte column is read with one 16-byte load per lane per pair.  Only complete
// tiles are loaded here (the host sends the ragged tail of a chunk to agg_general_kernel).
//
// The key / predicate registers of a tile are dead after phase 1 and the value registers of
// pair j after its two rows have been accumulated, so each is refilled with the tile after
// next as soon as it dies: every load has almost two tile periods to land.
template <int PK, int NV, int MODE>
__device__ __forceinline__ void load_tile_keys(const FastParams& p, int64_t tile, int tid, int nthreads, RawTile<PK, NV>& t) {
    const int64_t r0 = tile * (int64_t) (nthreads * FA_R) + tid * 2;
    const int64_t jstride = (int64_t) nthreads * 2;
    t.flags = (1u << FA_R) - 1u;
#pragma unroll
    for (int j = 0; j < FA_R / 2; ++j) {
        const int64_t r = r0 + j * jstride;
        if (MODE == FM_ALL8 || (MODE == FM_RUNTIME && p.key_mode == 0)) {
            t.kq[j] = ldg_stream16(p.key.data + r * 8);
        } else {
            const uint2 q = ldg_stream8(p.key.data + r * 4);
            t.kq[j].x = q.x;
            t.kq[j].y = q.y;
        }
        if constexpr (PK == PK_F64_VEC || PK == PK_I64_VEC) {
            t.pq[j] = ldg_stream16(p.pred.col.data + r * 8);
        } else if constexpr (PK == PK_MASK) {
            t.pq[j].x = ldg_stream2(p.pred.mask + r);  // two mask bytes; r is even, the base 2-byte aligned
        } else if constexpr (PK == PK_GENERIC) {
            bool f0, f1;
            pred_pair<PK>(p.pred, r, p.n, f0, f1);
            if (!f0) t.flags &= ~(1u << (2 * j));
            if (!f1) t.flags &= ~(2u << (2 * j));
        }
    }
}
template <int PK, int NV, int MODE>
__device__ __forceinline__ void load_tile_vals(const FastParams& p, int64_t tile, int tid, int nthreads, RawTile<PK, NV>& t, int j) {
    const int64_t r = tile * (int64_t) (nthreads * FA_R) + tid * 2 + (int64_t) j * (nthreads * 2);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        if (MODE != FM_RUNTIME || p.col_mode[v] == 0) {
            t.vq[v][j] = ldg_stream16(p.col[v].data + r * 8);
        } else {
            const uint2 q = ldg_stream8(p.col[v].data + r * 4);
            t.vq[v][j].x = q.x;
            t.vq[v][j].y = q.y;
        }
    }
}

// ---- decoding a row of a raw tile ---------------------------------------------------------
__device__ __forceinline__ uint64_t widen4(uint32_t raw, int mode) {
    if (mode == 1) return (uint64_t) (int64_t) (int32_t) raw;
    if (mode == 3) return (uint64_t) __double_as_longlong((double) __uint_as_float(raw));
    return raw;
}
// `r` is a compile-time constant after unrolling: these selects fold away.
__device__ __forceinline__ uint64_t q8(const uint4 (&q)[FA_R / 2], int r) {  // 8-byte element of row r
    return (r & 1) ? u64_of(q[r >> 1].z, q[r >> 1].w) : u64_of(q[r >> 1].x, q[r >> 1].y);
}
__device__ __forceinline__ uint32_t q4(const uint4 (&q)[FA_R / 2], int r) {  // 4-byte element of row r
    return (r & 1) ? q[r >> 1].y : q[r >> 1].x;
}
template <int MODE>
__device__ __forceinline__ uint64_t row_key(const FastParams& p, const uint4 (&kq)[FA_R / 2], int r) {
    if constexpr (MODE == FM_ALL8) return q8(kq, r);
    else if constexpr (MODE == FM_KEY4) {
        const uint32_t raw = q4(kq, r);
        const uint32_t hi = p.key_mode == 1 ? (uint32_t) ((int32_t) raw >> 31) : 0u;
        return u64_of(raw, hi);
    } else {
        if (p.key_mode == 0) return q8(kq, r);
        return widen4(q4(kq, r), p.key_mode == 1 ? 1 : 2);
    }
}
template <int MODE>
__device__ __forceinline__ uint64_t row_val(const FastParams& p, int v, const uint4 (&vq)[FA_R / 2], int r) {
    if constexpr (MODE != FM_RUNTIME) return q8(vq, r);
    else {
        if (p.col_mode[v] == 0) return q8(vq, r);
        return widen4(q4(vq, r), p.col_mode[v]);
    }
}

__device__ __forceinline__ uint64_t cell_apply(const FastCell& c, uint64_t cur, uint64_t v) {
    switch (c.op) {
        case CELL_ADD_F64:
            return (uint64_t) __double_as_longlong(__longlong_as_double((long long) cur) +
                                                   __longlong_as_double((long long) v));
        case CELL_MAXORD: {
            const uint64_t o = ord_transform(c.ord, c.is_min, v);
            return o > cur ? o : cur;
        }
        default: return cur + v;  // ADD_I64 and the low limb of ADD_I128
    }
}

__host__ __device__ __forceinline__ uint32_t fast_hash32(uint64_t key) {
    uint32_t x = (uint32_t) key ^ ((uint32_t) (key >> 32) * 0x85EBCA77u);
    x ^= x >> 16;
    return x * 0x9E3779B1u;  // Fibonacci hashing: the TOP bits index the table
}
// The dictionary's two slot choices for a key (cuckoo hashing; the host retries other seeds on a cycle).
// Both come from ONE 32-bit fold of the key (3 instructions) and one multiply each: the lookup runs per
// selected row, and the first version's two independent 64-bit hashes cost 1.6x the instructions of the
// direct-id kernel (profiles/r02_agg_fast_dict_ncu_full.md).
__host__ __device__ __forceinline__ uint32_t dict_fold(uint64_t key, uint64_t seed) {
    return ((uint32_t) key ^ (uint32_t) seed) ^ (((uint32_t) (key >> 32) ^ (uint32_t) (seed >> 32)) * 0x85EBCA77u);
}
__host__ __device__ __forceinline__ uint32_t dict_slot_a(uint32_t x, uint32_t shift) { return ((x ^ (x >> 16)) * 0x9E3779B1u) >> shift; }
__host__ __device__ __forceinline__ uint32_t dict_slot_b(uint32_t x, uint32_t shift) { return ((x ^ (x >> 13)) * 0xC2B2AE35u + 0x27D4EB2Fu) >> shift; }

// One folded group of a CTA (or one spilled row) into the global table.
__device__ __forceinline__ void fast_global_update(const FastParams& p, int64_t g, uint64_t cnt, const uint64_t* w /*cells*/) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p.table.count_star + g), (unsigned long long) cnt);
    for (int c = 0; c < p.n_cells; ++c) {
        const FastCell cell = p.cell[c];
        if (cell.op == CELL_I128_HI) continue;
        uint32_t fm = cell.func_mask;
        while (fm) {
            const int fi = __ffs(fm) - 1;
            fm &= fm - 1;
            switch (cell.op) {
                case CELL_ADD_F64:
                    atomicAdd(reinterpret_cast<double*>(p.table.acc_lo[fi] + g), __longlong_as_double((long long) w[c]));
                    break;
                case CELL_ADD_I64:
                    atomicAdd(reinterpret_cast<unsigned long long*>(p.table.acc_lo[fi] + g), (unsigned long long) w[c]);
                    break;
                case CELL_ADD_I128:
                    acc_add_i128(p.table.acc_lo[fi] + g, p.table.acc_hi[fi] + g, w[c], w[c + 1 < FA_MAX_CELLS ? c + 1 : c]);
                    break;
                default:
                    atomicMax(reinterpret_cast<unsigned long long*>(p.table.acc_lo[fi] + g), (unsigned long long) w[c]);
                    break;
            }
        }
    }
}

// A row that does not go through the CTA-local table: straight to the global one.
static __device__ __noinline__ void fast_global_row(const FastParams& p, uint64_t key, uint64_t v0, uint64_t v1, uint64_t v2,
                                             int64_t row) {
    const int64_t g = gt1_find_or_insert(p.table, key, false, hash_key1(key), p.row_limit);
    if (g < 0) {
        replay_append(p.replay, row);
        return;
    }
    const uint64_t vals[FA_MAX_COLS] = {v0, v1, v2};
    uint64_t w[FA_MAX_CELLS] = {0, 0, 0};
    for (int c = 0; c < p.n_cells; ++c) {
        const FastCell cell = p.cell[c];
        if (cell.op == CELL_I128_HI) {
            const FastCell lo = p.cell[c - 1];
            w[c] = (!lo.in_unsigned && (int64_t) vals[lo.col] < 0) ? ~0ULL : 0ULL;
        } else if (cell.op == CELL_MAXORD) {
            w[c] = ord_transform(cell.ord, cell.is_min, vals[cell.col]);
        } else {
            w[c] = vals[cell.col];
        }
    }
    fast_global_update(p, g, 1, w);
}

// Kernel context shared by the per-tile steps.
struct FastCtx {
    uint32_t a_keys, a_gid;  // shared key table (hash mode)
    uint32_t a_ent;          // this warp's entries
    uint32_t smask, hshift;
    int op;                  // VkCmpOp of the fused predicate
    uint64_t pscalar;
    uint32_t lane;
};

// ---- a hot group leaves the arbitration in one step ------------------------------------------------------
// Tag arbitration admits ONE lane per entry per round: a key that most lanes of a warp hold costs up to 32
// rounds of shared-memory traffic per row (26 Grows/s with one key against 193 uniform, profiles/r02_skew.md).
// Before a row slot arbitrates, the lanes that share the entry of the LOWEST pending lane are counted with one
// ballot; if they are FA_HOT_MIN or more, their values are summed with a warp butterfly (lanes outside the group
// add 0.0) and that lane alone updates the entry for all of them.  Uniform keys pay two votes and a shuffle per
// row slot and never take the branch -- but even that costs the HBM-bound north-star kernel 50 % (7.03 against
// 4.71 ms), so the step is a template parameter and the host picks the kernel that has it only when the learning
// launch saw one key hold >= 30 % of the rows (option AGG_HOT: 1 = that rule, 2 = always, 0 = never).
constexpr int FA_HOT_MIN = 6;
template <int MODE, int PK, int NV, int VAR>
__device__ __forceinline__ void fast_hot_group(const FastParams& p, const FastCtx& cx, const RawTile<PK, NV>& t, uint32_t ea_r,
                                               int r, uint32_t& pend, int k) {
    const bool on = (pend >> k) & 1u;
    const uint32_t m_on = __ballot_sync(0xffffffffu, on);
    if (m_on == 0u) return;
    const int l0 = __ffs(m_on) - 1;
    const uint32_t ea0 = __shfl_sync(0xffffffffu, ea_r, l0);
    const uint32_t grp = __ballot_sync(0xffffffffu, on && ea_r == ea0);
    if (__popc(grp) < FA_HOT_MIN) return;
    const bool mine = (grp >> cx.lane) & 1u;
    double v = mine ? __longlong_as_double((long long) row_val<MODE>(p, 0, t.vq[0], r)) : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    if ((int) cx.lane == l0) {
        if constexpr (VAR == 2) {
            const uint32_t g = (ea0 - cx.a_ent) >> 4;
            const uint32_t a_sum = cx.a_ent + g * 8u, a_ct = cx.a_ent + (uint32_t) p.gmax * 8u + g * 4u;
            const double sum = __longlong_as_double((long long) lds64(a_sum)) + v;
            sts64(a_sum, (uint64_t) __double_as_longlong(sum));
            sts32(a_ct, (lds32(a_ct) + (uint32_t) __popc(grp)) & 0x00FFFFFFu);
        } else {
            uint32_t e[4];
            lds128(ea0, e);
            e[0] += (uint32_t) __popc(grp);
            const double sum = __hiloint2double((int) e[3], (int) e[2]) + v;
            e[2] = (uint32_t) __double2loint(sum);
            e[3] = (uint32_t) __double2hiint(sum);
            sts128(ea0, e);
        }
    }
    if (mine) pend &= ~(1u << k);
    __syncwarp();
}

// Entry layout (NW 64-bit words, NW = 2 or 4):
//   word 0 = COUNT (low 32 bits: one warp in one launch) | arbitration tag (high 32 bits)
//   word 1.. = accumulator cells
// SUMF64: the entry's single cell is a float64 SUM (no run-time cell dispatch).
//
// A tile is processed in two phases.  Phase 1 is pure register work for all FA_R rows
// (decode, predicate, group id or first table probe) that the compiler interleaves freely.
// Phase 2 is one short critical section per row: store the lane id as the entry's tag,
// sync the warp, load the whole entry (tag + COUNT + cell in one LDS.128); the lane that
// reads its own tag back applies the row and stores the entry, the others go round again.
template <int PK, int NV, int NW, int MODE, bool DIRECT, bool SUMF64, int VAR, bool HOT>
__device__ __forceinline__ void fast_tile(const FastParams& p, const FastCtx& cx, RawTile<PK, NV>& t, uint8_t* smem,
                                          uint32_t* s_ngroups, int64_t row0, int tid, int nthreads, int64_t refill_tile,
                                          uint32_t& spilled) {
    constexpr int NCMAX = NW - 1;
    uint64_t key[FA_R];
    uint32_t ea[FA_R];       // shared address of the row's entry
    uint32_t todo = 0, spill = 0;

    // ---- phase 1: predicate and group id ----
    uint32_t okmask = (1u << FA_R) - 1u;
    if constexpr (PK == PK_F64_VEC || PK == PK_I64_VEC) {
        // one compare per row: the operator is uniform, so branch on it once per tile
        okmask = 0;
#define VK_PRED_ROWS(OP)                                                                                  \
        _Pragma("unroll") for (int r = 0; r < FA_R; ++r) {                                                \
            bool ok;                                                                                      \
            if constexpr (PK == PK_F64_VEC)                                                               \
                ok = __longlong_as_double((long long) q8(t.pq, r)) OP __longlong_as_double((long long) cx.pscalar); \
            else                                                                                          \
                ok = (int64_t) q8(t.pq, r) OP (int64_t) cx.pscalar;                                       \
            okmask |= (uint32_t) ok << r;                                                                 \
        }
        switch (cx.op) {
            case VK_EQ: VK_PRED_ROWS(==) break;
            case VK_NE: VK_PRED_ROWS(!=) break;
            case VK_GT: VK_PRED_ROWS(>) break;
            case VK_GE: VK_PRED_ROWS(>=) break;
            case VK_LT: VK_PRED_ROWS(<) break;
            default: VK_PRED_ROWS(<=) break;
        }
#undef VK_PRED_ROWS
    } else if constexpr (PK == PK_MASK) {
        okmask = 0;
#pragma unroll
        for (int r = 0; r < FA_R; ++r)
            okmask |= (uint32_t) (((t.pq[r >> 1].x >> (8 * (r & 1))) & 0xffu) != 0) << r;
    }
    const uint32_t act = t.flags & okmask;
#pragma unroll
    for (int r = 0; r < FA_R; ++r) key[r] = row_key<MODE>(p, t.kq, r);
    if constexpr (DIRECT) {
        uint32_t inmask = 0;
        const uint32_t g32 = (uint32_t) p.gmax;
#pragma unroll
        for (int r = 0; r < FA_R; ++r) {
            const uint64_t d = key[r] - p.direct_base;
            const uint32_t dl = (uint32_t) d, dh = (uint32_t) (d >> 32);
            inmask |= (uint32_t) ((dh == 0u) & (dl < g32)) << r;
            ea[r] = cx.a_ent + dl * (NW * 8);
        }
        todo = act & inmask;
        spill = act & ~inmask;
    } else if (p.dict_keys != nullptr) {
        // Read-only dictionary: a key the learning launch saw sits in one of its two slots, so the lookup
        // is two independent 8-byte probes + one 2-byte id load per row and NO loop -- with the
        // insert-as-you-go table below, one displaced key anywhere in a warp's 256 rows sends the whole
        // warp through the convergent probe loop (measured 4.7x slower than direct ids at 1000 keys).
        // A key that is in neither slot is new: that row goes to the global table.
        uint32_t hit = 0;
#pragma unroll
        for (int r = 0; r < FA_R; ++r) {
            ea[r] = cx.a_ent;
            if ((act >> r) & 1u) {
                const uint32_t x = dict_fold(key[r], p.seed_a);
                const uint32_t ha = dict_slot_a(x, cx.hshift), hb = dict_slot_b(x, cx.hshift);
                const uint64_t ka = lds64(cx.a_keys + ha * 8), kb = lds64(cx.a_keys + hb * 8);
                const bool in_a = ka == key[r], in_b = kb == key[r];
                const uint32_t g = lds16(cx.a_gid + (in_a ? ha : hb) * 2);
                if ((in_a | in_b) && key[r] != LK_EMPTY) {
                    hit |= 1u << r;
                    ea[r] = cx.a_ent + g * (NW * 8);
                }
            }
        }
        todo = act & hit;
        spill = act & ~hit;
    } else {
        uint32_t h[FA_R], gid[FA_R];
        uint32_t pend = 0;
        // first probe of every row, all loads in flight together
#pragma unroll
        for (int r = 0; r < FA_R; ++r) {
            h[r] = fast_hash32(key[r]) >> cx.hshift;
            gid[r] = GID_SPILL;
            if ((act >> r) & 1u) {
                if (key[r] == LK_EMPTY) {
                    spill |= 1u << r;  // the sentinel itself cannot live in the table
                } else {
                    const uint64_t k = lds64(cx.a_keys + h[r] * 8);
                    const uint32_t g = lds16(cx.a_gid + h[r] * 2);
                    if (k == key[r] && g != GID_PENDING) gid[r] = g;
                    else pend |= 1u << r;
                }
            }
        }
        // rows whose first probe did not hit: walk / insert, warp-convergent
        for (int it = 0; __any_sync(0xffffffffu, pend != 0); ++it) {
            if (it >= FA_MAXPROBE) {  // table region exhausted: leave these rows to the global table
                spill |= pend;
                pend = 0;
                break;
            }
#pragma unroll
            for (int r = 0; r < FA_R; ++r) {
                if (__any_sync(0xffffffffu, (pend >> r) & 1u)) {
                    if ((pend >> r) & 1u) {
                        const uint64_t k = lds64(cx.a_keys + h[r] * 8);
                        const uint32_t g = lds16(cx.a_gid + h[r] * 2);
                        if (k == key[r]) {
                            if (g != GID_PENDING) {
                                gid[r] = g;
                                pend &= ~(1u << r);
                            }
                        } else if (k == LK_EMPTY) {
                            const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(smem) + h[r],
                                                                     (unsigned long long) LK_EMPTY, (unsigned long long) key[r]);
                            if (old == LK_EMPTY) {
                                uint32_t ng = atomicAdd(s_ngroups, 1u);
                                if (ng >= (uint32_t) p.gmax) ng = GID_SPILL;
                                sts16(cx.a_gid + h[r] * 2, ng);
                                gid[r] = ng;
                                pend &= ~(1u << r);
                            } else if (old != key[r]) {
                                h[r] = (h[r] + 1) & cx.smask;
                            }  // old == key: another lane just claimed it; read its id next round
                        } else {
                            h[r] = (h[r] + 1) & cx.smask;
                        }
                    }
                    __syncwarp();
                }
            }
        }
#pragma unroll
        for (int r = 0; r < FA_R; ++r) {
            ea[r] = cx.a_ent + gid[r] * (NW * 8);
            if (((act & ~spill) >> r) & 1u) {
                if (gid[r] < (uint32_t) p.gmax) todo |= 1u << r;
                else spill |= 1u << r;
            }
        }
    }

    // ---- rows the CTA table could not take: global table, off the hot path ----
    if (__any_sync(0xffffffffu, spill != 0)) {
#pragma unroll
        for (int r = 0; r < FA_R; ++r) {
            if ((spill >> r) & 1u) {
                uint64_t v[FA_MAX_COLS] = {0, 0, 0};
#pragma unroll
                for (int k = 0; k < NV; ++k) v[k] = row_val<MODE>(p, k, t.vq[k], r);
                ++spilled;
                fast_global_row(p, key[r], v[0], v[1], v[2], row0 + (int64_t) (r >> 1) * (nthreads * 2) + (r & 1));
            }
        }
        __syncwarp();
    }

    // the key / predicate registers are dead: refill them with the tile after next
    if (refill_tile >= 0) load_tile_keys<PK, NV, MODE>(p, refill_tile, tid, nthreads, t);

    // ---- phase 2 (VAR 2): split entries -- SUM in an 8-byte array, {tag:8 | COUNT:24} in a 4-byte array ----
    // The 16-byte entry costs 7.3 + 10.4 + 10.4 shared-memory wavefronts per 32 rows (tag STS.32: the tag
    // words of 16-byte entries fall into only 8 of the 32 banks; LDS.128 / STS.128: four quarter-warp phases
    // of 8 random bank groups each) -- measured 31.5 with retries, 61 % of them bank-conflict replays, and
    // the pipe is 86 % busy at C3.  Split: tag byte store 3.5 + LDS.32 3.5 + LDS.64 6.2 + STS.64 6.2 +
    // STS.32 3.5 = 22.9.  Same arbitration protocol, same memory footprint (the region is carved in two).
    if constexpr (VAR == 2) {
        static_assert(SUMF64 && NW == 2, "the split layout is the COUNT + SUM(float64) entry");
        const uint32_t ct_off = (uint32_t) p.gmax * 8u;   // count/tag array behind the SUM array
#pragma unroll
        for (int q = 0; q < FA_R / FA_K; ++q) {
            uint32_t pend = (todo >> (q * FA_K)) & ((1u << FA_K) - 1u);
            uint32_t a_sum[FA_K], a_ct[FA_K];
#pragma unroll
            for (int k = 0; k < FA_K; ++k) {
                const uint32_t g = (ea[q * FA_K + k] - cx.a_ent) >> 4;
                a_sum[k] = cx.a_ent + g * 8u;
                a_ct[k] = cx.a_ent + ct_off + g * 4u;
            }
            if constexpr (HOT) {
#pragma unroll
                for (int k = 0; k < FA_K; ++k) fast_hot_group<MODE, PK, NV, 2>(p, cx, t, ea[q * FA_K + k], q * FA_K + k, pend, k);
            }
            while (__any_sync(0xffffffffu, pend != 0)) {
#pragma unroll
                for (int k = 0; k < FA_K; ++k)
                    if ((pend >> k) & 1u) sts8(a_ct[k] + 3, cx.lane | ((uint32_t) k << 5));
                __syncwarp();
                uint32_t ct[FA_K];
                uint64_t sm[FA_K];
#pragma unroll
                for (int k = 0; k < FA_K; ++k) {
                    ct[k] = 0xFFFFFFFFu;
                    sm[k] = 0;
                    if ((pend >> k) & 1u) {
                        ct[k] = lds32(a_ct[k]);
                        sm[k] = lds64(a_sum[k]);
                    }
                }
#pragma unroll
                for (int k = 0; k < FA_K; ++k) {
                    const uint32_t mine = cx.lane | ((uint32_t) k << 5);
                    if ((ct[k] >> 24) == mine) {   // 0xFF for rows that are not pending: never a tag
                        const int r = q * FA_K + k;
                        pend &= ~(1u << k);
                        const double sum = __longlong_as_double((long long) sm[k]) +
                                           __longlong_as_double((long long) row_val<MODE>(p, 0, t.vq[0], r));
                        sts64(a_sum[k], (uint64_t) __double_as_longlong(sum));
                        sts32(a_ct[k], ((ct[k] + 1u) & 0x00FFFFFFu) | (mine << 24));
                    }
                }
                __syncwarp();
            }
            if (refill_tile >= 0) {
#pragma unroll
                for (int j = q * FA_K / 2; j < (q + 1) * FA_K / 2; ++j) load_tile_vals<PK, NV, MODE>(p, refill_tile, tid, nthreads, t, j);
            }
        }
        return;
    }

    // ---- phase 2: accumulate into the warp-private entries, FA_K rows per round ----
    // The FA_K rows of a group arbitrate and update in the same round, so their
    // shared-memory round trips overlap (the kernel is bound by this dependent chain, not by
    // issue slots or by the shared-memory pipe: profiles/).  The tag is (lane, row slot), so
    // two rows of ONE lane that hit the same entry arbitrate like rows of different lanes.
    auto apply_row = [&](uint32_t (&e)[4], uint32_t (&f)[4], int r) {
        e[0] += 1;
        if constexpr (SUMF64) {
            const double sum = __hiloint2double((int) e[3], (int) e[2]) +
                               __longlong_as_double((long long) row_val<MODE>(p, 0, t.vq[0], r));
            e[2] = (uint32_t) __double2loint(sum);
            e[3] = (uint32_t) __double2hiint(sum);
        } else {
            uint64_t w[3] = {u64_of(e[2], e[3]), u64_of(f[0], f[1]), u64_of(f[2], f[3])};
#pragma unroll
            for (int c = 0; c < NCMAX; ++c) {
                if (c < p.n_cells) {
                    const FastCell cell = p.cell[c];
                    if (cell.op != CELL_I128_HI) {
                        uint64_t v = 0;
#pragma unroll
                        for (int k = 0; k < NV; ++k)
                            if (cell.col == k) v = row_val<MODE>(p, k, t.vq[k], r);
                        const uint64_t old = w[c];
                        const uint64_t nv = cell_apply(cell, old, v);
                        w[c] = nv;
                        if (NW == 4 && cell.op == CELL_ADD_I128 && c + 1 < NCMAX) {
                            const uint64_t ext = (!cell.in_unsigned && (int64_t) v < 0) ? ~0ULL : 0ULL;
                            w[c + 1 < 3 ? c + 1 : 2] += ext + (nv < old ? 1ULL : 0ULL);
                        }
                    }
                }
            }
            e[2] = (uint32_t) w[0];
            e[3] = (uint32_t) (w[0] >> 32);
            if constexpr (NW == 4) {
                f[0] = (uint32_t) w[1]; f[1] = (uint32_t) (w[1] >> 32);
                f[2] = (uint32_t) w[2]; f[3] = (uint32_t) (w[2] >> 32);
            }
        }
    };
#pragma unroll
    for (int q = 0; q < FA_R / FA_K; ++q) {
        uint32_t pend = (todo >> (q * FA_K)) & ((1u << FA_K) - 1u);
        if constexpr (SUMF64 && HOT) {
#pragma unroll
            for (int k = 0; k < FA_K; ++k) fast_hot_group<MODE, PK, NV, 0>(p, cx, t, ea[q * FA_K + k], q * FA_K + k, pend, k);
        }
        while (__any_sync(0xffffffffu, pend != 0)) {
#pragma unroll
            for (int k = 0; k < FA_K; ++k)
                if ((pend >> k) & 1u) sts32(ea[q * FA_K + k] + 4, cx.lane | ((uint32_t) k << 5));
            __syncwarp();
            uint32_t e[FA_K][4];
#pragma unroll
            for (int k = 0; k < FA_K; ++k) {
                e[k][1] = 0xFFFFFFFFu;
                if ((pend >> k) & 1u) lds128(ea[q * FA_K + k], e[k]);
            }
#pragma unroll
            for (int k = 0; k < FA_K; ++k) {
                if (e[k][1] == (cx.lane | ((uint32_t) k << 5))) {
                    const int r = q * FA_K + k;
                    pend &= ~(1u << k);
                    uint32_t f[4] = {0, 0, 0, 0};
                    if constexpr (NW == 4) lds128(ea[r] + 16, f);
                    apply_row(e[k], f, r);
                    sts128(ea[r], e[k]);
                    if constexpr (NW == 4) sts128(ea[r] + 16, f);
                }
            }
            __syncwarp();  // (measured: dropping this barrier buys nothing)
        }
        // the rows of this group are done: their value registers are dead, refill them
        if (refill_tile >= 0) {
#pragma unroll
            for (int j = q * FA_K / 2; j < (q + 1) * FA_K / 2; ++j) load_tile_vals<PK, NV, MODE>(p, refill_tile, tid, nthreads, t, j);
        }
    }
}

template <int PK, int NV, int NW, int MODE, bool DIRECT, bool SUMF64, int VAR = 0, bool HOT = false>
__global__ void __launch_bounds__(FA_MAX_THREADS, 1) agg_fast_kernel(const __grid_constant__ FastParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_ngroups;
    const int S = DIRECT ? 0 : (1 << p.log2s);
    const int G = p.gmax;
    const int nthreads = blockDim.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    constexpr int NCMAX = NW - 1;  // cells an entry can hold

    const uint32_t a_base = smem_addr(smem);
    const uint32_t a_warp0 = a_base + (uint32_t) fast_table_bytes(p.log2s, DIRECT);
    const uint32_t warp_bytes = (uint32_t) fast_warp_bytes(G, NW);
    FastCtx cx;
    cx.a_keys = a_base;
    cx.a_gid = a_base + (uint32_t) S * 8;
    cx.a_ent = a_warp0 + (uint32_t) warp * warp_bytes;
    cx.smask = (uint32_t) S - 1u;
    cx.hshift = 32 - p.log2s;
    cx.op = p.pred.op;
    cx.pscalar = p.pred.scalar.bits;
    cx.lane = (uint32_t) lane;

    {
        if constexpr (!DIRECT) {
            uint64_t* k = reinterpret_cast<uint64_t*>(smem);
            uint16_t* g = reinterpret_cast<uint16_t*>(smem + (size_t) S * 8);
            if (p.dict_keys != nullptr) {
                for (int i = tid; i < S; i += nthreads) {
                    k[i] = p.dict_keys[i];
                    g[i] = p.dict_gids[i];
                }
            } else {
                for (int i = tid; i < S; i += nthreads) {
                    k[i] = LK_EMPTY;
                    g[i] = (uint16_t) GID_PENDING;
                }
            }
        }
        uint32_t* z = reinterpret_cast<uint32_t*>(smem + fast_table_bytes(p.log2s, DIRECT));
        const size_t words = (size_t) warp_bytes * nwarps / 4;
        for (size_t i = tid; i < words; i += nthreads) z[i] = 0;
        if (tid == 0) s_ngroups = 0;
    }
    __syncthreads();

    uint32_t spilled = 0;
    const int64_t tile_rows = (int64_t) nthreads * FA_R;

    // two register tiles, explicitly alternated, each refilled piecewise while it is processed
    RawTile<PK, NV> ta, tb;
    int64_t tile = blockIdx.x;
    const int64_t stride = gridDim.x;
    auto load_all = [&](int64_t tl, RawTile<PK, NV>& t) {
        load_tile_keys<PK, NV, MODE>(p, tl, tid, nthreads, t);
#pragma unroll
        for (int j = 0; j < FA_R / 2; ++j) load_tile_vals<PK, NV, MODE>(p, tl, tid, nthreads, t, j);
    };
    auto process = [&](int64_t tl, RawTile<PK, NV>& t) {
        const int64_t nx = tl + 2 * stride;
        fast_tile<PK, NV, NW, MODE, DIRECT, SUMF64, VAR, HOT>(p, cx, t, smem, &s_ngroups, tl * tile_rows + tid * 2, tid, nthreads,
                                                    nx < p.num_tiles ? nx : (int64_t) -1, spilled);
    };
    if (tile < p.num_tiles) load_all(tile, ta);
    if (tile + stride < p.num_tiles) load_all(tile + stride, tb);
    // one lane per column streams the tile `pf_dist` rounds ahead into L2
    auto prefetch = [&](int64_t tl) {
        if (p.pf_dist <= 0 || lane != 0 || warp > NV + 1) return;
        const int64_t pt = tl + (int64_t) p.pf_dist * stride;
        if (pt >= p.num_tiles) return;
        const int64_t r = pt * tile_rows;
        if (warp == 0) {
            const bool k8 = MODE == FM_ALL8 || (MODE == FM_RUNTIME && p.key_mode == 0);
            l2_prefetch_bulk(p.key.data + r * (k8 ? 8 : 4), (uint32_t) tile_rows * (k8 ? 8u : 4u));
        } else if (warp == 1) {
            if constexpr (PK == PK_F64_VEC || PK == PK_I64_VEC) l2_prefetch_bulk(p.pred.col.data + r * 8, (uint32_t) tile_rows * 8u);
            else if constexpr (PK == PK_MASK) l2_prefetch_bulk(p.pred.mask + r, (uint32_t) tile_rows);
        } else {
            const int v = warp - 2;
            if (v < NV) {
                const bool v8 = MODE != FM_RUNTIME || p.col_mode[v] == 0;
                l2_prefetch_bulk(p.col[v].data + r * (v8 ? 8 : 4), (uint32_t) tile_rows * (v8 ? 8u : 4u));
            }
        }
    };
    while (tile < p.num_tiles) {
        prefetch(tile);
        process(tile, ta);
        tile += stride;
        if (tile >= p.num_tiles) break;
        prefetch(tile);
        process(tile, tb);
        tile += stride;
    }

    // ---- flush: fold the warps' private entries, one global update per group ----
    __syncthreads();
    for (int d = 16; d > 0; d >>= 1) spilled += __shfl_xor_sync(0xffffffffu, spilled, d);
    if (lane == 0 && spilled) atomicAdd(p.replay.spilled, (unsigned long long) spilled);
    const int n_iter = DIRECT ? G : S;
    unsigned long long selected = 0;
    long long kmin = INT64_MAX, kmax = INT64_MIN;
    for (int s = tid; s < n_iter; s += nthreads) {
        uint64_t key;
        uint32_t g16;
        if constexpr (DIRECT) {
            key = p.direct_base + (uint64_t) s;
            g16 = (uint32_t) s;
        } else {
            key = lds64(cx.a_keys + s * 8);
            if (key == LK_EMPTY) continue;
            g16 = lds16(cx.a_gid + s * 2);
            if (g16 >= (uint32_t) G) continue;
        }
        uint64_t cnt = 0;
        uint64_t acc[FA_MAX_CELLS] = {0, 0, 0};
        double facc[FA_MAX_CELLS] = {0.0, 0.0, 0.0};
        for (int w = 0; w < nwarps; ++w) {
            if constexpr (VAR == 2) {   // split entries: SUM array, then {tag | COUNT} array
                const uint32_t wb = a_warp0 + (uint32_t) w * warp_bytes;
                cnt += lds32(wb + (uint32_t) G * 8u + g16 * 4u) & 0x00FFFFFFu;
                facc[0] += __longlong_as_double((long long) lds64(wb + g16 * 8u));
                continue;
            }
            const uint32_t ea = a_warp0 + (uint32_t) w * warp_bytes + g16 * (NW * 8);
            uint32_t e[4], f[4] = {0, 0, 0, 0};
            lds128(ea, e);
            if constexpr (NW == 4) lds128(ea + 16, f);
            cnt += e[0];
            const uint64_t ev[3] = {u64_of(e[2], e[3]), u64_of(f[0], f[1]), u64_of(f[2], f[3])};
#pragma unroll
            for (int c = 0; c < NCMAX; ++c) {
                if (c < p.n_cells) {
                    const int op = p.cell[c].op;
                    const uint64_t v = ev[c];
                    if (op == CELL_ADD_F64) facc[c] += __longlong_as_double((long long) v);
                    else if (op == CELL_MAXORD) acc[c] = v > acc[c] ? v : acc[c];
                    else if (op == CELL_ADD_I128) {
                        const uint64_t nl = acc[c] + v;
                        if (c + 1 < NCMAX) acc[c + 1] += (nl < acc[c] ? 1ULL : 0ULL);
                        acc[c] = nl;
                    } else acc[c] += v;  // ADD_I64, I128_HI
                }
            }
        }
        if (cnt == 0) continue;
        selected += cnt;
        kmin = (long long) key < kmin ? (long long) key : kmin;
        kmax = (long long) key > kmax ? (long long) key : kmax;
#pragma unroll
        for (int c = 0; c < NCMAX; ++c)
            if (c < p.n_cells && p.cell[c].op == CELL_ADD_F64) acc[c] = (uint64_t) __double_as_longlong(facc[c]);
        // the host reserves capacity for every CTA's groups: this insert cannot fail
        const int64_t g = gt1_find_or_insert(p.table, key, false, hash_key1(key), INT64_MAX);
        if (g < 0) {
            atomicAdd(p.replay.lost, (unsigned long long) cnt);
            continue;
        }
        fast_global_update(p, g, cnt, acc);
    }
    // selectivity statistic for the host's geometry choice (one atomic per warp)
    for (int d = 16; d > 0; d >>= 1) selected += __shfl_xor_sync(0xffffffffu, selected, d);
    if (lane == 0 && selected) atomicAdd(p.replay.selected, selected);
    if (p.key_range != nullptr) {  // learning launch: the key range decides about direct group ids
        for (int d = 16; d > 0; d >>= 1) {
            const long long a = __shfl_xor_sync(0xffffffffu, kmin, d), b = __shfl_xor_sync(0xffffffffu, kmax, d);
            kmin = a < kmin ? a : kmin;
            kmax = b > kmax ? b : kmax;
        }
        if (lane == 0 && kmin <= kmax) {
            atomicMin(p.key_range, kmin);
            atomicMax(p.key_range + 1, kmax);
        }
    }
}

// ---- launch plumbing (instantiated per predicate kind in vk_agg_fast_pk*.cu) -------------
struct FastLaunch {
    int pk;        // PredKernelKind
    int nw;        // 2 or 4
    int mode;      // FastMode
    bool direct;
    bool sumf64;
    int variant;   // SUMF64 only: 0 = 16-byte entries, 2 = split (SoA) entries; both with tag arbitration
    bool hot;      // SUMF64 only: the kernel with the hot-group step (a key that holds a large share of the rows)
    int grid, threads;
    size_t smem;
};


#define VK_FAST_GO(PK, NV, NW, MODE, DIRECT, SUMF64, ...)                                                        \
    do {                                                                                                       \
        auto kernel = agg_fast_kernel<PK, NV, NW, MODE, DIRECT, SUMF64, ##__VA_ARGS__>;                         \
        VK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) l.smem));      \
        kernel<<<l.grid, l.threads, l.smem, s>>>(p);                                                           \
        VK_CHECK_LAUNCH("agg_fast_kernel");                                                                    \
        return VK_OK;                                                                                          \
    } while (0)

// Lean set: compile-time layout, optional direct mode (predicate kinds NONE / F64_VEC).
template <int PK, int MODE, bool DIRECT>
int launch_fast_lean(const FastParams& p, const FastLaunch& l, cudaStream_t s) {
    switch (p.n_cols) {
        case 0: VK_FAST_GO(PK, 0, 2, MODE, DIRECT, false);
        case 1:
            if (l.nw == 2 && l.sumf64 && l.hot && l.variant == 2) VK_FAST_GO(PK, 1, 2, MODE, DIRECT, true, 2, true);
            if (l.nw == 2 && l.sumf64 && l.hot) VK_FAST_GO(PK, 1, 2, MODE, DIRECT, true, 0, true);
            if (l.nw == 2 && l.sumf64 && l.variant == 2) VK_FAST_GO(PK, 1, 2, MODE, DIRECT, true, 2);
            if (l.nw == 2 && l.sumf64) VK_FAST_GO(PK, 1, 2, MODE, DIRECT, true);
            if (l.nw == 2) VK_FAST_GO(PK, 1, 2, MODE, DIRECT, false);
            VK_FAST_GO(PK, 1, 4, MODE, DIRECT, false);
        case 2: VK_FAST_GO(PK, 2, 4, MODE, DIRECT, false);
        default: VK_FAST_GO(PK, 3, 4, MODE, DIRECT, false);
    }
}
// Generic set: run-time column modes, hash mode only (every predicate kind).
template <int PK>
int launch_fast_generic(const FastParams& p, const FastLaunch& l, cudaStream_t s) {
    switch (p.n_cols) {
        case 0: VK_FAST_GO(PK, 0, 2, FM_RUNTIME, false, false);
        case 1:
            if (l.nw == 2) VK_FAST_GO(PK, 1, 2, FM_RUNTIME, false, false);
            VK_FAST_GO(PK, 1, 4, FM_RUNTIME, false, false);
        case 2: VK_FAST_GO(PK, 2, 4, FM_RUNTIME, false, false);
        default: VK_FAST_GO(PK, 3, 4, FM_RUNTIME, false, false);
    }
}
// The instantiations are spread over several translation units (vk_agg_fast_inst.cu is
// compiled once per VK_FAST_PART) so that they build in parallel.
#define VK_FAST_DECL(name) int name(const FastParams& p, const FastLaunch& l, cudaStream_t s)
VK_FAST_DECL(launch_fast_none_all8_d);   // _d: direct group ids, _t: CTA key table / dictionary
VK_FAST_DECL(launch_fast_none_all8_t);
VK_FAST_DECL(launch_fast_none_key4_d);
VK_FAST_DECL(launch_fast_none_key4_t);
VK_FAST_DECL(launch_fast_none_rt);
VK_FAST_DECL(launch_fast_f64_all8_d);
VK_FAST_DECL(launch_fast_f64_all8_t);
VK_FAST_DECL(launch_fast_f64_key4_d);
VK_FAST_DECL(launch_fast_f64_key4_t);
VK_FAST_DECL(launch_fast_f64_rt);
VK_FAST_DECL(launch_fast_mask_rt);
VK_FAST_DECL(launch_fast_i64_rt);
// PK_GENERIC carries the fused expression evaluator: one kernel per translation unit (each takes ~1 min of ptxas)
VK_FAST_DECL(launch_fast_gen_rt_c0);
VK_FAST_DECL(launch_fast_gen_rt_c1n);   // one value column, 2 warps per table
VK_FAST_DECL(launch_fast_gen_rt_c1);
VK_FAST_DECL(launch_fast_gen_rt_c2);
VK_FAST_DECL(launch_fast_gen_rt_c3);
inline int launch_fast_gen_rt(const FastParams& p, const FastLaunch& l, cudaStream_t s) {
    switch (p.n_cols) {
        case 0: return launch_fast_gen_rt_c0(p, l, s);
        case 1: return l.nw == 2 ? launch_fast_gen_rt_c1n(p, l, s) : launch_fast_gen_rt_c1(p, l, s);
        case 2: return launch_fast_gen_rt_c2(p, l, s);
        default: return launch_fast_gen_rt_c3(p, l, s);
    }
}

// Which predicate kinds have the lean set (the host only asks for direct / compile-time modes there).
__host__ inline bool fast_pk_is_lean(int pk) { return pk == PK_NONE || pk == PK_F64_VEC; }

}  // namespace vk
