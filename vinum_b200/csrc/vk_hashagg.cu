// Hash group-by aggregate for sm_100a.
//
// Replaces the row-at-a-time loop of BaseAggregate::Next
// (vinum_cpp/src/operators/aggregate/base_aggregate.cpp:23-45) and its
// SingleNumerical / MultiNumerical / OneGroup specialisations with these kernels:
//
//   agg_fast_kernel   low-cardinality single-key path (the north-star pipeline):
//                     persistent CTAs stream row tiles with 16-byte loads, evaluate the
//                     WHERE predicate in registers, resolve the key in a shared-memory
//                     open-addressing table and accumulate COUNT / SUM(f64) there; one
//                     flush per CTA into the global table.  No row is written back.
//   agg_general_kernel any key arity / dtype / NULLs / function: straight to the global
//                     (L2/HBM) table with native 64-bit atomics, one row per thread.
//   agg_wide_kernel   the same update for single-key tables (more groups than shared memory
//                     holds): four rows per thread, every stage's loads issued together, linear
//                     probing in lock step across the warp.
//   agg_part_scatter_kernel + agg_part_update_kernel
//                     tables beyond the L2: rows are scattered into buckets that are slices of
//                     the table, then the table is updated slice by slice (wide_update over the
//                     records) so that the slice stays in the L2.
//   agg_onegroup_kernel un-grouped reduction (OneGroupAggregate::Next,
//                     one_group_aggregate.cpp:9-26): register accumulate, warp shuffle,
//                     one atomic per CTA.
//
// Rows that cannot be inserted because the global table is at its load limit are
// appended to a replay list; the host grows the table and replays them, so an update
// never loses rows whatever the cardinality turns out to be.
#include "vk_hashagg.cuh"
#include "vk_agg_fast.cuh"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace vk {

// ============================================================ key range (direct mode)
// Smallest / largest key of a single-key table in SIGNED 64-bit order; out[0] = min, out[1] = max.
__global__ void __launch_bounds__(256) agg_key_range_kernel(GTable t, long long* out) {
    long long mn = INT64_MAX, mx = INT64_MIN;
    const int64_t slots = t.capacity + 1;  // + the slot of the all-ones key; the NULL group has no key
    for (int64_t s = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (int64_t) gridDim.x * blockDim.x) {
        if (!gt_slot_occupied(t, s)) continue;
        const long long k = s < t.capacity ? (long long) t.keys[s] : -1LL;
        mn = k < mn ? k : mn;
        mx = k > mx ? k : mx;
    }
    for (int d = 16; d > 0; d >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, mn, d), b = __shfl_xor_sync(0xffffffffu, mx, d);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0 && mn <= mx) {
        atomicMin(out, mn);
        atomicMax(out + 1, mx);
    }
}

// Rows of the largest group of a table (the learning launch's measure of key skew).
__global__ void __launch_bounds__(256) agg_max_count_kernel(GTable t, unsigned long long* out) {
    unsigned long long mx = 0;
    const int64_t slots = gt_total_slots(t);
    for (int64_t s = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (int64_t) gridDim.x * blockDim.x)
        if (gt_slot_occupied(t, s) && t.count_star[s] > mx) mx = t.count_star[s];
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, mx, d);
        mx = o > mx ? o : mx;
    }
    if ((threadIdx.x & 31) == 0 && mx) atomicMax(out, mx);
}

// Keys of a single-key table as a dense list (the dictionary builder reads them back).
__global__ void __launch_bounds__(256) agg_export_keys_kernel(GTable t, uint64_t* out, unsigned long long* cursor,
                                                              unsigned long long cap) {
    const int64_t slots = t.capacity + 1;  // + the slot of the all-ones key; the NULL group has no key
    for (int64_t s = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += (int64_t) gridDim.x * blockDim.x) {
        if (!gt_slot_occupied(t, s)) continue;
        const unsigned long long pos = atomicAdd(cursor, 1ULL);
        if (pos < cap) out[pos] = s < t.capacity ? t.keys[s] : ~0ULL;
    }
}

// ============================================================ general kernel
struct GenParams {
    Pred pred;
    int pk;                          // PredKernelKind (evaluated row-wise)
    int n_keys;
    Col keys[VK_AGG_MAX_KEYS];
    int n_funcs;
    FuncSpec specs[VK_AGG_MAX_FUNCS];
    Col vals[VK_AGG_MAX_FUNCS];
    int64_t n;
    int64_t row_begin;               // first row to process (the fast path leaves its ragged tail here)
    const uint32_t* row_list;        // replay: process these rows only
    const unsigned long long* row_list_count;
    GTable table;
    ReplayList replay;
};

__device__ __forceinline__ bool pred_row(const Pred& p, int64_t i) {
    if (p.kind == VK_PRED_NONE) return true;
    if (p.kind == VK_PRED_MASK) return p.mask[i] != 0;
    return pred_row_generic(p, i);
}

// Rows that passed the predicate (first visits only), one atomic per warp: the host sizes the table from
// (selected rows, groups) while the number of groups is still rising.
__device__ __forceinline__ void count_selected(const GenParams& p, unsigned n_selected) {
    if (p.row_list != nullptr || p.replay.selected == nullptr) return;
    __syncwarp();
    n_selected = __reduce_add_sync(0xffffffffu, n_selected);
    if ((threadIdx.x & 31) == 0 && n_selected) atomicAdd(p.replay.selected, (unsigned long long) n_selected);
}

// One row (that passed the predicate) into the global table: any key arity / dtype / NULLs / function.
__device__ __forceinline__ void general_update_row(const GenParams& p, int64_t i) {
    uint64_t kv[VK_AGG_MAX_KEYS];
    uint32_t nullmask = 0;
    for (int k = 0; k < p.n_keys; ++k) {
        bool valid = col_valid(p.keys[k], i);
        kv[k] = valid ? load_as_u64(p.keys[k], i) : 0;
        nullmask |= (uint32_t) (!valid) << k;
    }
    int64_t slot = gt_find_or_insert<0>(p.table, kv, nullmask, hash_keys(kv, nullmask, p.n_keys), p.table.max_groups);
    if (slot < 0) {
        replay_append(p.replay, i);
        return;
    }
    atomicAdd(reinterpret_cast<unsigned long long*>(p.table.count_star + slot), 1ULL);
    for (int f = 0; f < p.n_funcs; ++f) {
        const FuncSpec spec = p.specs[f];
        if (spec.acc == ACC_NONE) continue;
        if (!col_valid(p.vals[f], i)) {
            atomicAdd(reinterpret_cast<unsigned long long*>(p.table.nnull[f] + slot), 1ULL);
            continue;
        }
        if (spec.acc == ACC_COUNT) continue;
        acc_update_global(p.table, f, spec, slot, acc_load(spec, p.vals[f], i));
    }
}
static __device__ __noinline__ void general_update_row_cold(const GenParams& p, int64_t i) { general_update_row(p, i); }

__global__ void __launch_bounds__(256) agg_general_kernel(const __grid_constant__ GenParams p) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const bool replaying = p.row_list != nullptr;
    const int64_t total = replaying ? (int64_t) *p.row_list_count : p.n;
    unsigned n_selected = 0;
    for (int64_t it = (replaying ? 0 : p.row_begin) + (int64_t) blockIdx.x * blockDim.x + threadIdx.x; it < total; it += stride) {
        const int64_t i = replaying ? (int64_t) p.row_list[it] : it;
        if (!replaying && !pred_row(p.pred, i)) continue;  // listed rows already passed the predicate
        ++n_selected;
        general_update_row(p, i);
    }
    count_selected(p, n_selected);
}

// L2 residency control for the partitioned update: the slice of the table a bucket updates must stay in the L2 while
// records and gathered values stream through it.  Without hints every gathered 32-byte sector claims a 128-byte line,
// the L2 turns over in a few microseconds and the table slice is evicted between two touches (measured: RED hit rate
// 24 %, 227 DRAM bytes per record).  Table loads / reductions carry an evict_last policy, streams an evict_first one.
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ uint64_t ld_key_relaxed_hint(const uint64_t* p, uint64_t pol) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}
__device__ __forceinline__ void red_add_u64_hint(uint64_t* p, uint64_t v, uint64_t pol) {
    asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.u64 [%0], %1, %2;" ::"l"(p), "l"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void red_add_f64_hint(uint64_t* p, double v, uint64_t pol) {
    asm volatile("red.relaxed.gpu.global.add.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint64_t ldg_u64_stream_hint(const void* p, uint64_t pol) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ ulonglong2 ldg_rec_stream_hint(const ulonglong2* p, uint64_t pol) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0, %1}, [%2], %3;" : "=l"(v.x), "=l"(v.y) : "l"(p), "l"(pol));
    return v;
}

// Stage 1 of a tile of the global-table kernels: which of the thread's R rows (R/2 adjacent pairs, 512 rows apart) pass
// the predicate, and their keys.  `plain` = an ordinary key: not NULL (the value that marks a free slot is sorted out later).
template <int PK, int R>
__device__ __forceinline__ void wide_select_and_keys(const GenParams& p, int64_t base, bool whole, bool (&sel)[R], uint64_t (&key)[R],
                                                     bool (&plain)[R]) {
    const Col& kc = p.keys[0];
    if ((PK == PK_F64_VEC || PK == PK_I64_VEC) && whole) {
        // whole tile: the R/2 predicate loads first, the comparisons after them
        uint4 q[R / 2];
#pragma unroll
        for (int j = 0; j < R / 2; ++j) q[j] = ldg_stream16(p.pred.col.data + (base + j * 512) * 8);
#pragma unroll
        for (int j = 0; j < R / 2; ++j) {
            if (PK == PK_F64_VEC) {
                const double c = __longlong_as_double((long long) p.pred.scalar.bits);
                sel[2 * j] = apply_cmp(p.pred.op, __hiloint2double(q[j].y, q[j].x), c);
                sel[2 * j + 1] = apply_cmp(p.pred.op, __hiloint2double(q[j].w, q[j].z), c);
            } else {
                const int64_t c = (int64_t) p.pred.scalar.bits;
                sel[2 * j] = apply_cmp(p.pred.op, (int64_t) (((uint64_t) q[j].y << 32) | q[j].x), c);
                sel[2 * j + 1] = apply_cmp(p.pred.op, (int64_t) (((uint64_t) q[j].w << 32) | q[j].z), c);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < R; j += 2) pred_pair<PK>(p.pred, base + (j / 2) * 512, p.n, sel[j], sel[j + 1]);
    }
    if (kc.validity == nullptr && (kc.dtype == VK_I64 || kc.dtype == VK_U64 || kc.dtype == VK_F64)) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
            key[j] = 0;
            plain[j] = sel[j];
            if (sel[j]) key[j] = reinterpret_cast<const uint64_t*>(kc.data)[base + (j / 2) * 512 + (j & 1)];
        }
    } else if (kc.validity == nullptr && kc.dtype == VK_I32) {
#pragma unroll
        for (int j = 0; j < R; ++j) {
            key[j] = 0;
            plain[j] = sel[j];
            if (sel[j]) key[j] = (uint64_t) (int64_t) reinterpret_cast<const int32_t*>(kc.data)[base + (j / 2) * 512 + (j & 1)];
        }
    } else {
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int64_t i = base + (j / 2) * 512 + (j & 1);
            key[j] = 0;
            plain[j] = false;
            if (sel[j]) {
                plain[j] = col_valid(kc, i);
                if (plain[j]) key[j] = load_as_u64(kc, i);
            }
        }
    }
}

// Stages 2 - 4: the slot of every selected row (lock-step probing), then the accumulator updates.  `row` = row ids
// (for the value columns and the replay list).  Must be called by all 32 lanes of a warp together.
template <int R, bool HINTS = false, bool HOT = false>
__device__ __forceinline__ void wide_update(const GenParams& p, bool (&sel)[R], const uint64_t (&key)[R], const bool (&plain)[R],
                                            const uint32_t (&row)[R]) {   // row ids fit 32 bits (chunks are cut accordingly)
    const GTable& t = p.table;
    const Col& kc = p.keys[0];
    (void) kc;
    const uint64_t mask = (uint64_t) t.capacity - 1;
    const uint64_t keep = HINTS ? l2_policy_keep() : 0, stream = HINTS ? l2_policy_stream() : 0;
    int64_t slot[R];
    bool pend[R];   // rows still looking for their slot
#pragma unroll
    for (int j = 0; j < R; ++j) {
        pend[j] = plain[j] && key[j] != GT_EMPTY;
        slot[j] = (int64_t) (hash_key1(key[j]) & mask);
        if (sel[j] && !pend[j]) {
            // NULL key / the key value that marks a free slot: their two dedicated slots
            slot[j] = gt1_find_or_insert(t, key[j], !plain[j], 0, t.max_groups);
            if (slot[j] < 0) {
                replay_append(p.replay, (int64_t) row[j]);
                sel[j] = false;
            }
        }
    }
    // Linear probing in lock step: one round loads the current slot of every pending row of the warp
    // (R loads in flight per lane); each row is then found, inserted by CAS (the protocol of
    // gt1_find_or_insert) or moves one slot on.  Votes keep the warp converged; the group counter, one
    // hot address, is read once per warp and round and advanced once per warp.  (Loading two slots per
    // round was measured: fewer rounds, but 64 registers instead of 56 and 10 % slower at 1e4 - 1e6 groups.)
    const unsigned lane = threadIdx.x & 31u;
    for (int64_t probes = 0; probes < t.capacity; ++probes) {
        uint64_t seen[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            seen[j] = key[j];
            if (pend[j]) seen[j] = HINTS ? ld_key_relaxed_hint(t.keys + slot[j], keep) : ld_key_relaxed(t.keys + slot[j]);
        }
        // rows at a free slot claim it by CAS; the group counter (one hot address) is read once per warp and round and
        // advanced once, whatever the number of claims (an insert-heavy launch spent most of its time on that line)
        bool claims = false;
        unsigned settled = 0;   // bit j: row slot j was dealt with in the claim block (found, inserted, deferred or moved on)
#pragma unroll
        for (int j = 0; j < R; ++j) claims = claims || (pend[j] && seen[j] == GT_EMPTY);
        if (__any_sync(0xffffffffu, claims)) {   // warp-uniform; nothing of this block is live on the common path
            unsigned long long groups_now = 0;
            if (lane == 0) groups_now = *reinterpret_cast<volatile unsigned long long*>(t.num_groups);
            groups_now = __shfl_sync(0xffffffffu, groups_now, 0);
            unsigned before = 0, won_total = 0;   // claims of earlier row slots / inserts of this round
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const bool claim = pend[j] && seen[j] == GT_EMPTY;
                const unsigned claimers = __ballot_sync(0xffffffffu, claim);
                if (!claimers) continue;   // warp-uniform
                bool inserted = false;
                if (claim) {
                    settled |= 1u << j;
                    if (groups_now + before + __popc(claimers & ((1u << lane) - 1u)) >= (unsigned long long) t.max_groups) {
                        replay_append(p.replay, (int64_t) row[j]);
                        sel[j] = false;
                        pend[j] = false;
                    } else {
                        const unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(t.keys + slot[j]),
                                                                 (unsigned long long) GT_EMPTY, (unsigned long long) key[j]);
                        inserted = old == GT_EMPTY;
                        if (inserted || old == key[j]) pend[j] = false;
                        else slot[j] = (int64_t) (((uint64_t) slot[j] + 1) & mask);   // another key took it
                    }
                }
                won_total += __popc(__ballot_sync(0xffffffffu, inserted));
                before += __popc(claimers);
            }
            if (lane == 0 && won_total) atomicAdd(t.num_groups, (unsigned long long) won_total);
        }
        bool more = false;
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if (pend[j] && !((settled >> j) & 1u)) {
                if (seen[j] == key[j]) pend[j] = false;
                else slot[j] = (int64_t) (((uint64_t) slot[j] + 1) & mask);
            }
            more = more || pend[j];
        }
        if (!__any_sync(0xffffffffu, more)) break;
    }
#pragma unroll
    for (int j = 0; j < R; ++j)
        if (pend[j]) {   // every slot visited: cannot happen below the load limit, but no row may be lost
            replay_append(p.replay, (int64_t) row[j]);
            sel[j] = false;
        }
    __syncwarp();
    // A key that many lanes of the warp hold (skewed keys): same-address reductions serialise in one L2 slice at about one
    // lane per cycle for the whole GPU, so half the rows on one key would cost 0.2 s per 1e9 rows.  The lanes that share the
    // slot of the LOWEST selected lane are found with one shuffle and one vote per row slot; from HOT_MIN of them on, that
    // lane alone updates the slot for all of them (COUNT: the population count; SUM(float64): a warp butterfly).  A template
    // parameter: the host launches this variant when the learning launch saw one key hold >= 30 % of the rows (AGG_HOT).
    constexpr int HOT_MIN = 4;
    unsigned hot[R];
    int64_t hot_slot[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
        hot[j] = 0;
        hot_slot[j] = -1;
        if (HOT) {
            const unsigned live = __ballot_sync(0xffffffffu, sel[j]);
            if (live) {   // warp-uniform
                hot_slot[j] = __shfl_sync(0xffffffffu, slot[j], __ffs(live) - 1);
                const unsigned same = __ballot_sync(0xffffffffu, sel[j] && slot[j] == hot_slot[j]);
                if (__popc(same) >= HOT_MIN) hot[j] = same;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < R; ++j)
        if (sel[j] && !(HOT && hot[j] && slot[j] == hot_slot[j])) {
            if (HINTS) red_add_u64_hint(t.count_star + slot[j], 1ULL, keep);
            else atomicAdd(reinterpret_cast<unsigned long long*>(t.count_star + slot[j]), 1ULL);
        }
    if (HOT) {
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (hot[j] && lane == (unsigned) (__ffs(hot[j]) - 1))
                atomicAdd(reinterpret_cast<unsigned long long*>(t.count_star + slot[j]), (unsigned long long) __popc(hot[j]));
    }
    for (int f = 0; f < p.n_funcs; ++f) {
        const FuncSpec spec = p.specs[f];
        if (spec.acc == ACC_NONE) continue;
        const Col& vc = p.vals[f];
        uint64_t v[R];
        bool ok[R];
        if (spec.acc == ACC_SUM_F64 && vc.dtype == VK_F64 && vc.validity == nullptr) {
            // the common shape, free of per-row switches so that the R loads issue back to back
#pragma unroll
            for (int j = 0; j < R; ++j) {
                ok[j] = sel[j];
                v[j] = 0;
                if (sel[j])
                    v[j] = HINTS ? ldg_u64_stream_hint(reinterpret_cast<const uint64_t*>(vc.data) + row[j], stream)
                                 : reinterpret_cast<const uint64_t*>(vc.data)[row[j]];
            }
#pragma unroll
            for (int j = 0; j < R; ++j) {
                double x = __longlong_as_double((long long) v[j]);
                bool mine = ok[j];
                if (HOT && hot[j]) {   // warp-uniform
                    const bool member = ok[j] && slot[j] == hot_slot[j];
                    double part = member ? x : 0.0;
                    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
                    if (member) {
                        mine = lane == (unsigned) (__ffs(hot[j]) - 1);
                        x = part;
                    }
                }
                if (mine) {
                    if (HINTS) red_add_f64_hint(t.acc_lo[f] + slot[j], x, keep);
                    else atomicAdd(reinterpret_cast<double*>(t.acc_lo[f] + slot[j]), x);
                }
            }
            continue;
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int64_t i = (int64_t) row[j];
            ok[j] = false;
            v[j] = 0;
            if (!sel[j]) continue;
            if (!col_valid(vc, i)) {
                atomicAdd(reinterpret_cast<unsigned long long*>(t.nnull[f] + slot[j]), 1ULL);
                continue;
            }
            if (spec.acc == ACC_COUNT) continue;
            ok[j] = true;
            v[j] = acc_load(spec, vc, i);
        }
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (ok[j]) acc_update_global(t, f, spec, slot[j], v[j]);
    }
}

// ---- the same update for single-key tables, R rows per thread ---------------------------------
// agg_general_kernel keeps ONE row per thread in flight and walks a chain of dependent loads for it
// (predicate -> key -> table slot -> value): at 1e6 groups ncu shows 62 % long-scoreboard stalls at
// 94 % occupancy (profiles/r02_agg_general_1e6_ncu_full.md).  Here a thread owns R rows of a tile
// (as R/2 adjacent pairs, so that 8-byte predicate columns arrive by 16-byte loads) and every stage
// is issued for all R rows before the next stage waits on it: R predicate loads in flight, then R
// key loads, then R first-probe loads of the table (the wait-free single-key protocol of
// gt1_find_or_insert makes a plain load a complete look-up when the key is already there), then the
// value loads, then the fire-and-forget reductions.  Rows that meet a foreign slot probe on in lock step
// with the other rows of the warp; an empty slot is claimed by the same CAS as in gt1_find_or_insert.
template <int PK, int R, bool HOT>
__global__ void __launch_bounds__(256) agg_wide_kernel(const __grid_constant__ GenParams p) {
    constexpr int64_t TILE = 256 * R;
    const int64_t n_tiles = (p.n - p.row_begin + TILE - 1) / TILE;
    unsigned n_selected = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // the tile loop is warp-uniform; a warp that entered a tile diverged was measured at 4.3 active lanes
        // per issued instruction and 3.6 x the DRAM traffic
        __syncwarp();
        const int64_t base = p.row_begin + tile * TILE + 2 * (int64_t) threadIdx.x;   // row_begin is even
        bool sel[R], plain[R];
        uint64_t key[R];
        uint32_t row[R];
        wide_select_and_keys<PK, R>(p, base, p.row_begin + (tile + 1) * TILE <= p.n, sel, key, plain);
#pragma unroll
        for (int j = 0; j < R; ++j) {
            n_selected += sel[j] ? 1u : 0u;
            row[j] = (uint32_t) (base + (j / 2) * 512 + (j & 1));
        }
        wide_update<R, false, HOT>(p, sel, key, plain, row);
    }
    count_selected(p, n_selected);
}

// One launch: as many CTAs as are resident at once (the tile loop is persistent).
static int launch_wide(const GenParams& gp, bool hot, int sms, cudaStream_t s) {
    constexpr int R = 4;
    const int64_t tiles = (gp.n - gp.row_begin + 256 * R - 1) / (256 * R);
    auto go = [&](auto kernel) -> int {
        static int per_sm = 0;   // same for every device of this architecture
        if (per_sm == 0) {
            int b = 0;
            VK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, 256, 0));
            per_sm = b > 0 ? b : 1;
        }
        const int64_t cap = (int64_t) sms * per_sm;
        kernel<<<(unsigned) (tiles < cap ? tiles : cap), 256, 0, s>>>(gp);
        VK_CHECK_LAUNCH("agg_wide_kernel");
        return VK_OK;
    };
    switch (gp.pk) {
        case PK_NONE: return hot ? go(agg_wide_kernel<PK_NONE, R, true>) : go(agg_wide_kernel<PK_NONE, R, false>);
        case PK_MASK: return hot ? go(agg_wide_kernel<PK_MASK, R, true>) : go(agg_wide_kernel<PK_MASK, R, false>);
        case PK_F64_VEC: return hot ? go(agg_wide_kernel<PK_F64_VEC, R, true>) : go(agg_wide_kernel<PK_F64_VEC, R, false>);
        case PK_I64_VEC: return hot ? go(agg_wide_kernel<PK_I64_VEC, R, true>) : go(agg_wide_kernel<PK_I64_VEC, R, false>);
        default: return hot ? go(agg_wide_kernel<PK_GENERIC, R, true>) : go(agg_wide_kernel<PK_GENERIC, R, false>);
    }
}

// ============================================================ partitioned aggregate (tables beyond the L2)
// Once the global table no longer fits the L2 (1.5e6 groups and more) every row costs three random DRAM sectors:
// 72 ms at 8.4e6 groups for 1e9 rows.  Two passes that keep the table traffic inside the L2 cost less:
//   scatter  every selected row's (key, row id) is appended to one of P buckets, the bucket being the RANGE OF TABLE
//            SLOTS the key hashes into (top bits of hash & mask): one atomic on the bucket's cursor and ONE 16-byte store
//            per row (the pass is bound by uncoalesced transactions, ~60 G/s, not by bytes; the L2 merges neighbouring
//            appends into full sectors);
//   update   the records are consumed bucket by bucket by the very update of agg_wide_kernel (wide_update: lock-step
//            probing, RED accumulation, value columns gathered by row id): all CTAs work on the same one or two buckets
//            at a time, whose slice of the table (capacity / P slots, ~4 MB) stays in the L2.
// Everything the global-table kernel supports is supported (any key dtype, NULL keys, every function); rows that cannot
// be placed (bucket full, table at its limit) go to the replay list like everywhere else.
struct PartParams {
    GenParams g;             // predicate, key, value columns, functions, global table, replay list
    int log2p;               // buckets = 1 << log2p
    int shift;               // bucket = (hash & mask) >> shift
    uint32_t cap;            // records per bucket
    ulonglong2* recs;        // [P][cap] records {key, row id relative to the chunk}
    unsigned int* cursor;    // [P * 32]: one counter per 128-byte line (8 cursors per bucket, chosen by further hash bits, were
                             // measured: no change, the pass is not bound by same-address atomics)
};

template <int PK>
__global__ void __launch_bounds__(256) agg_part_scatter_kernel(const __grid_constant__ PartParams p) {
    constexpr int R = 8;
    constexpr int64_t TILE = 256 * R;
    const GenParams& g = p.g;
    const uint64_t mask = (uint64_t) g.table.capacity - 1;
    const int64_t n_tiles = (g.n + TILE - 1) / TILE;
    unsigned n_selected = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        __syncwarp();
        const int64_t base = tile * TILE + 2 * (int64_t) threadIdx.x;
        bool sel[R], plain[R];
        uint64_t key[R];
        uint32_t row[R];
        wide_select_and_keys<PK, R>(g, base, (tile + 1) * TILE <= g.n, sel, key, plain);
        uint32_t pos[R], bucket[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            n_selected += sel[j] ? 1u : 0u;
            row[j] = (uint32_t) (base + (j / 2) * 512 + (j & 1));
            bucket[j] = (uint32_t) ((hash_key1(key[j]) & mask) >> p.shift);
            pos[j] = 0xFFFFFFFFu;
            // (a NULL key, or the key value that marks a free slot, has a dedicated slot: straight to the table, below)
            if (sel[j] && plain[j] && key[j] != GT_EMPTY) pos[j] = atomicAdd(p.cursor + (size_t) bucket[j] * 32, 1u);
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            if (!sel[j]) continue;
            if (pos[j] == 0xFFFFFFFFu) general_update_row_cold(g, (int64_t) row[j]);
            else if (pos[j] < p.cap) p.recs[(size_t) bucket[j] * p.cap + pos[j]] = make_ulonglong2(key[j], (unsigned long long) row[j]);
            else replay_append(g.replay, (int64_t) row[j]);   // bucket full (skewed keys): the replay pass takes the row
        }
    }
    count_selected(g, n_selected);
}

template <bool HOT>
__global__ void __launch_bounds__(256) agg_part_update_kernel(const __grid_constant__ PartParams p) {
    constexpr int R = 4;
    constexpr uint32_t TILE = 256 * R;
    const GenParams& g = p.g;
    const uint32_t tiles_per_bucket = (p.cap + TILE - 1) / TILE;
    const int64_t n_tiles = (int64_t) tiles_per_bucket << p.log2p;
    const uint64_t stream = l2_policy_stream();
    // Tiles are handed out in bucket-major order from a queue, not by a fixed stride: with a stride the CTAs drift apart
    // (a quarter of the tiles are empty tails that cost nothing) and the slices of dozens of buckets compete for the L2
    // (measured: 42 % hit rate on the table even with the evict_last hint).
    __shared__ unsigned long long s_tile;
    unsigned long long* queue = reinterpret_cast<unsigned long long*>(p.cursor + ((size_t) 32 << p.log2p));
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_tile = atomicAdd(queue, 1ULL);
        __syncthreads();
        const int64_t tile = (int64_t) s_tile;
        if (tile >= n_tiles) break;
        const uint32_t b = (uint32_t) (tile / tiles_per_bucket);
        const uint32_t first = (uint32_t) (tile % tiles_per_bucket) * TILE;
        uint32_t n_b = p.cursor[(size_t) b * 32];
        if (n_b > p.cap) n_b = p.cap;
        if (first >= n_b) continue;   // CTA-uniform
        const ulonglong2* recs = p.recs + (size_t) b * p.cap;
        bool sel[R], plain[R];
        uint64_t key[R];
        uint32_t row[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const uint32_t r = first + j * 256 + threadIdx.x;
            sel[j] = plain[j] = r < n_b;
            key[j] = 0;
            row[j] = 0;
            if (sel[j]) {
                const ulonglong2 rec = ldg_rec_stream_hint(recs + r, stream);
                key[j] = rec.x;
                row[j] = (uint32_t) rec.y;
            }
        }
        wide_update<R, true, HOT>(g, sel, key, plain, row);
    }
}

// option AGG_PARTITION = 1: partition when the table (all its arrays) is larger than this.  Measured (profiles/r02_tuning.md):
// at 1e6 groups the table (48 MB) lives in the L2 and the global-table kernel alone wins; at 8.4e6 groups (805 MB) it loses
constexpr int64_t PART_MIN_TABLE_BYTES = (int64_t) 96 << 20;
constexpr int64_t PART_BUCKET_BYTES = (int64_t) 4 << 20;   // slice of the table one bucket updates (256 KB - 16 MB measured: no difference)
constexpr int64_t PART_MAX_CHUNK = (int64_t) 1 << 27;    // rows per scatter + reduce round (bounds the scratch: 16 B per selected row)

static int launch_part(const PartParams& pp, bool hot, int sms, cudaStream_t s) {
    auto resident = [&](auto kernel, int* per_sm) -> int {
        if (*per_sm == 0) {
            int b = 0;
            VK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, kernel, 256, 0));
            *per_sm = b > 0 ? b : 1;
        }
        return VK_OK;
    };
    const int64_t tiles = (pp.g.n + 2047) / 2048;
    auto scatter = [&](auto kernel) -> int {
        static int per_sm = 0;
        const int rc = resident(kernel, &per_sm);
        if (rc != VK_OK) return rc;
        const int64_t cap = (int64_t) sms * per_sm;
        kernel<<<(unsigned) (tiles < cap ? tiles : cap), 256, 0, s>>>(pp);
        VK_CHECK_LAUNCH("agg_part_scatter_kernel");
        return VK_OK;
    };
    int rc;
    switch (pp.g.pk) {
        case PK_NONE: rc = scatter(agg_part_scatter_kernel<PK_NONE>); break;
        case PK_MASK: rc = scatter(agg_part_scatter_kernel<PK_MASK>); break;
        case PK_F64_VEC: rc = scatter(agg_part_scatter_kernel<PK_F64_VEC>); break;
        case PK_I64_VEC: rc = scatter(agg_part_scatter_kernel<PK_I64_VEC>); break;
        default: rc = scatter(agg_part_scatter_kernel<PK_GENERIC>); break;
    }
    if (rc != VK_OK) return rc;
    const int64_t update_tiles = (int64_t) ((pp.cap + 1023) / 1024) << pp.log2p;
    auto update = [&](auto kernel) -> int {
        static int per_sm = 0;
        const int urc = resident(kernel, &per_sm);
        if (urc != VK_OK) return urc;
        const int64_t cap = (int64_t) sms * per_sm;
        kernel<<<(unsigned) (update_tiles < cap ? update_tiles : cap), 256, 0, s>>>(pp);
        VK_CHECK_LAUNCH("agg_part_update_kernel");
        return VK_OK;
    };
    return hot ? update(agg_part_update_kernel<true>) : update(agg_part_update_kernel<false>);
}

// ============================================================ one-group kernel
// One launch per aggregate function; slot 0 of the table is the single group.
struct OneParams {
    Pred pred;
    FuncSpec spec;
    Col val;
    int fi;            // function index, -1: count rows only (COUNT(*) / the row counter)
    int64_t n;
    GTable table;
};

// Warp reduction of the per-thread partials and one atomic per warp into slot 0.
__device__ __forceinline__ void onegroup_flush(const GTable& table, int fi, int acc, uint64_t rows, uint64_t nn, uint64_t lo,
                                               uint64_t hi, double fsum) {
    // warp reduction
    for (int d = 16; d > 0; d >>= 1) {
        rows += __shfl_xor_sync(0xffffffffu, rows, d);
        nn += __shfl_xor_sync(0xffffffffu, nn, d);
        if (acc == ACC_SUM_F64) fsum += __shfl_xor_sync(0xffffffffu, fsum, d);
        else if (acc == ACC_SUM_I64) lo += __shfl_xor_sync(0xffffffffu, lo, d);
        else if (acc == ACC_SUM_I128) {
            uint64_t ol = __shfl_xor_sync(0xffffffffu, lo, d), oh = __shfl_xor_sync(0xffffffffu, hi, d);
            uint64_t nl = lo + ol;
            hi += oh + (nl < lo ? 1ULL : 0ULL);
            lo = nl;
        } else if (acc == ACC_MAXORD) {
            uint64_t o = __shfl_xor_sync(0xffffffffu, lo, d);
            lo = o > lo ? o : lo;
        }
    }
    if ((threadIdx.x & 31) != 0) return;
    if (fi < 0) {
        if (rows) atomicAdd(reinterpret_cast<unsigned long long*>(table.count_star), (unsigned long long) rows);
        return;
    }
    if (nn) atomicAdd(reinterpret_cast<unsigned long long*>(table.nnull[fi]), (unsigned long long) nn);
    switch (acc) {
        case ACC_SUM_F64:
            if (rows - nn) atomicAdd(reinterpret_cast<double*>(table.acc_lo[fi]), fsum);
            break;
        case ACC_SUM_I64:
            if (lo) atomicAdd(reinterpret_cast<unsigned long long*>(table.acc_lo[fi]), (unsigned long long) lo);
            break;
        case ACC_SUM_I128:
            if (lo | hi) acc_add_i128(table.acc_lo[fi], table.acc_hi[fi], lo, hi);
            break;
        case ACC_MAXORD:
            if (rows - nn) atomicMax(reinterpret_cast<unsigned long long*>(table.acc_lo[fi]), (unsigned long long) lo);
            break;
        default: break;
    }
}

__global__ void __launch_bounds__(256) agg_onegroup_kernel(const __grid_constant__ OneParams p) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    uint64_t lo = 0, hi = 0, nn = 0, rows = 0;
    double fsum = 0.0;
    const int acc = p.fi < 0 ? ACC_NONE : p.spec.acc;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        if (!pred_row(p.pred, i)) continue;
        ++rows;
        if (acc == ACC_NONE) continue;
        if (!col_valid(p.val, i)) { ++nn; continue; }
        if (acc == ACC_COUNT) continue;
        uint64_t v = acc_load(p.spec, p.val, i);
        switch (acc) {
            case ACC_SUM_F64: fsum += __longlong_as_double((long long) v); break;
            case ACC_SUM_I64: lo += v; break;
            case ACC_SUM_I128: {
                uint64_t nl = lo + v;
                hi += ((!p.spec.in_unsigned && (int64_t) v < 0) ? ~0ULL : 0ULL) + (nl < lo ? 1ULL : 0ULL);
                lo = nl;
                break;
            }
            default: {
                uint64_t o = ord_transform(p.spec.ord, p.spec.is_min, v);
                lo = o > lo ? o : lo;
                break;
            }
        }
    }
    onegroup_flush(p.table, p.fi, acc, rows, nn, lo, hi, fsum);
}

// The common un-grouped case -- no predicate or a vectorisable `column <op> literal` one, value
// column 8 bytes wide without a validity bitmap -- with U lane-contiguous row pairs per thread and
// the 16-byte loads of a round issued together (agg_onegroup_kernel has one 8-byte load per operand
// in flight per thread).  Measured at 1e8 rows, COUNT(*) + SUM(f64) WHERE f64 > c (two launches, 24 B/row):
// 0.676 ms = 3.55 TB/s against 1.305 ms (profiles/r01_variants.md); VINUM_B200_ONEGROUP_FAST=0 selects
// agg_onegroup_kernel.
template <int PK, int U>
__global__ void __launch_bounds__(256) agg_onegroup8_kernel(const __grid_constant__ OneParams p) {
    static_assert(PK == PK_NONE || PK == PK_F64_VEC || PK == PK_I64_VEC, "vectorisable predicates only");
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t npairs = p.n >> 1;
    uint64_t lo = 0, hi = 0, rows = 0;
    double fsum = 0.0;
    const int acc = p.fi < 0 ? ACC_NONE : p.spec.acc;
    const bool need_val = acc >= ACC_SUM_F64;
    auto add = [&](uint64_t v) {
        switch (acc) {
            case ACC_SUM_F64: fsum += __longlong_as_double((long long) v); break;
            case ACC_SUM_I64: lo += v; break;
            case ACC_SUM_I128: {
                const uint64_t nl = lo + v;
                hi += ((!p.spec.in_unsigned && (int64_t) v < 0) ? ~0ULL : 0ULL) + (nl < lo ? 1ULL : 0ULL);
                lo = nl;
                break;
            }
            default: {
                const uint64_t o = ord_transform(p.spec.ord, p.spec.is_min, v);
                lo = o > lo ? o : lo;
                break;
            }
        }
    };
    auto pass = [&](uint64_t bits) -> bool {
        if constexpr (PK == PK_F64_VEC)
            return apply_cmp(p.pred.op, __longlong_as_double((long long) bits), __longlong_as_double((long long) p.pred.scalar.bits));
        else if constexpr (PK == PK_I64_VEC) return apply_cmp(p.pred.op, (int64_t) bits, (int64_t) p.pred.scalar.bits);
        else return true;
    };
    for (int64_t q0 = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; q0 < npairs; q0 += stride * U) {
        uint4 pq[U], vq[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = q0 + u * stride;
            pq[u] = vq[u] = make_uint4(0, 0, 0, 0);
            if (q < npairs) {
                if constexpr (PK != PK_NONE) pq[u] = ldg_stream16(p.pred.col.data + q * 16);
                if (need_val) vq[u] = ldg_stream16(p.val.data + q * 16);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (q0 + u * stride < npairs) {
                const bool ok0 = pass(((uint64_t) pq[u].y << 32) | pq[u].x), ok1 = pass(((uint64_t) pq[u].w << 32) | pq[u].z);
                rows += (uint64_t) ok0 + (uint64_t) ok1;
                if (need_val) {
                    if (ok0) add(((uint64_t) vq[u].y << 32) | vq[u].x);
                    if (ok1) add(((uint64_t) vq[u].w << 32) | vq[u].z);
                }
            }
        }
    }
    if ((p.n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // last, unpaired row
        const int64_t i = p.n - 1;
        if (pred_row(p.pred, i)) {
            ++rows;
            if (need_val) add(acc_load(p.spec, p.val, i));
        }
    }
    onegroup_flush(p.table, p.fi, acc, rows, 0, lo, hi, fsum);
}

// ================================================================== finalize
struct FinalParams {
    GTable table;
    FuncSpec specs[VK_AGG_MAX_FUNCS];
    int64_t num_groups;
    int null_last;                      // single key: NULL group goes to row num_groups-1
    unsigned long long* cursor;         // output row allocator
    int64_t capacity;                   // rows the output arrays hold (packed result: may be < num_groups)
    uint64_t* header;                   // packed result: [0] = number of groups (read from the device counter)
    uint64_t* out_keys[VK_AGG_MAX_KEYS];
    uint8_t* out_key_valid[VK_AGG_MAX_KEYS];
    uint64_t* out_count_star;
    uint64_t* out_lo[VK_AGG_MAX_FUNCS];
    uint64_t* out_hi[VK_AGG_MAX_FUNCS];
    uint8_t* out_valid[VK_AGG_MAX_FUNCS];
};

// Hugeint::TryCast<double>, vinum_cpp/src/common/huge_int.cpp:394-406
__device__ __forceinline__ double hugeint_to_double(__int128 v) {
    uint64_t lower = (uint64_t) v;
    int64_t upper = (int64_t) (v >> 64);
    if (upper == -1) return -(double) (0xFFFFFFFFFFFFFFFFULL - lower) - 1.0;
    return (double) lower + (double) upper * 18446744073709551616.0;  // double(UINT64_MAX) rounds to 2^64
}

__global__ void __launch_bounds__(256) agg_finalize_kernel(const __grid_constant__ FinalParams p) {
    const GTable& t = p.table;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t slots = gt_total_slots(t);
    // packed result: the group count is whatever the stream-ordered kernels before this one left in the
    // device counter -- the host has not seen it yet
    const int64_t num_groups = p.header != nullptr ? (int64_t) *t.num_groups : p.num_groups;
    if (p.header != nullptr && blockIdx.x == 0 && threadIdx.x == 0) p.header[0] = (uint64_t) num_groups;
    for (int64_t s = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += stride) {
        if (!gt_slot_occupied(t, s)) continue;
        uint64_t kv[VK_AGG_MAX_KEYS];
        uint32_t nullmask;
        gt_slot_key(t, s, kv, &nullmask);
        int64_t row;
        if (p.null_last && nullmask) row = num_groups - 1;
        else row = (int64_t) atomicAdd(p.cursor, 1ULL);
        if (row >= num_groups || row >= p.capacity) continue;  // more groups than the caller's block holds: it retries
        for (int k = 0; k < t.n_keys; ++k) {
            p.out_keys[k][row] = kv[k];
            p.out_key_valid[k][row] = !((nullmask >> k) & 1);
        }
        const uint64_t rows = t.count_star[s];
        p.out_count_star[row] = rows;
        for (int f = 0; f < t.n_funcs; ++f) {
            const FuncSpec spec = p.specs[f];
            const uint64_t nn = t.nnull[f] ? t.nnull[f][s] : 0;
            const uint64_t nvalid = rows - nn;
            uint64_t lo = 0, hi = 0;
            bool valid = nvalid > 0;
            const uint64_t alo = t.acc_lo[f] ? t.acc_lo[f][s] : 0;
            const uint64_t ahi = t.acc_hi[f] ? t.acc_hi[f][s] : 0;
            switch (spec.func) {
                case VK_AGG_COUNT_STAR: lo = rows; valid = true; break;
                case VK_AGG_COUNT: lo = nvalid; valid = true; break;
                case VK_AGG_MIN: case VK_AGG_MAX: lo = valid ? ord_inverse(spec.ord, spec.is_min, alo) : 0; break;
                case VK_AGG_SUM: lo = alo; hi = ahi; break;
                case VK_AGG_AVG: {
                    if (!valid) break;
                    double avg;
                    if (spec.acc == ACC_SUM_F64) {
                        avg = __longlong_as_double((long long) alo) / (double) nvalid;  // agg_funcs.h:519-522
                    } else if (spec.acc == ACC_SUM_I64) {
                        double sum = spec.in_unsigned ? (double) alo : (double) (int64_t) alo;
                        avg = sum / (double) nvalid;
                        if (dtype_size(spec.in_dtype) <= 2) avg = (double) (float) avg;  // T_OUT = float_t
                    } else {
                        // hugeint quotient + remainder, agg_funcs.h:524-540
                        __int128 sum = (__int128) (((unsigned __int128) ahi << 64) | alo);
                        __int128 cnt = (__int128) (int64_t) nvalid;
                        __int128 q = sum / cnt, rem = sum % cnt;
                        avg = hugeint_to_double(q);
                        avg += hugeint_to_double(rem) / (double) nvalid;
                    }
                    lo = (uint64_t) __double_as_longlong(avg);
                    break;
                }
                default: break;
            }
            p.out_lo[f][row] = lo;
            if (p.out_hi[f]) p.out_hi[f][row] = hi;
            p.out_valid[f][row] = valid;
        }
    }
}

// ==================================================================== rehash
struct RehashParams {
    GTable src, dst;
};
__global__ void __launch_bounds__(256) agg_rehash_kernel(const __grid_constant__ RehashParams p) {
    const GTable& a = p.src;
    const GTable& b = p.dst;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t slots = gt_total_slots(a);
    for (int64_t s = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += stride) {
        if (!gt_slot_occupied(a, s)) continue;
        uint64_t kv[VK_AGG_MAX_KEYS];
        uint32_t nullmask;
        gt_slot_key(a, s, kv, &nullmask);
        int64_t d = gt_find_or_insert<0>(b, kv, nullmask, hash_keys(kv, nullmask, a.n_keys), b.max_groups);
        if (d < 0) continue;  // cannot happen: dst is larger
        b.count_star[d] = a.count_star[s];
        for (int f = 0; f < a.n_funcs; ++f) {
            if (a.acc_lo[f]) b.acc_lo[f][d] = a.acc_lo[f][s];
            if (a.acc_hi[f]) b.acc_hi[f][d] = a.acc_hi[f][s];
            if (a.nnull[f]) b.nnull[f][d] = a.nnull[f][s];
        }
    }
}

// ============================================================ partial exchange
// Record (u64 words): keys[n_keys] | nullmask | count_star | per func: lo, hi, nnull
struct ExchParams {
    GTable table;
    FuncSpec specs[VK_AGG_MAX_FUNCS];
    int n_ranks;
    int words;
    long long* counts;             // [n_ranks]
    const long long* offsets;      // [n_ranks] record offsets
    unsigned long long* cursors;   // [n_ranks]
    uint64_t* out;
    const uint64_t* in;
    int64_t n_in;
    ReplayList replay;
};
__device__ __forceinline__ int dest_rank(uint64_t hash, int n_ranks) { return (int) ((hash >> 33) % (uint64_t) n_ranks); }

__global__ void __launch_bounds__(256) agg_partition_count_kernel(const __grid_constant__ ExchParams p) {
    const GTable& t = p.table;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t slots = gt_total_slots(t);
    for (int64_t s = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += stride) {
        if (!gt_slot_occupied(t, s)) continue;
        uint64_t kv[VK_AGG_MAX_KEYS];
        uint32_t nullmask;
        gt_slot_key(t, s, kv, &nullmask);
        atomicAdd(reinterpret_cast<unsigned long long*>(p.counts) + dest_rank(hash_keys(kv, nullmask, t.n_keys), p.n_ranks), 1ULL);
    }
}
__global__ void __launch_bounds__(256) agg_export_kernel(const __grid_constant__ ExchParams p) {
    const GTable& t = p.table;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t slots = gt_total_slots(t);
    for (int64_t s = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += stride) {
        if (!gt_slot_occupied(t, s)) continue;
        uint64_t kv[VK_AGG_MAX_KEYS];
        uint32_t nullmask;
        gt_slot_key(t, s, kv, &nullmask);
        int d = dest_rank(hash_keys(kv, nullmask, t.n_keys), p.n_ranks);
        int64_t rec = p.offsets[d] + (int64_t) atomicAdd(p.cursors + d, 1ULL);
        uint64_t* o = p.out + rec * p.words;
        int w = 0;
        for (int k = 0; k < t.n_keys; ++k) o[w++] = kv[k];
        o[w++] = nullmask;
        o[w++] = t.count_star[s];
        for (int f = 0; f < t.n_funcs; ++f) {
            o[w++] = t.acc_lo[f] ? t.acc_lo[f][s] : 0;
            o[w++] = t.acc_hi[f] ? t.acc_hi[f][s] : 0;
            o[w++] = t.nnull[f] ? t.nnull[f][s] : 0;
        }
    }
}
// One partial-group record folded into the table; false when the table is at its load limit.
__device__ __forceinline__ bool merge_record(const GTable& t, const FuncSpec* specs, const uint64_t* o) {
    uint64_t kv[VK_AGG_MAX_KEYS];
    int w = 0;
    for (int k = 0; k < t.n_keys; ++k) kv[k] = o[w++];
    const uint32_t nullmask = (uint32_t) o[w++];
    int64_t slot = gt_find_or_insert<0>(t, kv, nullmask, hash_keys(kv, nullmask, t.n_keys), t.max_groups);
    if (slot < 0) return false;
    atomicAdd(reinterpret_cast<unsigned long long*>(t.count_star + slot), (unsigned long long) o[w++]);
    for (int f = 0; f < t.n_funcs; ++f) {
        const FuncSpec spec = specs[f];
        uint64_t lo = o[w++], hi = o[w++], nn = o[w++];
        if (nn && t.nnull[f]) atomicAdd(reinterpret_cast<unsigned long long*>(t.nnull[f] + slot), (unsigned long long) nn);
        switch (spec.acc) {
            case ACC_SUM_F64:
                atomicAdd(reinterpret_cast<double*>(t.acc_lo[f] + slot), __longlong_as_double((long long) lo));
                break;
            case ACC_SUM_I64:
                atomicAdd(reinterpret_cast<unsigned long long*>(t.acc_lo[f] + slot), (unsigned long long) lo);
                break;
            case ACC_SUM_I128: acc_add_i128(t.acc_lo[f] + slot, t.acc_hi[f] + slot, lo, hi); break;
            case ACC_MAXORD:
                atomicMax(reinterpret_cast<unsigned long long*>(t.acc_lo[f] + slot), (unsigned long long) lo);
                break;
            default: break;
        }
    }
    return true;
}
__global__ void __launch_bounds__(256) agg_merge_kernel(const __grid_constant__ ExchParams p) {
    const GTable& t = p.table;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t r = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; r < p.n_in; r += stride)
        if (!merge_record(t, p.specs, p.in + r * p.words)) replay_append(p.replay, r);
}

// ============================================================== peer exchange
// Low-cardinality group-by across the GPUs of one box WITHOUT a collective call: every
// rank owns a window of device memory that its peers map (cudaIpc) and write into over NVLink.
//   window (u64 words): flag[16] | ack[16] | decision[2] count[16] cursor done pad.. | slots[2][world][slot_words]
//   slot = [n_groups | n_groups records of `words` u64]           (two parities: epoch & 1)
// Rank r's send kernel stores its partial groups straight into slot r of the OWNER's window, then
// releases flag[r] = epoch at system scope; the owner's wait kernel acquires every flag, decides
// (fits / a rank overflowed its slot / the owner's table needs to grow), the merge kernel folds the
// records into the owner's own table, and the ack kernel stores ack[owner] = epoch and the decision
// into every peer's window, which frees the slot parity for epoch + 2 and tells the peers how the
// query went.  No NCCL call and no host round trip until the final read-back.
constexpr int PEER_MAX = 16;
constexpr int PW_FLAG = 0, PW_ACK = 16, PW_DECISION = 32, PW_COUNT = 34, PW_CURSOR = 50, PW_DONE = 51, PW_SLOTS = 64;
constexpr uint64_t PEER_OK = 1, PEER_OVERFLOW = 2, PEER_NEED_GROW = 3, PEER_TIMEOUT = 4;
constexpr unsigned long long PEER_TIMEOUT_NS = 4000000000ULL;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const uint64_t* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint64_t* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Spin until *p >= want; false after PEER_TIMEOUT_NS (a dead peer must not hang this GPU).
__device__ __forceinline__ bool wait_at_least(const uint64_t* p, unsigned long long want) {
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(p) < want) {
        __nanosleep(200);
        if (global_ns() - t0 > PEER_TIMEOUT_NS) return false;
    }
    return true;
}

struct PeerSendParams {
    GTable table;
    int words;
    int has_table;
    int64_t cap;                 // records a slot holds
    uint64_t* slot;              // the owner's slot for this rank (peer memory)
    uint64_t* flag;              // the owner's flag word for this rank (peer memory)
    const uint64_t* ack;         // this rank's own window: what the owner has merged so far
    unsigned long long* cursor;  // local scratch
    unsigned long long* done;
    unsigned long long epoch;
};
__global__ void __launch_bounds__(256) agg_peer_send_kernel(const __grid_constant__ PeerSendParams p) {
    __shared__ int s_ok;
    __shared__ int s_last;
    // the slot parity of this epoch was last used by epoch - 2: the owner must have merged that one
    if (threadIdx.x == 0) s_ok = p.epoch < 3 || wait_at_least(p.ack, p.epoch - 2);
    __syncthreads();
    if (s_ok && p.has_table) {
        const GTable& t = p.table;
        const int64_t stride = (int64_t) gridDim.x * blockDim.x;
        const int64_t slots = gt_total_slots(t);
        for (int64_t s = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; s < slots; s += stride) {
            if (!gt_slot_occupied(t, s)) continue;
            const int64_t rec = (int64_t) atomicAdd(p.cursor, 1ULL);
            if (rec >= p.cap) continue;   // counted, not stored: the owner sees n_groups > cap and everybody falls back
            uint64_t kv[VK_AGG_MAX_KEYS];
            uint32_t nullmask;
            gt_slot_key(t, s, kv, &nullmask);
            uint64_t* o = p.slot + 1 + rec * p.words;
            int w = 0;
            for (int k = 0; k < t.n_keys; ++k) o[w++] = kv[k];
            o[w++] = nullmask;
            o[w++] = t.count_star[s];
            for (int f = 0; f < t.n_funcs; ++f) {
                o[w++] = t.acc_lo[f] ? t.acc_lo[f][s] : 0;
                o[w++] = t.acc_hi[f] ? t.acc_hi[f][s] : 0;
                o[w++] = t.nnull[f] ? t.nnull[f][s] : 0;
            }
        }
    }
    // last block out publishes: every block's remote stores are fenced at system scope before its
    // arrival is counted, so they are visible to the owner before the flag is
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(p.done, 1ULL) == (unsigned long long) gridDim.x - 1;
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence_system();
        const unsigned long long n = s_ok ? *reinterpret_cast<volatile unsigned long long*>(p.cursor) : ~0ULL;
        p.slot[0] = n;
        __threadfence_system();
        st_release_sys(p.flag, p.epoch);
    }
}

struct PeerMergeParams {
    GTable table;
    FuncSpec specs[VK_AGG_MAX_FUNCS];
    int words;
    int rank, world;
    int64_t cap;
    int64_t slot_words;
    uint64_t* window;            // this rank's own window
    uint64_t* peer_window[PEER_MAX];
    unsigned long long epoch;
    unsigned long long* lost;
};
__device__ __forceinline__ const uint64_t* peer_slot(const PeerMergeParams& p, int r) {
    return p.window + PW_SLOTS + ((int64_t) (p.epoch & 1) * p.world + r) * p.slot_words;
}
// One warp: lane r waits for rank r's message, then lane 0 decides.
__global__ void __launch_bounds__(32) agg_peer_wait_kernel(const __grid_constant__ PeerMergeParams p) {
    const int r = threadIdx.x;
    bool ok = true;
    unsigned long long n = 0;
    if (r < p.world && r != p.rank) {
        ok = wait_at_least(p.window + PW_FLAG + r, p.epoch);
        if (ok) n = peer_slot(p, r)[0];
        p.window[PW_COUNT + r] = n;
    }
    const unsigned all_ok = __all_sync(0xffffffffu, ok);
    const unsigned fits = __all_sync(0xffffffffu, n <= (unsigned long long) p.cap);
    unsigned long long total = n <= (unsigned long long) p.cap ? n : 0;
    for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(0xffffffffu, total, d);
    if (r == 0) {
        uint64_t decision = PEER_OK;
        if (!all_ok) decision = PEER_TIMEOUT;
        else if (!fits) decision = PEER_OVERFLOW;
        else if (p.table.max_groups - (int64_t) *p.table.num_groups < (int64_t) total) decision = PEER_NEED_GROW;
        p.window[PW_DECISION + (p.epoch & 1)] = decision;
    }
}
__global__ void __launch_bounds__(256) agg_peer_merge_kernel(const __grid_constant__ PeerMergeParams p) {
    if (p.window[PW_DECISION + (p.epoch & 1)] != PEER_OK) return;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int r = 0; r < p.world; ++r) {
        if (r == p.rank) continue;
        const int64_t n = (int64_t) p.window[PW_COUNT + r];
        const uint64_t* recs = peer_slot(p, r) + 1;
        for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
            if (!merge_record(p.table, p.specs, recs + i * p.words)) atomicAdd(p.lost, 1ULL);
    }
}
// Lane r tells rank r: merged (ack) and how it went (decision); stream order puts this after the merge.
__global__ void __launch_bounds__(32) agg_peer_ack_kernel(const __grid_constant__ PeerMergeParams p) {
    const int r = threadIdx.x;
    if (r >= p.world || r == p.rank) return;
    const uint64_t decision = p.window[PW_DECISION + (p.epoch & 1)];
    uint64_t* w = p.peer_window[r];
    w[PW_DECISION + (p.epoch & 1)] = decision;
    __threadfence_system();
    st_release_sys(w + PW_ACK + p.rank, p.epoch);
}
// A non-owner rank: wait until the owner has merged this epoch (its decision word is then in our window).
__global__ void __launch_bounds__(32) agg_peer_wait_ack_kernel(uint64_t* window, int owner, unsigned long long epoch) {
    if (threadIdx.x == 0 && !wait_at_least(window + PW_ACK + owner, epoch)) window[PW_DECISION + (epoch & 1)] = PEER_TIMEOUT;
}

}  // namespace vk

using namespace vk;

// =================================================================== host side
struct VkAgg {
    int n_keys = 0;
    int key_dtypes[VK_AGG_MAX_KEYS] = {0};
    int n_funcs = 0;
    FuncSpec specs[VK_AGG_MAX_FUNCS];
    GTable t{};
    bool table_ready = false;
    int64_t expected_groups = 0;
    int64_t groups_ub = 0;            // upper bound on groups in the table
    unsigned long long* d_ctr = nullptr;  // [0] num_groups [1] list count [2] lost [3] spilled [4] finalize cursor [5..] scratch
    unsigned long long* h_ctr = nullptr;  // pinned mirror
    int device = 0;                   // device the object (its counter block, list and table) lives on
    bool ctr_ready = false;           // d_ctr zeroed on the first stream this object is used on
    cudaStream_t last_stream = nullptr;   // every call so far ran on this stream ...
    bool stream_seen = false, multi_stream = false;  // ... unless multi_stream
    uint64_t list_max = 0;            // replay list cap (entries), from free memory at first use
    uint32_t* list = nullptr;
    uint64_t list_cap = 0;
    int last_path = 0;
    // fast (shared-memory) path configuration; env overrides are for tuning runs
    int fast_log2s = 12;              // CTA key table slots
    int fast_pf = -1;                 // L2 bulk prefetch distance of the fused kernel, in tiles (-1: automatic)
    double fast_selectivity = 1.0;    // share of rows that passed the fused predicate so far
    int fast_warps = FA_MAX_THREADS / 32;  // warps per CTA (fewer warps = more groups per warp-private table)
    bool fast_warps_fixed = false;
    int fast_direct_policy = 1;       // 1: use direct (key - base) group ids when the key range allows, 0: always hash
    int64_t learn_rows = (int64_t) 1 << 20;   // rows of the learning launch
    int wide_rows = 1;                        // option AGG_WIDE
    // hash mode: read-only cuckoo dictionary of the keys seen so far (built by the host after the learning
    // launch, vk_agg_fast.cuh); rows with other keys go to the global table
    bool dict_ready = false, dict_failed = false;
    int dict_policy = 1;
    int hot_policy = 1;               // option AGG_HOT
    double hot_share = 0.0;           // largest group's share of the selected rows (learning launch)
    int match_policy = 1;             // option AGG_ENTRY: layout / update protocol of the COUNT + SUM(f64) entry
    int dict_n = 0;                   // keys in the dictionary
    int dict_log2s = 0;
    uint8_t* dict_dev = nullptr;      // [S] u64 keys | [S] u16 ids
    uint64_t seed_a = 0;
    std::vector<uint64_t> dict_hk;
    std::vector<uint16_t> dict_hg;
    bool direct_known = false;        // key range of the table measured
    bool direct_ok = false;
    uint64_t direct_min = 0, direct_span = 0;  // smallest key (signed order) and max - min
    bool fast_disabled = false;       // cardinality turned out too high for the shared-memory table
    uint64_t fast_rows_seen = 0, fast_spilled = 0;
    int64_t fast_groups_seen = 0;
    // global-table path while the number of groups is still rising: chunks grow geometrically and the
    // table is sized at every chunk boundary from an estimate of the final number of groups
    bool sizing = true;
    double est_groups = 0;            // latest estimate of the final number of groups (0: none yet)
    int part_policy = 1;              // option AGG_PARTITION
    int64_t rows_seen = 0;            // rows of every update so far
    int64_t sized_groups = 0;         // groups / selected rows at the last estimate
    uint64_t sized_selected = 0;
    // optional per-launch timing of the update kernels (bench.py roofline): events are
    // recorded on the launch stream and resolved lazily, so profiling adds no sync
    bool profile = false;
    struct ProfSpan { cudaEvent_t e0, e1; int64_t rows; int path; };
    std::vector<ProfSpan> prof_pending;
    std::vector<ProfSpan> prof_free;
    double prof_ms[4] = {0, 0, 0, 0};
    int64_t prof_launches[4] = {0, 0, 0, 0};
    int64_t prof_rows[4] = {0, 0, 0, 0};
};

namespace {

bool debug_on() {
    return opt(OPT_DEBUG) != 0;
}
#define VK_DBG(...) do { if (debug_on()) { fprintf(stderr, "[vk_agg] " __VA_ARGS__); fputc('\n', stderr); fflush(stderr); } } while (0)

constexpr double kMaxLoad = 0.5;
constexpr int CTR_GROUPS = 0, CTR_LIST = 1, CTR_LOST = 2, CTR_SPILL = 3, CTR_CURSOR = 4, CTR_SELECTED = 6, CTR_RANGE = 8, CTR_MAXCOUNT = 10, CTR_WORDS = 16;

// Counter blocks (16 device words + a pinned host mirror) are recycled across aggregate
// objects: cudaMalloc / cudaHostAlloc / cudaFree cost far more than a whole 1e9-row scan.
struct CtrBlock { int device; unsigned long long* d; unsigned long long* h; };
// leaked on purpose: objects may be destroyed after static destructors have run
std::mutex& g_ctr_mu = *new std::mutex();
std::vector<CtrBlock>& g_ctr_free = *new std::vector<CtrBlock>();

int ctr_acquire(VkAgg* a) {
    int dev = 0;
    cudaGetDevice(&dev);
    a->device = dev;
    {
        std::lock_guard<std::mutex> lk(g_ctr_mu);
        for (size_t i = 0; i < g_ctr_free.size(); ++i) {
            if (g_ctr_free[i].device == dev) {
                a->d_ctr = g_ctr_free[i].d;
                a->h_ctr = g_ctr_free[i].h;
                g_ctr_free.erase(g_ctr_free.begin() + i);
                return VK_OK;
            }
        }
    }
    cudaError_t e = cudaMalloc((void**) &a->d_ctr, CTR_WORDS * sizeof(unsigned long long));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(counters)");
    e = cudaHostAlloc((void**) &a->h_ctr, CTR_WORDS * sizeof(unsigned long long), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaFree(a->d_ctr);
        a->d_ctr = nullptr;
        return cuda_fail(e, "cudaHostAlloc(counters)");
    }
    return VK_OK;
}
void ctr_release(VkAgg* a) {
    if (!a->d_ctr) return;
    const int dev = a->device;
    std::lock_guard<std::mutex> lk(g_ctr_mu);
    g_ctr_free.push_back(CtrBlock{dev, a->d_ctr, a->h_ctr});
    a->d_ctr = nullptr;
    a->h_ctr = nullptr;
}
// Every entry point that touches the counters on a stream goes through here first.
int ctr_init(VkAgg* a, cudaStream_t s) {
    if (a->stream_seen && a->last_stream != s) a->multi_stream = true;
    a->stream_seen = true;
    a->last_stream = s;
    if (!a->ctr_ready) {
        VK_CUDA(cudaMemsetAsync(a->d_ctr, 0, CTR_WORDS * sizeof(unsigned long long), s));
        a->ctr_ready = true;
    }
    return VK_OK;
}

int64_t pow2_ceil(int64_t x) {
    int64_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

int alloc_table(VkAgg* a, GTable* t, int64_t capacity, cudaStream_t s) {
    memset(t, 0, sizeof(GTable));
    t->capacity = capacity;
    t->max_groups = (int64_t) (capacity * kMaxLoad);
    t->n_keys = a->n_keys;
    t->n_funcs = a->n_funcs;
    t->single = a->n_keys == 1;
    t->num_groups = a->d_ctr + CTR_GROUPS;
    const int64_t slots = capacity + (t->single ? 2 : 0);
    auto zalloc = [&](void** p, size_t bytes) -> int {
        VK_CUDA(cudaMallocAsync(p, bytes, s));
        VK_CUDA(cudaMemsetAsync(*p, 0, bytes, s));
        return VK_OK;
    };
    int rc;
    if (t->single) {
        // key array doubles as the slot tag: all-ones == free
        if ((rc = zalloc((void**) &t->state, 2 * sizeof(uint32_t)))) return rc;
        VK_CUDA(cudaMallocAsync((void**) &t->keys, (size_t) slots * sizeof(uint64_t), s));
        VK_CUDA(cudaMemsetAsync(t->keys, 0xFF, (size_t) slots * sizeof(uint64_t), s));
    } else {
        if ((rc = zalloc((void**) &t->state, capacity * sizeof(uint32_t)))) return rc;
        if (a->n_keys) {
            if ((rc = zalloc((void**) &t->keys, (size_t) a->n_keys * capacity * sizeof(uint64_t)))) return rc;
            if ((rc = zalloc((void**) &t->knull, capacity * sizeof(uint32_t)))) return rc;
        }
    }
    if ((rc = zalloc((void**) &t->count_star, slots * sizeof(uint64_t)))) return rc;
    for (int f = 0; f < a->n_funcs; ++f) {
        int acc = a->specs[f].acc;
        if (acc >= ACC_SUM_F64)
            if ((rc = zalloc((void**) &t->acc_lo[f], slots * sizeof(uint64_t)))) return rc;
        if (acc == ACC_SUM_I128)
            if ((rc = zalloc((void**) &t->acc_hi[f], slots * sizeof(uint64_t)))) return rc;
        if (acc != ACC_NONE)
            if ((rc = zalloc((void**) &t->nnull[f], slots * sizeof(uint64_t)))) return rc;
    }
    return VK_OK;
}

void free_table(GTable* t, cudaStream_t s) {
    if (t->state) cudaFreeAsync(t->state, s);
    if (t->keys) cudaFreeAsync(t->keys, s);
    if (t->knull) cudaFreeAsync(t->knull, s);
    if (t->count_star) cudaFreeAsync(t->count_star, s);
    for (int f = 0; f < VK_AGG_MAX_FUNCS; ++f) {
        if (t->acc_lo[f]) cudaFreeAsync(t->acc_lo[f], s);
        if (t->acc_hi[f]) cudaFreeAsync(t->acc_hi[f], s);
        if (t->nnull[f]) cudaFreeAsync(t->nnull[f], s);
    }
    memset(t, 0, sizeof(GTable));
}

int read_counters(VkAgg* a, cudaStream_t s) {
    VK_CUDA(cudaMemcpyAsync(a->h_ctr, a->d_ctr, CTR_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    VK_CUDA(cudaStreamSynchronize(s));
    return VK_OK;
}

// Grow (rehash) so that at least `need_free` more groups fit.
int grow_table(VkAgg* a, int64_t groups_now, int64_t need_free, cudaStream_t s) {
    int64_t cap = a->t.capacity;
    while ((int64_t) (cap * kMaxLoad) - groups_now < need_free) cap <<= 1;
    if (cap == a->t.capacity) return VK_OK;
    GTable nt;
    // the group counter is shared: reset it, the rehash re-counts
    int rc = alloc_table(a, &nt, cap, s);
    if (rc != VK_OK) return rc;
    VK_CUDA(cudaMemsetAsync(a->d_ctr + CTR_GROUPS, 0, sizeof(unsigned long long), s));
    RehashParams rp{a->t, nt};
    int64_t need = (gt_total_slots(a->t) + 255) / 256, capb = (int64_t) sm_count() * 8;
    agg_rehash_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(rp);
    VK_CHECK_LAUNCH("agg_rehash_kernel");
    free_table(&a->t, s);
    a->t = nt;
    return VK_OK;
}

// Number of distinct keys behind (selected rows n, groups g so far) if the keys are drawn evenly from G
// values: g = G (1 - exp(-n / G)).  Solved for x = n / G by bisection on (1 - exp(-x)) / x = g / n.  Skewed
// keys make this an under-estimate (the caller treats it as a lower bound and keeps the replay list as the
// safety net); g == n gives no upper bound at all and returns 0.
double estimate_groups(double n, double g) {
    if (g <= 0 || n <= 0) return 0;
    if (g >= n) return 0;
    const double r = g / n;
    double lo = 1e-9, hi = 2.0 / r + 64.0;   // x -> 1 / r when the keys have all been seen many times (G -> g)
    for (int it = 0; it < 100; ++it) {
        const double x = 0.5 * (lo + hi);
        if ((1.0 - exp(-x)) / x > r) lo = x;
        else hi = x;
    }
    return n / (0.5 * (lo + hi));
}

// Bytes of one table slot (key, row counter, per function: accumulator lanes + NULL counter).
int64_t slot_bytes(const VkAgg* a) {
    int64_t b = a->n_keys == 1 ? 16 : 8 * a->n_keys + 16;
    for (int f = 0; f < a->n_funcs; ++f) b += a->specs[f].acc == ACC_SUM_I128 ? 24 : 16;
    return b;
}

// Replay lists are sized for the worst case of a whole chunk (up to 4 GB for a 1e9-row batch)
// and almost never written.  Carving such a block out of the stream-ordered pool on every
// query fragments it (measured: sporadic multi-ms stalls), so one list per device is kept
// across aggregate objects, like the counter blocks.
struct ListBlock { int device; uint32_t* ptr; uint64_t cap; };
std::vector<ListBlock>& g_list_free = *new std::vector<ListBlock>();

int ensure_list(VkAgg* a, uint64_t want, cudaStream_t s) {
    if (a->list_cap >= want) return VK_OK;
    if (a->list) VK_CUDA(cudaFreeAsync(a->list, s));
    a->list = nullptr;
    a->list_cap = 0;
    const int dev = a->device;
    {
        std::lock_guard<std::mutex> lk(g_ctr_mu);
        for (size_t i = 0; i < g_list_free.size(); ++i) {
            if (g_list_free[i].device == dev && g_list_free[i].cap >= want) {
                a->list = g_list_free[i].ptr;
                a->list_cap = g_list_free[i].cap;
                g_list_free.erase(g_list_free.begin() + i);
                return VK_OK;
            }
        }
    }
    VK_CUDA(cudaMallocAsync((void**) &a->list, want * sizeof(uint32_t), s));
    a->list_cap = want;
    return VK_OK;
}
// Called after the object's stream has been drained.
void release_list(VkAgg* a, cudaStream_t s) {
    if (!a->list) return;
    const int dev = a->device;
    uint32_t* drop = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_ctr_mu);
        size_t i = 0;
        for (; i < g_list_free.size(); ++i)
            if (g_list_free[i].device == dev) break;
        if (i == g_list_free.size()) {
            g_list_free.push_back(ListBlock{dev, a->list, a->list_cap});
        } else if (g_list_free[i].cap < a->list_cap) {  // keep the larger one
            drop = g_list_free[i].ptr;
            g_list_free[i] = ListBlock{dev, a->list, a->list_cap};
        } else {
            drop = a->list;
        }
    }
    if (drop) cudaFreeAsync(drop, s);
    a->list = nullptr;
    a->list_cap = 0;
}

ReplayList make_replay(VkAgg* a) {
    ReplayList l;
    l.rows = a->list;
    l.count = a->d_ctr + CTR_LIST;
    l.lost = a->d_ctr + CTR_LOST;
    l.spilled = a->d_ctr + CTR_SPILL;
    l.selected = a->d_ctr + CTR_SELECTED;
    l.capacity = a->list_cap;
    return l;
}

// Fast-path plan for one update call: which columns are streamed and which accumulator
// cells a group entry holds.
struct FastPlan {
    int key_mode = 0;
    int n_cols = 0;
    VkColumn cols[FA_MAX_COLS];
    int col_mode[FA_MAX_COLS] = {0, 0, 0};
    int n_cells = 0;
    FastCell cells[FA_MAX_CELLS];
    int nw = 2;  // 64-bit words per entry
    int mode = FM_RUNTIME;
    bool sumf64 = false;
};

// Largest number of dense group ids a CTA of `warps` warps can hold.
int fast_gmax(int log2s, bool direct, int nw, int warps) {
    const int64_t budget = (int64_t) max_smem_optin() - 256;
    const int64_t avail = budget - (int64_t) fast_table_bytes(log2s, direct);
    if (avail <= 0) return 0;
    int64_t g = avail / ((int64_t) warps * (nw * 8));
    g &= ~(int64_t) 15;
    while (g > 0 && (int64_t) fast_smem_bytes(log2s, direct, (int) g, nw, warps) > budget) g -= 16;
    if (!direct) {
        const int64_t cap = ((int64_t) 1 << log2s) / 2;  // keep the key table at most half full
        if (g > cap) g = cap;
    }
    if (g > 0xFFF0) g = 0xFFF0;
    return (int) g;
}

int launch_fast(const FastParams& p, const FastLaunch& l, cudaStream_t s) {
    switch (l.pk) {
        case PK_NONE:
            return l.mode == FM_ALL8 ? (l.direct ? launch_fast_none_all8_d(p, l, s) : launch_fast_none_all8_t(p, l, s))
                 : l.mode == FM_KEY4 ? (l.direct ? launch_fast_none_key4_d(p, l, s) : launch_fast_none_key4_t(p, l, s))
                                     : launch_fast_none_rt(p, l, s);
        case PK_F64_VEC:
            return l.mode == FM_ALL8 ? (l.direct ? launch_fast_f64_all8_d(p, l, s) : launch_fast_f64_all8_t(p, l, s))
                 : l.mode == FM_KEY4 ? (l.direct ? launch_fast_f64_key4_d(p, l, s) : launch_fast_f64_key4_t(p, l, s))
                                     : launch_fast_f64_rt(p, l, s);
        case PK_MASK: return launch_fast_mask_rt(p, l, s);
        case PK_I64_VEC: return launch_fast_i64_rt(p, l, s);
        default: return launch_fast_gen_rt(p, l, s);
    }
}

VkColumn slice_col(const VkColumn& c, int64_t off, int64_t len) {
    VkColumn r = c;
    r.offset += off;
    r.length = len;
    return r;
}

// profiling spans ----------------------------------------------------------------
int prof_begin(VkAgg* a, cudaStream_t s, int64_t rows, int path) {
    if (!a->profile) return -1;
    VkAgg::ProfSpan sp;
    if (!a->prof_free.empty()) {
        sp = a->prof_free.back();
        a->prof_free.pop_back();
    } else {
        if (cudaEventCreate(&sp.e0) != cudaSuccess || cudaEventCreate(&sp.e1) != cudaSuccess) return -1;
    }
    sp.rows = rows;
    sp.path = path;
    cudaEventRecord(sp.e0, s);
    a->prof_pending.push_back(sp);
    return (int) a->prof_pending.size() - 1;
}
void prof_end(VkAgg* a, cudaStream_t s, int idx) {
    if (idx >= 0) cudaEventRecord(a->prof_pending[idx].e1, s);
}
void prof_resolve(VkAgg* a) {
    for (auto& sp : a->prof_pending) {
        float ms = 0;
        if (cudaEventSynchronize(sp.e1) == cudaSuccess && cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) {
            a->prof_ms[sp.path] += ms;
            a->prof_launches[sp.path] += 1;
            a->prof_rows[sp.path] += sp.rows;
        }
        a->prof_free.push_back(sp);
    }
    a->prof_pending.clear();
}

// Key range of the (single-key) table -> a->direct_*; decides whether later chunks may use
// direct group ids.  One small scan of the table plus a counter read-back.
void apply_key_range(VkAgg* a);
int measure_key_range(VkAgg* a, cudaStream_t s) {
    const long long init[2] = {INT64_MAX, INT64_MIN};
    VK_CUDA(cudaMemcpyAsync(a->d_ctr + CTR_RANGE, init, sizeof(init), cudaMemcpyHostToDevice, s));
    int64_t need = (a->t.capacity + 1 + 255) / 256, capb = (int64_t) sm_count() * 8;
    agg_key_range_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(a->t, reinterpret_cast<long long*>(a->d_ctr + CTR_RANGE));
    VK_CHECK_LAUNCH("agg_key_range_kernel");
    int rc = read_counters(a, s);
    if (rc != VK_OK) return rc;
    apply_key_range(a);
    return VK_OK;
}
void apply_key_range(VkAgg* a) {
    const long long mn = (long long) a->h_ctr[CTR_RANGE], mx = (long long) a->h_ctr[CTR_RANGE + 1];
    a->direct_known = true;
    a->direct_ok = mn <= mx && (uint64_t) mx - (uint64_t) mn < 0xFFF0ULL;
    a->direct_min = (uint64_t) mn;
    a->direct_span = (uint64_t) mx - (uint64_t) mn;
    VK_DBG("key range: min=%lld max=%lld direct_ok=%d", mn, mx, (int) a->direct_ok);
}

// Hash mode: place the table's keys into a cuckoo dictionary (two slot choices per key, S = 1 << log2s
// slots) and upload it.  Leaves dict_ready false when there are too many keys or no seed pair works.
int build_dict(VkAgg* a, int64_t n_groups, int log2s, cudaStream_t s) {
    a->dict_ready = false;
    const int64_t S = (int64_t) 1 << log2s;
    if (n_groups < 1 || n_groups > S / 4 || n_groups >= 0xFFF0) return VK_OK;
    uint64_t* d_keys = nullptr;
    VK_CUDA(cudaMallocAsync((void**) &d_keys, (size_t) n_groups * 8, s));
    VK_CUDA(cudaMemsetAsync(a->d_ctr + CTR_CURSOR, 0, sizeof(unsigned long long), s));
    int64_t need = (a->t.capacity + 1 + 255) / 256, capb = (int64_t) sm_count() * 8;
    agg_export_keys_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(a->t, d_keys, a->d_ctr + CTR_CURSOR,
                                                                                   (unsigned long long) n_groups);
    VK_CHECK_LAUNCH("agg_export_keys_kernel");
    std::vector<uint64_t> keys((size_t) n_groups);
    VK_CUDA(cudaMemcpyAsync(keys.data(), d_keys, (size_t) n_groups * 8, cudaMemcpyDeviceToHost, s));
    int rc = read_counters(a, s);   // synchronises: the keys are on the host
    VK_CUDA(cudaFreeAsync(d_keys, s));
    if (rc != VK_OK) return rc;
    if ((int64_t) a->h_ctr[CTR_CURSOR] != n_groups) return VK_OK;   // the table moved on (or holds a NULL group): no dictionary
    const uint32_t shift = 32 - log2s;
    a->dict_hk.assign((size_t) S, LK_EMPTY);
    a->dict_hg.assign((size_t) S, 0);
    bool placed = false;
    uint64_t sa = 0x9E3779B97F4A7C15ULL;
    for (int attempt = 0; attempt < 32 && !placed; ++attempt) {
        std::fill(a->dict_hk.begin(), a->dict_hk.end(), LK_EMPTY);
        placed = true;
        for (int64_t i = 0; i < n_groups && placed; ++i) {
            uint64_t key = keys[(size_t) i];
            uint16_t gid = (uint16_t) i;
            if (key == LK_EMPTY) continue;   // the sentinel cannot live in the dictionary: its rows take the global path
            uint32_t slot = dict_slot_a(dict_fold(key, sa), shift);
            bool done = false;
            for (int kick = 0; kick < 256; ++kick) {
                if (a->dict_hk[slot] == LK_EMPTY) {
                    a->dict_hk[slot] = key;
                    a->dict_hg[slot] = gid;
                    done = true;
                    break;
                }
                std::swap(key, a->dict_hk[slot]);
                std::swap(gid, a->dict_hg[slot]);
                const uint32_t x = dict_fold(key, sa);
                const uint32_t ha = dict_slot_a(x, shift), hb = dict_slot_b(x, shift);
                slot = slot == ha ? hb : ha;   // the evicted key moves to its other slot
            }
            placed = done;
        }
        if (!placed) sa = splitmix64(sa + attempt);
    }
    if (!placed) {
        a->dict_failed = true;
        return VK_OK;
    }
    if (a->dict_dev == nullptr || a->dict_log2s != log2s) {
        if (a->dict_dev) VK_CUDA(cudaFreeAsync(a->dict_dev, s));
        a->dict_dev = nullptr;
        VK_CUDA(cudaMallocAsync((void**) &a->dict_dev, (size_t) S * 10, s));
        a->dict_log2s = log2s;
    }
    VK_CUDA(cudaMemcpyAsync(a->dict_dev, a->dict_hk.data(), (size_t) S * 8, cudaMemcpyHostToDevice, s));
    VK_CUDA(cudaMemcpyAsync(a->dict_dev + (size_t) S * 8, a->dict_hg.data(), (size_t) S * 2, cudaMemcpyHostToDevice, s));
    a->seed_a = sa;
    a->dict_n = (int) n_groups;
    a->dict_ready = true;
    VK_DBG("dictionary: %d keys in %lld slots", a->dict_n, (long long) S);
    return VK_OK;
}

bool aligned_for_pairs(const VkColumn& c) {
    const int es = dtype_size(c.dtype);
    uintptr_t addr = reinterpret_cast<uintptr_t>(c.data) + (uintptr_t) c.offset * es;
    return (addr % (2 * es)) == 0;
}

// Can this update run on the shared-memory path?  Single fixed-width key without NULLs,
// value columns without NULLs, at most FA_MAX_COLS distinct columns / FA_MAX_CELLS cells.
bool build_fast_plan(const VkAgg* a, const VkColumn* keys, const VkColumn* values, FastPlan* out) {
    FastPlan pl;
    if (keys[0].validity != nullptr || !aligned_for_pairs(keys[0])) return false;
    switch (keys[0].dtype) {
        case VK_I64: case VK_U64: case VK_F64: pl.key_mode = 0; break;
        case VK_I32: pl.key_mode = 1; break;
        case VK_U32: case VK_F32: pl.key_mode = 2; break;
        default: return false;
    }
    for (int f = 0; f < a->n_funcs; ++f) {
        const FuncSpec& sp = a->specs[f];
        if (sp.acc == ACC_NONE) continue;
        const VkColumn& vc = values[f];
        if (vc.validity != nullptr) return false;
        if (sp.acc == ACC_COUNT) continue;  // no NULLs: COUNT(col) == COUNT(*)
        int mode;
        switch (vc.dtype) {
            case VK_I64: case VK_U64: case VK_F64: mode = 0; break;
            case VK_I32: mode = 1; break;
            case VK_U32: mode = 2; break;
            case VK_F32: mode = 3; break;
            default: return false;
        }
        if (!aligned_for_pairs(vc)) return false;
        FastCell cell{};
        switch (sp.acc) {
            case ACC_SUM_F64:
                if (!dtype_is_float(vc.dtype)) return false;
                cell.op = CELL_ADD_F64;
                break;
            case ACC_SUM_I64:
                if (mode != 1 && mode != 2) return false;
                cell.op = CELL_ADD_I64;
                break;
            case ACC_SUM_I128:
                if (mode != 0 || dtype_is_float(vc.dtype)) return false;
                cell.op = CELL_ADD_I128;
                cell.in_unsigned = sp.in_unsigned;
                break;
            case ACC_MAXORD:
                cell.op = CELL_MAXORD;
                cell.ord = sp.ord;
                cell.is_min = sp.is_min;
                break;
            default: return false;
        }
        int ci = 0;
        for (; ci < pl.n_cols; ++ci)
            if (pl.cols[ci].data == vc.data && pl.cols[ci].offset == vc.offset && pl.cols[ci].dtype == vc.dtype) break;
        if (ci == pl.n_cols) {
            if (pl.n_cols == FA_MAX_COLS) return false;
            pl.cols[pl.n_cols] = vc;
            pl.col_mode[pl.n_cols] = mode;
            ++pl.n_cols;
        }
        cell.col = ci;
        int k = 0;
        for (; k < pl.n_cells; ++k) {
            const FastCell& e = pl.cells[k];
            if (e.op == cell.op && e.col == cell.col && e.ord == cell.ord && e.is_min == cell.is_min &&
                e.in_unsigned == cell.in_unsigned) break;
        }
        if (k == pl.n_cells) {
            const int need = cell.op == CELL_ADD_I128 ? 2 : 1;
            if (pl.n_cells + need > FA_MAX_CELLS) return false;
            pl.cells[pl.n_cells] = cell;
            pl.cells[pl.n_cells].func_mask = 0;
            ++pl.n_cells;
            if (need == 2) {
                FastCell hi{};
                hi.op = CELL_I128_HI;
                hi.col = ci;
                pl.cells[pl.n_cells++] = hi;
            }
        }
        pl.cells[k].func_mask |= 1u << f;
    }
    pl.nw = pl.n_cells <= 1 ? 2 : 4;
    bool vals8 = true;
    for (int v = 0; v < pl.n_cols; ++v) vals8 = vals8 && pl.col_mode[v] == 0;
    pl.mode = !vals8 ? FM_RUNTIME : (pl.key_mode == 0 ? FM_ALL8 : FM_KEY4);
    pl.sumf64 = pl.n_cells == 1 && pl.cells[0].op == CELL_ADD_F64 && pl.n_cols == 1;
    *out = pl;
    return true;
}

}  // namespace

extern "C" {

int vk_agg_create(VkAgg** out, int n_keys, const int32_t* key_dtypes, int n_funcs, const int32_t* funcs,
                  const int32_t* func_in_dtypes, int64_t expected_groups) {
    VK_REQUIRE(out, "vk_agg_create: out is NULL");
    *out = nullptr;
    VK_REQUIRE(n_keys >= 0 && n_keys <= VK_AGG_MAX_KEYS, "vk_agg_create: 0..8 key columns supported");
    VK_REQUIRE(n_funcs >= 0 && n_funcs <= VK_AGG_MAX_FUNCS, "vk_agg_create: 0..16 aggregate functions supported");
    VK_REQUIRE(n_keys == 0 || key_dtypes, "vk_agg_create: key_dtypes is NULL");
    VK_REQUIRE(n_funcs == 0 || (funcs && func_in_dtypes), "vk_agg_create: funcs is NULL");
    VkAgg* a = new VkAgg();
    a->n_keys = n_keys;
    a->n_funcs = n_funcs;
    a->expected_groups = expected_groups;
    for (int k = 0; k < n_keys; ++k) {
        if (!dtype_valid(key_dtypes[k])) { delete a; return fail(VK_ERR_ARG, "vk_agg_create: bad key dtype"); }
        a->key_dtypes[k] = key_dtypes[k];
    }
    for (int f = 0; f < n_funcs; ++f) {
        FuncSpec s{};
        s.func = funcs[f];
        s.in_dtype = func_in_dtypes[f];
        const int dt = s.in_dtype;
        if (s.func != VK_AGG_COUNT_STAR && !dtype_valid(dt)) { delete a; return fail(VK_ERR_ARG, "vk_agg_create: bad input dtype"); }
        s.in_unsigned = dtype_is_unsigned(dt);
        switch (s.func) {
            case VK_AGG_COUNT_STAR: s.acc = ACC_NONE; break;
            case VK_AGG_COUNT: s.acc = ACC_COUNT; break;
            case VK_AGG_MIN: case VK_AGG_MAX:
                s.acc = ACC_MAXORD;
                s.is_min = s.func == VK_AGG_MIN;
                s.ord = dtype_is_float(dt) ? ORD_F64 : (dtype_is_unsigned(dt) ? ORD_U64 : ORD_S64);
                break;
            case VK_AGG_SUM: case VK_AGG_AVG:
                // agg_func_factory.cpp:107-131 (SUM) / :176-211 (AVG)
                if (dt == VK_BOOL8) { delete a; return fail(VK_ERR_UNSUPPORTED, "Column data type is not supported by sum()/avg()."); }
                if (dtype_is_float(dt)) s.acc = ACC_SUM_F64;
                else if (dtype_size(dt) == 8) s.acc = ACC_SUM_I128;
                else s.acc = ACC_SUM_I64;
                break;
            default:
                delete a;
                return fail(VK_ERR_ARG, "vk_agg_create: unknown aggregate function");
        }
        a->specs[f] = s;
    }
    // kernel-selection knobs are sampled per object (vk_set_option / VINUM_B200_*)
    a->fast_log2s = (int) opt(OPT_AGG_LOG2S);
    a->fast_pf = (int) opt(OPT_AGG_PF);
    if (a->fast_log2s < 8) a->fast_log2s = 8;
    if (a->fast_log2s > 13) a->fast_log2s = 13;
    if (opt(OPT_AGG_WARPS) > 0) {
        a->fast_warps = (int) opt(OPT_AGG_WARPS);
        a->fast_warps_fixed = true;
    }
    if (a->fast_warps < 1) a->fast_warps = 1;
    if (a->fast_warps > FA_MAX_THREADS / 32) a->fast_warps = FA_MAX_THREADS / 32;
    a->fast_direct_policy = (int) opt(OPT_AGG_DIRECT);
    a->dict_policy = (int) opt(OPT_AGG_DICT);
    a->hot_policy = (int) opt(OPT_AGG_HOT);
    a->match_policy = (int) opt(OPT_AGG_ENTRY);
    if (a->match_policy < 0 || a->match_policy > 2) a->match_policy = 1;
    if (opt(OPT_AGG_NOFAST)) a->fast_disabled = true;
    a->wide_rows = (int) opt(OPT_AGG_WIDE);
    a->part_policy = (int) opt(OPT_AGG_PARTITION);
    a->learn_rows = (int64_t) 1 << (opt(OPT_AGG_LEARN_LOG2) < 10 ? 10 : (opt(OPT_AGG_LEARN_LOG2) > 30 ? 30 : opt(OPT_AGG_LEARN_LOG2)));
    const int rc = ctr_acquire(a);
    if (rc != VK_OK) { delete a; return rc; }
    *out = a;
    return VK_OK;
}

int vk_agg_destroy(VkAgg* a) {
    if (!a) return VK_OK;
    // Stream-ordered teardown: the table and the list go back to the pool behind the work
    // already queued on the object's stream.  The counter block is recycled by the next
    // aggregate (possibly on another stream), so that stream is drained first.
    cudaStream_t s = a->last_stream;
    if (a->multi_stream) {
        cudaDeviceSynchronize();
    } else if (a->stream_seen) {
        cudaStreamSynchronize(s);
    }
    prof_resolve(a);
    for (auto& sp : a->prof_free) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
    if (a->table_ready) free_table(&a->t, s);
    if (a->dict_dev) cudaFreeAsync(a->dict_dev, s);
    release_list(a, s);
    ctr_release(a);
    delete a;
    return VK_OK;
}

int vk_agg_last_path(VkAgg* a) { return a ? a->last_path : 0; }
double vk_agg_estimate_groups(double selected_rows, double groups) { return estimate_groups(selected_rows, groups); }

int vk_agg_profile(VkAgg* a, int enable) {
    VK_REQUIRE(a, "vk_agg_profile: agg is NULL");
    a->profile = enable != 0;
    return VK_OK;
}
int vk_agg_profile_read(VkAgg* a, int path, double* out_ms, int64_t* out_launches, int64_t* out_rows) {
    VK_REQUIRE(a && path >= 1 && path <= 3, "vk_agg_profile_read: bad argument");
    prof_resolve(a);
    if (out_ms) *out_ms = a->prof_ms[path];
    if (out_launches) *out_launches = a->prof_launches[path];
    if (out_rows) *out_rows = a->prof_rows[path];
    return VK_OK;
}

static int run_replay_until_empty(VkAgg* a, GenParams gp, int64_t chunk_rows, cudaStream_t s) {
    // Precondition: counters read back in a->h_ctr after the chunk's kernel.
    while (true) {
        if (a->h_ctr[CTR_LOST] != 0)
            return fail(VK_ERR_STATE, "vk_agg_update: internal error: replay list overflowed (rows lost)");
        const int64_t groups = (int64_t) a->h_ctr[CTR_GROUPS];
        const int64_t pending = (int64_t) a->h_ctr[CTR_LIST];
        a->groups_ub = groups;
        VK_DBG("replay: groups=%lld pending=%lld capacity=%lld", (long long) groups, (long long) pending,
               (long long) a->t.capacity);
        if (pending == 0) return VK_OK;
        // grow so that every pending row can become a new group, then replay the list
        int rc = grow_table(a, groups, pending + 1024, s);
        if (rc != VK_OK) return rc;
        // the list being replayed must not be appended to while it is read: with enough
        // free slots no row can fail, so `count` is only read.
        uint32_t* replay_rows = nullptr;
        VK_CUDA(cudaMallocAsync((void**) &replay_rows, (size_t) pending * sizeof(uint32_t), s));
        VK_CUDA(cudaMemcpyAsync(replay_rows, a->list, (size_t) pending * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        VK_CUDA(cudaMemcpyAsync(a->d_ctr + CTR_CURSOR + 1, a->d_ctr + CTR_LIST, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, s));
        VK_CUDA(cudaMemsetAsync(a->d_ctr + CTR_LIST, 0, sizeof(unsigned long long), s));
        gp.table = a->t;
        gp.replay = make_replay(a);
        gp.row_list = replay_rows;
        gp.row_list_count = a->d_ctr + CTR_CURSOR + 1;
        gp.n = chunk_rows;
        int64_t need = (pending + 255) / 256, capb = (int64_t) sm_count() * 8;
        agg_general_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(gp);
        VK_CHECK_LAUNCH("agg_general_kernel(replay)");
        VK_CUDA(cudaFreeAsync(replay_rows, s));
        rc = read_counters(a, s);
        if (rc != VK_OK) return rc;
    }
}

int vk_agg_update(VkAgg* a, const VkPredicate* pred, int64_t n_rows, const VkColumn* keys, const VkColumn* values,
                  VkStream stream) {
    VK_REQUIRE(a, "vk_agg_update: agg is NULL");
    VK_REQUIRE(n_rows >= 0, "vk_agg_update: negative n_rows");
    VK_REQUIRE(a->n_keys == 0 || keys, "vk_agg_update: keys is NULL");
    VK_REQUIRE(a->n_funcs == 0 || values, "vk_agg_update: values is NULL");
    cudaStream_t s = (cudaStream_t) stream;
    { const int irc = ctr_init(a, s); if (irc != VK_OK) return irc; }
    for (int k = 0; k < a->n_keys; ++k) {
        VK_REQUIRE(keys[k].dtype == a->key_dtypes[k], "vk_agg_update: key dtype changed between batches");
        VK_REQUIRE(keys[k].length == n_rows, "vk_agg_update: key column length != n_rows");
    }
    for (int f = 0; f < a->n_funcs; ++f) {
        if (a->specs[f].acc == ACC_NONE) continue;
        VK_REQUIRE(values[f].dtype == a->specs[f].in_dtype, "vk_agg_update: value dtype changed between batches");
        VK_REQUIRE(values[f].length == n_rows, "vk_agg_update: value column length != n_rows");
    }
    VkPredicate none{};
    none.kind = VK_PRED_NONE;
    if (!pred) pred = &none;

    // ---- table on first use ----
    if (!a->table_ready) {
        int64_t cap;
        if (a->n_keys == 0) cap = 1;
        else {
            int64_t guess = a->expected_groups > 0 ? a->expected_groups * 4 : (n_rows < (1 << 19) ? n_rows * 4 : (1 << 21));
            cap = pow2_ceil(guess < 4096 ? 4096 : guess);
        }
        int rc = alloc_table(a, &a->t, cap, s);
        if (rc != VK_OK) return rc;
        a->table_ready = true;
        a->groups_ub = 0;
        if (a->n_keys == 0) {
            // the single group exists even for empty input (COUNT(*) = 0, hash_agg_test.cpp:761-778)
            uint32_t two = 2;
            VK_CUDA(cudaMemcpyAsync(a->t.state, &two, sizeof(two), cudaMemcpyHostToDevice, s));
            unsigned long long one = 1;
            VK_CUDA(cudaMemcpyAsync(a->d_ctr + CTR_GROUPS, &one, sizeof(one), cudaMemcpyHostToDevice, s));
            a->groups_ub = 1;
            a->t.max_groups = 1;
        }
    }
    if (n_rows == 0) return VK_OK;

    // ---- un-grouped reduction ----
    if (a->n_keys == 0) {
        a->last_path = 3;
        int64_t need = (n_rows + 255) / 256, capb = (int64_t) sm_count() * 8;
        const unsigned grid = (unsigned) (need < capb ? need : capb);
        for (int f = -1; f < a->n_funcs; ++f) {
            if (f >= 0 && a->specs[f].acc == ACC_NONE) continue;
            OneParams op{};
            int pk;
            int rc = make_pred(*pred, n_rows, &op.pred, &pk);
            if (rc != VK_OK) return rc;
            op.fi = f;
            op.n = n_rows;
            op.table = a->t;
            if (f >= 0) {
                op.spec = a->specs[f];
                op.val = make_col(values[f]);
            }
            // row pairs per thread of agg_onegroup8_kernel (0: agg_onegroup_kernel).  Measured on one box
            // (profiles/r02_variants.md, COUNT(*), SUM(f64) WHERE f64 > c at 1e8 rows): 2 -> 0.52 ms, 4 -> 0.68 ms,
            // general kernel 1.30 ms
            const int fast1 = (int) opt(OPT_ONEGROUP_FAST);
            bool raw8 = pk == PK_NONE || pk == PK_F64_VEC || pk == PK_I64_VEC;
            if (f >= 0 && op.spec.acc != ACC_COUNT) {
                const int dt = op.val.dtype;
                const bool f64_acc = op.spec.acc == ACC_SUM_F64 || (op.spec.acc == ACC_MAXORD && op.spec.ord == ORD_F64);
                raw8 = raw8 && (f64_acc ? dt == VK_F64 : (dt == VK_I64 || dt == VK_U64)) && op.spec.acc != ACC_SUM_I64 &&
                       (reinterpret_cast<uintptr_t>(op.val.data) & 15) == 0;
            }
            if (f >= 0) raw8 = raw8 && op.val.validity == nullptr;
            if (fast1 >= 2 && raw8) {
                int64_t need8 = (n_rows / 2 + 256 * 4 - 1) / (256 * 4);
                const unsigned g8 = (unsigned) (need8 < 1 ? 1 : (need8 < capb ? need8 : capb));
#define VK_ONE8_GO(PK)                                                               \
                do {                                                                 \
                    if (fast1 >= 4) agg_onegroup8_kernel<PK, 4><<<g8, 256, 0, s>>>(op); \
                    else agg_onegroup8_kernel<PK, 2><<<g8, 256, 0, s>>>(op);         \
                } while (0)
                if (pk == PK_NONE) VK_ONE8_GO(PK_NONE);
                else if (pk == PK_F64_VEC) VK_ONE8_GO(PK_F64_VEC);
                else VK_ONE8_GO(PK_I64_VEC);
#undef VK_ONE8_GO
                VK_CHECK_LAUNCH("agg_onegroup8_kernel");
                continue;
            }
            agg_onegroup_kernel<<<grid, 256, 0, s>>>(op);
            VK_CHECK_LAUNCH("agg_onegroup_kernel");
        }
        return VK_OK;
    }

    // ---- path selection ----
    FastPlan plan;
    const bool plan_ok = a->n_keys == 1 && build_fast_plan(a, keys, values, &plan);
    bool fast = plan_ok && !a->fast_disabled;
    // the partitioned plan (scatter + update per table slice) takes every single-key aggregate
    const bool part_ok = a->n_keys == 1 && a->part_policy != 0;
    Pred dpred;
    int pk;
    int rc = make_pred(*pred, n_rows, &dpred, &pk);
    if (rc != VK_OK) return rc;

    const int sms = sm_count();
    const int log2s = a->fast_log2s;
    const bool lean = fast && fast_pk_is_lean(pk) && plan.mode != FM_RUNTIME;
    int warps = a->fast_warps, gmax = 0;
    bool direct = false;
    uint64_t direct_base = 0;
    size_t smem = 0;
    int pf_dist = 0;
    auto configure_fast = [&](int64_t groups_hint) {
        direct = false;
        // Two measured regimes (profiles/): when the predicate drops a good share of the rows
        // the kernel is HBM-bound and 8 warps + an L2 bulk prefetch 6 tiles ahead stream best;
        // when (nearly) every row reaches the tables it is bound by the shared-memory pipe
        // and wants all 12 warps and no prefetch.
        const bool selective = pk != PK_NONE && a->fast_rows_seen > 0 && a->fast_selectivity <= 0.7;
        pf_dist = a->fast_pf >= 0 ? a->fast_pf : (selective ? 6 : 0);
        const int w_auto = selective && FA_MAX_THREADS / 32 >= 8 ? 8 : FA_MAX_THREADS / 32;
        const int w_hi = a->fast_warps_fixed ? a->fast_warps : w_auto;
        const int w_lo = a->fast_warps_fixed ? a->fast_warps : 2;
        auto next_w = [](int w) { return w > 8 ? w - 2 : w >> 1; };  // 12, 10, 8, 4, 2
        // direct group ids (key - base): the key range seen so far fits the warp-private tables
        if (lean && a->fast_direct_policy && a->direct_known && a->direct_ok) {
            for (int w = w_hi; w >= w_lo; w = next_w(w)) {
                const int g = fast_gmax(log2s, true, plan.nw, w);
                if (g >= 16 && a->direct_span < (uint64_t) g) {
                    direct = true;
                    warps = w;
                    gmax = g;
                    // centre the observed range in the window so that unseen neighbours still map
                    direct_base = a->direct_min - ((uint64_t) g - 1 - a->direct_span) / 2;
                    break;
                }
            }
        }
        if (!direct && a->dict_ready) {
            // read-only dictionary: the dense ids are exactly 0 .. dict_n - 1.  The lookup makes the kernel
            // instruction / latency bound rather than HBM bound (8 warps: 42 % issue utilisation, 0.5 eligible
            // warps per cycle, profiles/r02_agg_fast_dict_ncu_full.md): as many warps as the tables leave room for
            warps = a->fast_warps_fixed ? a->fast_warps : FA_MAX_THREADS / 32;
            const int need = (a->dict_n + 15) & ~15;
            while (warps > w_lo && need > fast_gmax(log2s, false, plan.nw, warps)) warps = next_w(warps);
            gmax = need < 16 ? 16 : need;
            if (gmax > fast_gmax(log2s, false, plan.nw, warps)) fast = false;
        } else if (!direct) {
            // fewer warps per CTA leave more shared memory per warp-private table
            warps = w_hi;
            while (warps > w_lo && groups_hint + groups_hint / 32 + 8 > fast_gmax(log2s, false, plan.nw, warps)) warps = next_w(warps);
            gmax = fast_gmax(log2s, false, plan.nw, warps);
            if (gmax < 16 || groups_hint > gmax) fast = false;
        }
        smem = fast_smem_bytes(log2s, direct, gmax, plan.nw, warps);
    };
    if (fast) configure_fast(a->fast_groups_seen);
    const int fast_grid_max = sms;  // one persistent CTA per SM

    // ---- chunk loop: never more rows in flight than (free slots + replay capacity) ----
    if (a->list_max == 0) {
        // once per device, not per query: cudaMemGetInfo walks the driver's allocation state and
        // has been seen to take milliseconds with tens of GB resident
        static uint64_t g_list_max[64] = {0};
        const int dev = a->device >= 0 && a->device < 64 ? a->device : 0;
        if (g_list_max[dev] == 0) {
            // large enough that a 1e9-row batch is ONE launch (the list is only touched by rows
            // that could not be inserted; the memory is kept per device and never written otherwise)
            int l2 = (int) opt(OPT_LIST_LOG2);
            if (l2 < 16) l2 = 16;
            if (l2 > 31) l2 = 31;
            uint64_t lm = (uint64_t) 1 << l2;
            size_t fr = 0, tot = 0;
            if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
                while (lm * sizeof(uint32_t) > fr / 16 && lm > (1u << 16)) lm >>= 1;
            }
            g_list_max[dev] = lm;
        }
        a->list_max = g_list_max[dev];
    }
    const uint64_t list_max = a->list_max;
    int64_t pos = 0;
    while (pos < n_rows) {
        const int64_t remaining = n_rows - pos;
        const int64_t flush_reserve = fast ? (int64_t) fast_grid_max * gmax : 0;
        int64_t free_slots = a->t.max_groups - a->groups_ub - flush_reserve;
        if (free_slots < 1024) {
            // bound is pessimistic: refresh it, then grow if it is real
            rc = read_counters(a, s);
            if (rc != VK_OK) return rc;
            a->groups_ub = (int64_t) a->h_ctr[CTR_GROUPS];
            free_slots = a->t.max_groups - a->groups_ub - flush_reserve;
            if (free_slots < 1024) {
                rc = grow_table(a, a->groups_ub, a->groups_ub + flush_reserve + 4096, s);
                if (rc != VK_OK) return rc;
                free_slots = a->t.max_groups - a->groups_ub - flush_reserve;
            }
        }
        int64_t chunk = remaining;
        bool may_fail = false;
        // the first fast chunk is small: it tells the cardinality, the key range and the
        // selectivity before the geometry is fixed (and it needs no replay list)
        if (fast && a->fast_rows_seen == 0) {
            const int64_t learn = free_slots < a->learn_rows ? free_slots : a->learn_rows;
            if (chunk > learn) chunk = learn;
        }
        // global-table path, groups still rising: no more than 8 x the rows seen so far in one launch, so that
        // a table that turns out too small is found out (and sized from an estimate) before most rows arrive
        const bool sizing = !fast && a->sizing;
        if (sizing) {
            const int64_t lim = a->rows_seen < ((int64_t) 1 << 17) ? ((int64_t) 1 << 20) : 8 * a->rows_seen;
            if (chunk > lim) chunk = lim;
        }
        // skewed keys on the global-table kernels: the variant that merges a warp's updates of one slot (option AGG_HOT)
        const bool hot_keys = a->hot_policy == 2 || (a->hot_policy == 1 && a->hot_share >= 0.3);
        // table beyond the L2: scatter into buckets (= slices of the table) + update slice by slice instead of at random
        // (the part of the table that is touched counts, not what is allocated: groups at the load limit)
        // (and not with one key holding a third of the rows: they would all land in ONE bucket, overflow it and come back
        // through the replay list -- measured 245 ms against 17 for the global-table kernel with the hot-slot merge)
        const bool use_part = !fast && part_ok &&
                              (a->part_policy == 2 ||
                               (!hot_keys && a->est_groups * (double) slot_bytes(a) / kMaxLoad > (double) PART_MIN_TABLE_BYTES));
        if (use_part && chunk > PART_MAX_CHUNK) chunk = PART_MAX_CHUNK;
        if (use_part) {
            // any row may be deferred (its bucket full, its group not creatable): the list must hold a whole chunk
            uint64_t want = (uint64_t) chunk;
            if (want > list_max) want = list_max;
            rc = ensure_list(a, want, s);
            if (rc != VK_OK) return rc;
            if (chunk > (int64_t) a->list_cap) chunk = (int64_t) a->list_cap;
            may_fail = true;
        } else if (chunk > free_slots) {
            uint64_t want = (uint64_t) (chunk - free_slots);
            if (want > list_max) want = list_max;
            rc = ensure_list(a, want, s);
            if (rc != VK_OK) return rc;
            if (chunk > free_slots + (int64_t) a->list_cap) chunk = free_slots + (int64_t) a->list_cap;
            may_fail = true;
        }
        if (chunk > (int64_t) 0xfffff000LL) chunk = (int64_t) 0xfffff000LL;  // row ids in the list are 32-bit
        if (chunk < remaining) {
            // chunk starts stay pair-aligned, and fast chunks are whole tiles (only the last one has a tail)
            const int64_t unit = fast ? (int64_t) warps * 32 * FA_R : 4096;
            if (chunk >= unit) chunk -= chunk % unit;
            else chunk &= ~(int64_t) 1;
        }
        if (chunk <= 0) return fail(VK_ERR_STATE, "vk_agg_update: internal error: empty chunk");
        VK_CUDA(cudaMemsetAsync(a->d_ctr + CTR_LIST, 0, 3 * sizeof(unsigned long long), s));  // list, lost, spilled
        VK_DBG("chunk pos=%lld rows=%lld fast=%d may_fail=%d free_slots=%lld capacity=%lld groups_ub=%lld list_cap=%llu",
               (long long) pos, (long long) chunk, (int) fast, (int) may_fail, (long long) free_slots,
               (long long) a->t.capacity, (long long) a->groups_ub, (unsigned long long) a->list_cap);

        VkPredicate cpred = *pred;
        VkExprCompare cexpr;
        if (cpred.kind == VK_PRED_MASK) cpred.mask += pos;
        if (cpred.kind == VK_PRED_CMP) cpred.column = slice_col(cpred.column, pos, chunk);
        if (cpred.kind == VK_PRED_EXPR && cpred.expr != nullptr) {
            cexpr = *cpred.expr;
            for (VkExprChain* ch : {&cexpr.lhs, &cexpr.rhs})
                for (int k = 0; k < ch->n_terms && k < VK_EXPR_MAX_TERMS; ++k)
                    if (ch->terms[k].is_column) ch->terms[k].column = slice_col(ch->terms[k].column, pos, chunk);
            cpred.expr = &cexpr;
        }
        rc = make_pred(cpred, chunk, &dpred, &pk);
        if (rc != VK_OK) return rc;

        bool fused_range = false;
        GenParams gp{};
        gp.pred = dpred;
        gp.pk = pk;
        gp.n_keys = a->n_keys;
        for (int k = 0; k < a->n_keys; ++k) gp.keys[k] = make_col(slice_col(keys[k], pos, chunk));
        gp.n_funcs = a->n_funcs;
        for (int f = 0; f < a->n_funcs; ++f) {
            gp.specs[f] = a->specs[f];
            if (a->specs[f].acc != ACC_NONE) gp.vals[f] = make_col(slice_col(values[f], pos, chunk));
        }
        gp.n = chunk;
        gp.table = a->t;
        gp.replay = make_replay(a);

        if (fast) {
            a->last_path = 1;
            FastParams fp{};
            fp.pred = dpred;
            fp.key = gp.keys[0];
            fp.key_mode = plan.key_mode;
            fp.n_cols = plan.n_cols;
            for (int v = 0; v < plan.n_cols; ++v) {
                fp.col[v] = make_col(slice_col(plan.cols[v], pos, chunk));
                fp.col_mode[v] = plan.col_mode[v];
            }
            fp.n_cells = plan.n_cells;
            for (int c = 0; c < plan.n_cells; ++c) fp.cell[c] = plan.cells[c];
            const int threads = warps * 32;
            const int64_t tile_rows = (int64_t) threads * FA_R;
            fp.n = chunk;
            fp.num_tiles = chunk / tile_rows;  // complete tiles; the ragged tail goes to the general kernel below
            fp.log2s = log2s;
            fp.gmax = gmax;
            fp.direct_base = direct_base;
            fp.row_limit = a->t.max_groups - flush_reserve;
            fp.pf_dist = warps >= 5 ? pf_dist : 0;  // one prefetching lane per column: needs NV + 2 warps
            fp.table = a->t;
            fp.replay = gp.replay;
            if (!direct && a->dict_ready) {
                fp.dict_keys = reinterpret_cast<const uint64_t*>(a->dict_dev);
                fp.dict_gids = reinterpret_cast<const uint16_t*>(a->dict_dev + ((size_t) 8 << log2s));
                fp.seed_a = a->seed_a;
            }
            FastLaunch fl;
            fl.pk = pk;
            fl.nw = plan.nw;
            fl.mode = lean ? plan.mode : FM_RUNTIME;
            fl.direct = direct;
            fl.sumf64 = plan.sumf64;
            // COUNT + SUM(f64) entry layout (option AGG_ENTRY): split entries move 11 % fewer shared-memory
            // wavefronts and win where that pipe binds (C3: 3.89 ms against 4.17), 16-byte entries win by 2 %
            // where HBM binds (selective predicate); 1 = choose by the selectivity the learning launch measured
            {
                const bool pipe_bound = !(pk != PK_NONE && a->fast_rows_seen > 0 && a->fast_selectivity <= 0.7);
                const int want = a->match_policy == 1 ? (pipe_bound ? 2 : 0) : a->match_policy;
                fl.variant = (lean && plan.sumf64 && plan.nw == 2) ? want : 0;
                // the kernel with the hot-group step only where a key holds a large share of the rows: for
                // uniform keys its two extra votes per row slot cost the HBM-bound kernel 50 %
                fl.hot = lean && plan.sumf64 && plan.nw == 2 && (a->hot_policy == 2 || (a->hot_policy == 1 && a->hot_share >= 0.3));
            }
            fl.grid = fp.num_tiles < fast_grid_max ? (int) fp.num_tiles : fast_grid_max;
            fl.threads = threads;
            fl.smem = smem;
            VK_DBG("fast launch: pk=%d mode=%d direct=%d base=%lld gmax=%d warps=%d nw=%d smem=%zu", pk, fl.mode, (int) direct,
                   (long long) direct_base, gmax, warps, plan.nw, smem);
            const int64_t fast_rows = fp.num_tiles * tile_rows;
            // the learning launch on an empty table also reports its key range (saves a scan + a sync)
            fused_range = lean && a->fast_direct_policy && !a->direct_known && a->groups_ub == 0 && fast_rows == chunk;
            if (fused_range) {
                const long long init[2] = {INT64_MAX, INT64_MIN};
                VK_CUDA(cudaMemcpyAsync(a->d_ctr + CTR_RANGE, init, sizeof(init), cudaMemcpyHostToDevice, s));
                fp.key_range = reinterpret_cast<long long*>(a->d_ctr + CTR_RANGE);
            }
            if (fp.num_tiles > 0) {
                const int span = prof_begin(a, s, fast_rows, 1);
                rc = launch_fast(fp, fl, s);
                prof_end(a, s, span);
                if (rc != VK_OK) return rc;
            }
            if (fast_rows < chunk) {
                GenParams tp = gp;
                tp.row_begin = fast_rows;
                int64_t need = (chunk - fast_rows + 255) / 256;
                agg_general_kernel<<<(unsigned) need, 256, 0, s>>>(tp);
                VK_CHECK_LAUNCH("agg_general_kernel(tail)");
            }
            a->groups_ub += chunk;
        } else if (use_part) {
            a->last_path = 4;
            PartParams pp{};
            pp.g = gp;
            // buckets = slices of the table of ~4 MB (all its arrays), so that the slice being updated stays in the L2
            int log2cap = 0;
            while (((int64_t) 1 << log2cap) < a->t.capacity) ++log2cap;
            pp.log2p = 0;
            while (pp.log2p < 12 && pp.log2p + 8 < log2cap && (a->t.capacity >> pp.log2p) * slot_bytes(a) > PART_BUCKET_BYTES) ++pp.log2p;
            pp.shift = log2cap - pp.log2p;
            const int64_t n_buckets = (int64_t) 1 << pp.log2p;
            // records per bucket: from the share of the rows selected so far, 15 % over
            const uint64_t selected_so_far = a->h_ctr[CTR_SELECTED] + a->fast_spilled;
            double sel = a->sized_selected > 0 && a->rows_seen > 0 ? (double) selected_so_far / (double) a->rows_seen : 1.0;
            sel = sel * 1.15 + 0.01;
            if (sel > 1.0 || pk == PK_NONE) sel = 1.0;
            const double mean = (double) chunk * sel / (double) n_buckets;
            pp.cap = (uint32_t) (mean * 1.1 + 6.0 * sqrt(mean) + 64.0);   // the hash spreads keys evenly: Poisson tails + 10 %
            const size_t recs = (size_t) n_buckets * pp.cap;
            uint8_t* scratch = nullptr;
            const size_t cursor_bytes = (size_t) n_buckets * 128 + 128;   // + the tile queue of the update pass
            VK_CUDA(cudaMallocAsync((void**) &scratch, cursor_bytes + recs * 16, s));
            pp.cursor = reinterpret_cast<unsigned int*>(scratch);
            pp.recs = reinterpret_cast<ulonglong2*>(scratch + cursor_bytes);
            VK_CUDA(cudaMemsetAsync(pp.cursor, 0, cursor_bytes, s));
            VK_DBG("partitioned: buckets=%lld cap=%u table=%.1f MB scratch=%.1f MB", (long long) n_buckets, pp.cap,
                   (double) (a->t.capacity * slot_bytes(a)) / 1e6, (double) (cursor_bytes + recs * 16) / 1e6);
            const int span = prof_begin(a, s, chunk, 2);
            rc = launch_part(pp, hot_keys, sms, s);
            prof_end(a, s, span);
            VK_CUDA(cudaFreeAsync(scratch, s));
            if (rc != VK_OK) return rc;
            a->groups_ub += chunk;
        } else {
            a->last_path = 2;
            int64_t need = (chunk + 255) / 256, capb = (int64_t) sms * 8;
            const int span = prof_begin(a, s, chunk, 2);
            // single-key tables: several rows per thread, the loads of each stage issued together
            if (a->t.single && a->wide_rows != 0) rc = launch_wide(gp, hot_keys, sms, s);
            else {
                agg_general_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(gp);
                VK_CHECK_LAUNCH("agg_general_kernel");
            }
            prof_end(a, s, span);
            if (rc != VK_OK) return rc;
            a->groups_ub += chunk;
        }

        a->rows_seen += chunk;
        if (may_fail || sizing || (fast && a->fast_rows_seen < ((uint64_t) 1 << 20))) {
            // rows may have been deferred (or we are still learning the cardinality)
            if (fast && a->hot_policy == 1 && plan.sumf64) {
                // key skew so far: the largest group's share of the selected rows (one scan of the table)
                VK_CUDA(cudaMemsetAsync(a->d_ctr + CTR_MAXCOUNT, 0, sizeof(unsigned long long), s));
                int64_t need = (gt_total_slots(a->t) + 255) / 256, capb = (int64_t) sms * 8;
                agg_max_count_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(a->t, a->d_ctr + CTR_MAXCOUNT);
                VK_CHECK_LAUNCH("agg_max_count_kernel");
            }
            rc = read_counters(a, s);
            if (rc != VK_OK) return rc;
            if (fast) {
                const uint64_t spill_now = a->h_ctr[CTR_SPILL];
                a->fast_rows_seen += (uint64_t) chunk;
                a->fast_spilled += spill_now;
                a->fast_selectivity = (double) (a->h_ctr[CTR_SELECTED] + a->fast_spilled) / (double) a->fast_rows_seen;
                a->fast_groups_seen = (int64_t) a->h_ctr[CTR_GROUPS];
                if (a->hot_policy == 1 && plan.sumf64) {
                    const double selected = (double) (a->h_ctr[CTR_SELECTED] + a->fast_spilled);
                    a->hot_share = selected > 0 ? (double) a->h_ctr[CTR_MAXCOUNT] / selected : 0.0;
                    VK_DBG("largest group holds %.1f %% of the rows", 100.0 * a->hot_share);
                }
                if (chunk >= 65536 && spill_now * 4 > (uint64_t) chunk) {
                    // this configuration thrashes: most rows fell through to the global path
                    if (direct) a->direct_ok = false;   // the key range moved: back to hash mode
                    else if (a->dict_ready) a->dict_ready = false;   // stale dictionary: rebuilt below from the table as it is now
                    else a->fast_disabled = true;       // cardinality too high for shared memory
                } else if (lean && a->fast_direct_policy && !a->direct_known &&
                           a->fast_groups_seen <= fast_gmax(log2s, true, plan.nw, 2)) {
                    if (fused_range && spill_now == 0) {
                        apply_key_range(a);
                    } else {
                        rc = measure_key_range(a, s);
                        if (rc != VK_OK) return rc;
                    }
                }
            }
            rc = run_replay_until_empty(a, gp, chunk, s);
            if (rc != VK_OK) return rc;
            if (a->fast_disabled) fast = false;
            // no dense key range: look the keys up in a read-only dictionary from now on; a dictionary that
            // most rows miss (the keys moved on) is rebuilt from the table as it is now
            const bool want_dict = fast && a->dict_policy && !a->dict_failed &&
                                   !(a->fast_direct_policy && a->direct_known && a->direct_ok && lean);
            if (want_dict && !a->dict_ready) {
                rc = build_dict(a, (int64_t) a->h_ctr[CTR_GROUPS], log2s, s);
                if (rc != VK_OK) return rc;
            }
            if (fast) configure_fast(a->fast_groups_seen);
            if (!fast && a->sizing) {
                // counters as of the end of this chunk (the replay loop may have moved them)
                rc = read_counters(a, s);
                if (rc != VK_OK) return rc;
                const int64_t groups = (int64_t) a->h_ctr[CTR_GROUPS];
                const uint64_t selected = a->h_ctr[CTR_SELECTED] + a->fast_spilled;
                const int64_t new_groups = groups - a->sized_groups;
                const uint64_t new_selected = selected - a->sized_selected;
                a->groups_ub = groups;
                if (sizing && (uint64_t) new_groups * 64 < new_selected) {
                    a->sizing = false;   // fewer than 1 row in 64 opened a group: the table has seen most keys
                } else if (selected > 0) {
                    // rows still to come: known within this update; after its last chunk the caller may or may not send more
                    // batches (BaseAggregate::Next is called per batch), so room is made for at most 4 x the groups so far
                    const bool last = pos + chunk >= n_rows;
                    const double left = last ? 3.0 * (double) groups
                                             : (double) (n_rows - pos - chunk) * ((double) selected / (double) a->rows_seen);
                    double est = estimate_groups((double) selected, (double) groups);
                    if (est <= 0 || est > (double) groups + left) est = (double) groups + left;   // every row left a new group
                    a->est_groups = est;
                    // 2 % over the estimate (its own error is a few tenths of a percent; a table one doubling larger
                    // than needed costs L2 residency: 1e6 groups fit the 2 Mi-slot table, 48 MB, and must stay there)
                    int64_t want = (int64_t) (est * 1.02);
                    // never more than a third of the free memory in one step
                    if (want > a->t.max_groups && want * 2 >= ((int64_t) 1 << 26)) {
                        size_t fr = 0, tot = 0;
                        if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) {
                            const int64_t fit = (int64_t) (fr / 3) / slot_bytes(a) / 2;   // groups at the load limit
                            if (want > fit) want = fit;
                        }
                    }
                    VK_DBG("sizing: groups=%lld selected=%llu estimate=%.0f want=%lld max_groups=%lld", (long long) groups,
                           (unsigned long long) selected, est, (long long) want, (long long) a->t.max_groups);
                    if (want > a->t.max_groups) {
                        rc = grow_table(a, groups, want - groups, s);
                        if (rc != VK_OK) return rc;
                    }
                }
                a->sized_groups = groups;
                a->sized_selected = selected;
            }
        }
        pos += chunk;
    }
    return VK_OK;
}

int vk_agg_num_groups(VkAgg* a, int64_t* out_groups, VkStream stream) {
    VK_REQUIRE(a && out_groups, "vk_agg_num_groups: NULL argument");
    if (!a->table_ready) {
        *out_groups = a->n_keys == 0 ? 1 : 0;
        return VK_OK;
    }
    int rc = ctr_init(a, (cudaStream_t) stream);
    if (rc != VK_OK) return rc;
    rc = read_counters(a, (cudaStream_t) stream);
    if (rc != VK_OK) return rc;
    a->groups_ub = (int64_t) a->h_ctr[CTR_GROUPS];
    *out_groups = a->groups_ub;
    return VK_OK;
}

int vk_agg_result(VkAgg* a, int64_t num_groups, uint64_t* const* out_keys, uint8_t* const* out_key_valid,
                  uint64_t* out_count_star, uint64_t* const* out_vals_lo, uint64_t* const* out_vals_hi,
                  uint8_t* const* out_vals_valid, VkStream stream) {
    VK_REQUIRE(a, "vk_agg_result: agg is NULL");
    cudaStream_t s = (cudaStream_t) stream;
    { const int irc = ctr_init(a, s); if (irc != VK_OK) return irc; }
    if (num_groups == 0) return VK_OK;
    VK_REQUIRE(out_count_star, "vk_agg_result: out_count_star is NULL");
    if (!a->table_ready) {
        // OneGroup result() without any batch: a single all-NULL / zero-count row
        VK_REQUIRE(a->n_keys == 0 && num_groups == 1, "vk_agg_result: no batches were aggregated");
        VK_CUDA(cudaMemsetAsync(out_count_star, 0, 8, s));
        for (int f = 0; f < a->n_funcs; ++f) {
            VK_CUDA(cudaMemsetAsync(out_vals_lo[f], 0, 8, s));
            if (out_vals_hi && out_vals_hi[f]) VK_CUDA(cudaMemsetAsync(out_vals_hi[f], 0, 8, s));
            int is_count = a->specs[f].func == VK_AGG_COUNT_STAR || a->specs[f].func == VK_AGG_COUNT;
            VK_CUDA(cudaMemsetAsync(out_vals_valid[f], is_count ? 1 : 0, 1, s));
        }
        return VK_OK;
    }
    FinalParams fp{};
    fp.table = a->t;
    fp.num_groups = num_groups;
    fp.null_last = a->n_keys == 1;
    fp.cursor = a->d_ctr + CTR_CURSOR;
    fp.capacity = num_groups;
    fp.header = nullptr;
    VK_CUDA(cudaMemsetAsync(fp.cursor, 0, sizeof(unsigned long long), s));
    for (int k = 0; k < a->n_keys; ++k) {
        VK_REQUIRE(out_keys && out_keys[k] && out_key_valid && out_key_valid[k], "vk_agg_result: NULL key output");
        fp.out_keys[k] = out_keys[k];
        fp.out_key_valid[k] = out_key_valid[k];
    }
    fp.out_count_star = out_count_star;
    for (int f = 0; f < a->n_funcs; ++f) {
        VK_REQUIRE(out_vals_lo && out_vals_lo[f] && out_vals_valid && out_vals_valid[f], "vk_agg_result: NULL value output");
        fp.specs[f] = a->specs[f];
        fp.out_lo[f] = out_vals_lo[f];
        fp.out_hi[f] = out_vals_hi ? out_vals_hi[f] : nullptr;
        fp.out_valid[f] = out_vals_valid[f];
    }
    int64_t need = (gt_total_slots(a->t) + 255) / 256, capb = (int64_t) sm_count() * 8;
    agg_finalize_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(fp);
    VK_CHECK_LAUNCH("agg_finalize_kernel");
    return VK_OK;
}

int vk_agg_record_words(VkAgg* a, int* out_words) {
    VK_REQUIRE(a && out_words, "vk_agg_record_words: NULL argument");
    *out_words = a->n_keys + 2 + 3 * a->n_funcs;
    return VK_OK;
}

static int exch_params(VkAgg* a, ExchParams* p) {
    memset(p, 0, sizeof(*p));
    p->table = a->t;
    for (int f = 0; f < a->n_funcs; ++f) p->specs[f] = a->specs[f];
    p->words = a->n_keys + 2 + 3 * a->n_funcs;
    return VK_OK;
}

int vk_agg_partition_counts(VkAgg* a, int n_ranks, int64_t* out_counts_dev, VkStream stream) {
    VK_REQUIRE(a && out_counts_dev && n_ranks >= 1, "vk_agg_partition_counts: bad argument");
    cudaStream_t s = (cudaStream_t) stream;
    { const int irc = ctr_init(a, s); if (irc != VK_OK) return irc; }
    VK_CUDA(cudaMemsetAsync(out_counts_dev, 0, sizeof(int64_t) * n_ranks, s));
    if (!a->table_ready) return VK_OK;
    ExchParams p;
    exch_params(a, &p);
    p.n_ranks = n_ranks;
    p.counts = reinterpret_cast<long long*>(out_counts_dev);
    int64_t need = (gt_total_slots(a->t) + 255) / 256, capb = (int64_t) sm_count() * 8;
    agg_partition_count_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(p);
    VK_CHECK_LAUNCH("agg_partition_count_kernel");
    return VK_OK;
}

int vk_agg_export_partials(VkAgg* a, int n_ranks, const int64_t* offsets_dev, uint64_t* out_records, VkStream stream) {
    VK_REQUIRE(a && offsets_dev && n_ranks >= 1, "vk_agg_export_partials: bad argument");
    if (!a->table_ready) return VK_OK;
    VK_REQUIRE(out_records, "vk_agg_export_partials: out_records is NULL");
    cudaStream_t s = (cudaStream_t) stream;
    { const int irc = ctr_init(a, s); if (irc != VK_OK) return irc; }
    ExchParams p;
    exch_params(a, &p);
    p.n_ranks = n_ranks;
    p.offsets = reinterpret_cast<const long long*>(offsets_dev);
    p.out = out_records;
    unsigned long long* cursors = nullptr;
    VK_CUDA(cudaMallocAsync((void**) &cursors, sizeof(unsigned long long) * n_ranks, s));
    VK_CUDA(cudaMemsetAsync(cursors, 0, sizeof(unsigned long long) * n_ranks, s));
    p.cursors = cursors;
    int64_t need = (gt_total_slots(a->t) + 255) / 256, capb = (int64_t) sm_count() * 8;
    agg_export_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(p);
    VK_CHECK_LAUNCH("agg_export_kernel");
    VK_CUDA(cudaFreeAsync(cursors, s));
    return VK_OK;
}

int vk_agg_merge_partials(VkAgg* a, const uint64_t* records, int64_t n_records, VkStream stream) {
    VK_REQUIRE(a && n_records >= 0, "vk_agg_merge_partials: bad argument");
    VK_REQUIRE(a->n_keys > 0, "vk_agg_merge_partials: un-grouped aggregates merge on the host");
    if (n_records == 0) return VK_OK;
    VK_REQUIRE(records, "vk_agg_merge_partials: records is NULL");
    cudaStream_t s = (cudaStream_t) stream;
    { const int irc = ctr_init(a, s); if (irc != VK_OK) return irc; }
    if (!a->table_ready) {
        int rc = alloc_table(a, &a->t, pow2_ceil(n_records * 4 < 4096 ? 4096 : n_records * 4), s);
        if (rc != VK_OK) return rc;
        a->table_ready = true;
        a->groups_ub = 0;
    }
    // make room for every record becoming a new group: merging can then never fail
    int rc = read_counters(a, s);
    if (rc != VK_OK) return rc;
    a->groups_ub = (int64_t) a->h_ctr[CTR_GROUPS];
    rc = grow_table(a, a->groups_ub, n_records + 1024, s);
    if (rc != VK_OK) return rc;
    rc = ensure_list(a, 1024, s);
    if (rc != VK_OK) return rc;
    VK_CUDA(cudaMemsetAsync(a->d_ctr + CTR_LIST, 0, 3 * sizeof(unsigned long long), s));
    ExchParams p;
    exch_params(a, &p);
    p.in = records;
    p.n_in = n_records;
    p.replay = make_replay(a);
    int64_t need = (n_records + 255) / 256, capb = (int64_t) sm_count() * 8;
    agg_merge_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(p);
    VK_CHECK_LAUNCH("agg_merge_kernel");
    a->groups_ub += n_records;
    return VK_OK;
}

// ------------------------------------------------------------ packed result ----
// Finalise into ONE caller-owned device block so that the host needs a single copy and a single
// synchronisation to get the groups (the separate vk_agg_num_groups + vk_agg_result + two copies cost
// three stream round trips: 0.2 ms of a 4.2 ms query).  Block (u64 words), C = capacity:
//   [0] groups in the table  [1] reserved
//   keys[n_keys][C] | count_star[C] | lo[n_funcs][C] | hi[n_funcs][C] | bytes: key_valid[n_keys][C] valid[n_funcs][C]
// Word 0 may exceed C: the caller then retries with a larger block.
uint64_t vk_agg_result_packed_bytes(int n_keys, int n_funcs, int64_t capacity) {
    if (capacity < 1) capacity = 1;
    const uint64_t words = 2 + (uint64_t) (n_keys + 1 + 2 * n_funcs) * (uint64_t) capacity;
    const uint64_t bytes = (uint64_t) (n_keys + n_funcs) * (uint64_t) capacity;
    return words * 8 + ((bytes + 7) & ~(uint64_t) 7);
}

int vk_agg_result_packed(VkAgg* a, int64_t capacity, void* out_block, VkStream stream) {
    VK_REQUIRE(a && out_block && capacity >= 1, "vk_agg_result_packed: bad argument");
    VK_REQUIRE(a->n_keys >= 1, "vk_agg_result_packed: group-by aggregates only");
    cudaStream_t s = (cudaStream_t) stream;
    { const int irc = ctr_init(a, s); if (irc != VK_OK) return irc; }
    uint64_t* blk = reinterpret_cast<uint64_t*>(out_block);
    if (!a->table_ready) {   // no batch was aggregated: zero groups
        VK_CUDA(cudaMemsetAsync(blk, 0, 16, s));
        return VK_OK;
    }
    FinalParams fp{};
    fp.table = a->t;
    fp.num_groups = -1;
    fp.null_last = a->n_keys == 1;
    fp.cursor = a->d_ctr + CTR_CURSOR;
    fp.capacity = capacity;
    fp.header = blk;
    VK_CUDA(cudaMemsetAsync(fp.cursor, 0, sizeof(unsigned long long), s));
    uint64_t* w = blk + 2;
    for (int k = 0; k < a->n_keys; ++k) { fp.out_keys[k] = w; w += capacity; }
    fp.out_count_star = w; w += capacity;
    for (int f = 0; f < a->n_funcs; ++f) { fp.out_lo[f] = w; w += capacity; }
    for (int f = 0; f < a->n_funcs; ++f) { fp.out_hi[f] = w; w += capacity; }
    uint8_t* b = reinterpret_cast<uint8_t*>(w);
    for (int k = 0; k < a->n_keys; ++k) { fp.out_key_valid[k] = b; b += capacity; }
    for (int f = 0; f < a->n_funcs; ++f) { fp.specs[f] = a->specs[f]; fp.out_valid[f] = b; b += capacity; }
    int64_t need = (gt_total_slots(a->t) + 255) / 256, capb = (int64_t) sm_count() * 8;
    agg_finalize_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(fp);
    VK_CHECK_LAUNCH("agg_finalize_kernel");
    return VK_OK;
}

// ------------------------------------------------------------- peer exchange ----
struct VkPeer {
    int rank = 0, world = 1;
    int device = 0;
    int64_t cap = 0;
    int words_max = 0;
    int64_t slot_words = 0;
    size_t bytes = 0;
    uint64_t* window = nullptr;
    uint64_t* peer[PEER_MAX] = {nullptr};
    bool opened[PEER_MAX] = {false};
};

int vk_peer_create(VkPeer** out, int rank, int world, int64_t cap_groups, int max_record_words) {
    VK_REQUIRE(out, "vk_peer_create: out is NULL");
    *out = nullptr;
    VK_REQUIRE(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world, "vk_peer_create: 1..16 ranks supported");
    VK_REQUIRE(cap_groups >= 1 && max_record_words >= 3, "vk_peer_create: bad slot geometry");
    VkPeer* p = new VkPeer();
    p->rank = rank;
    p->world = world;
    p->cap = cap_groups;
    p->words_max = max_record_words;
    p->slot_words = (1 + cap_groups * (int64_t) max_record_words + 15) & ~(int64_t) 15;
    p->bytes = (size_t) (PW_SLOTS + 2 * (int64_t) world * p->slot_words) * 8;
    cudaGetDevice(&p->device);
    // plain cudaMalloc: memory of the stream-ordered pool cannot be exported with cudaIpcGetMemHandle
    cudaError_t e = cudaMalloc((void**) &p->window, p->bytes);
    if (e != cudaSuccess) { delete p; return cuda_fail(e, "cudaMalloc(peer window)"); }
    e = cudaMemset(p->window, 0, p->bytes);
    if (e != cudaSuccess) { cudaFree(p->window); delete p; return cuda_fail(e, "cudaMemset(peer window)"); }
    p->peer[rank] = p->window;
    *out = p;
    return VK_OK;
}
int vk_peer_destroy(VkPeer* p) {
    if (!p) return VK_OK;
    cudaDeviceSynchronize();
    for (int r = 0; r < p->world; ++r)
        if (p->opened[r] && p->peer[r]) cudaIpcCloseMemHandle(p->peer[r]);
    if (p->window) cudaFree(p->window);
    delete p;
    return VK_OK;
}
int vk_peer_handle(VkPeer* p, void* out_handle64) {
    VK_REQUIRE(p && out_handle64, "vk_peer_handle: NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    VK_CUDA(cudaIpcGetMemHandle(&h, p->window));
    memcpy(out_handle64, &h, sizeof(h));
    return VK_OK;
}
int vk_peer_open(VkPeer* p, int peer_rank, const void* handle64) {
    VK_REQUIRE(p && handle64 && peer_rank >= 0 && peer_rank < p->world && peer_rank != p->rank, "vk_peer_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* ptr = nullptr;
    VK_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->peer[peer_rank] = reinterpret_cast<uint64_t*>(ptr);
    p->opened[peer_rank] = true;
    return VK_OK;
}
int vk_peer_attach_local(VkPeer* p, int peer_rank, VkPeer* other) {
    // ranks that live in ONE process (tests: several logical ranks on one GPU) share the address space
    VK_REQUIRE(p && other && peer_rank >= 0 && peer_rank < p->world && other->rank == peer_rank, "vk_peer_attach_local: bad argument");
    VK_REQUIRE(other->cap == p->cap && other->words_max == p->words_max && other->world == p->world,
               "vk_peer_attach_local: windows of different geometry");
    if (other->device != p->device) {
        int can = 0;
        VK_CUDA(cudaDeviceCanAccessPeer(&can, p->device, other->device));
        VK_REQUIRE(can, "vk_peer_attach_local: devices are not peer-accessible");
        cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess");
        cudaGetLastError();
    }
    p->peer[peer_rank] = other->window;
    return VK_OK;
}
void* vk_peer_decision_ptr(VkPeer* p, uint64_t epoch) {
    return p ? (void*) (p->window + PW_DECISION + (epoch & 1)) : nullptr;
}

int vk_agg_peer_send(VkAgg* a, VkPeer* p, int owner, uint64_t epoch, VkStream stream) {
    VK_REQUIRE(a && p && owner >= 0 && owner < p->world && owner != p->rank && epoch >= 1, "vk_agg_peer_send: bad argument");
    VK_REQUIRE(p->peer[owner], "vk_agg_peer_send: the owner's window is not mapped");
    const int words = a->n_keys + 2 + 3 * a->n_funcs;
    VK_REQUIRE(words <= p->words_max, "vk_agg_peer_send: records wider than the window's slots");
    cudaStream_t s = (cudaStream_t) stream;
    { const int irc = ctr_init(a, s); if (irc != VK_OK) return irc; }
    PeerSendParams sp{};
    sp.has_table = a->table_ready ? 1 : 0;
    if (a->table_ready) sp.table = a->t;
    sp.words = words;
    sp.cap = p->cap;
    sp.slot = p->peer[owner] + PW_SLOTS + ((int64_t) (epoch & 1) * p->world + p->rank) * p->slot_words;
    sp.flag = p->peer[owner] + PW_FLAG + p->rank;
    sp.ack = p->window + PW_ACK + owner;
    sp.cursor = reinterpret_cast<unsigned long long*>(p->window + PW_CURSOR);
    sp.done = reinterpret_cast<unsigned long long*>(p->window + PW_DONE);
    sp.epoch = epoch;
    VK_CUDA(cudaMemsetAsync(p->window + PW_CURSOR, 0, 16, s));
    int64_t need = a->table_ready ? (gt_total_slots(a->t) + 255) / 256 : 1, capb = (int64_t) sm_count() * 4;
    agg_peer_send_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(sp);
    VK_CHECK_LAUNCH("agg_peer_send_kernel");
    agg_peer_wait_ack_kernel<<<1, 32, 0, s>>>(p->window, owner, epoch);
    VK_CHECK_LAUNCH("agg_peer_wait_ack_kernel");
    return VK_OK;
}

int vk_agg_peer_merge(VkAgg* a, VkPeer* p, uint64_t epoch, VkStream stream) {
    VK_REQUIRE(a && p && epoch >= 1, "vk_agg_peer_merge: bad argument");
    VK_REQUIRE(a->n_keys > 0, "vk_agg_peer_merge: un-grouped aggregates merge on the host");
    const int words = a->n_keys + 2 + 3 * a->n_funcs;
    VK_REQUIRE(words <= p->words_max, "vk_agg_peer_merge: records wider than the window's slots");
    cudaStream_t s = (cudaStream_t) stream;
    { const int irc = ctr_init(a, s); if (irc != VK_OK) return irc; }
    if (!a->table_ready) {   // the owner's own shard was empty
        int rc = alloc_table(a, &a->t, pow2_ceil(4 * p->cap * p->world < 4096 ? 4096 : 4 * p->cap * p->world), s);
        if (rc != VK_OK) return rc;
        a->table_ready = true;
        a->groups_ub = 0;
    }
    PeerMergeParams mp{};
    mp.table = a->t;
    for (int f = 0; f < a->n_funcs; ++f) mp.specs[f] = a->specs[f];
    mp.words = words;
    mp.rank = p->rank;
    mp.world = p->world;
    mp.cap = p->cap;
    mp.slot_words = p->slot_words;
    mp.window = p->window;
    for (int r = 0; r < p->world; ++r) {
        VK_REQUIRE(p->peer[r], "vk_agg_peer_merge: a peer's window is not mapped");
        mp.peer_window[r] = p->peer[r];
    }
    mp.epoch = epoch;
    mp.lost = a->d_ctr + CTR_LOST;
    agg_peer_wait_kernel<<<1, 32, 0, s>>>(mp);
    VK_CHECK_LAUNCH("agg_peer_wait_kernel");
    int64_t need = ((int64_t) p->cap + 255) / 256, capb = (int64_t) sm_count() * 2;
    agg_peer_merge_kernel<<<(unsigned) (need < capb ? need : capb), 256, 0, s>>>(mp);
    VK_CHECK_LAUNCH("agg_peer_merge_kernel");
    agg_peer_ack_kernel<<<1, 32, 0, s>>>(mp);
    VK_CHECK_LAUNCH("agg_peer_ack_kernel");
    a->groups_ub += (int64_t) p->cap * (p->world - 1);
    return VK_OK;
}

}  // extern "C"
