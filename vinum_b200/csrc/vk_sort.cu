// Device radix sort + gather.
//
// Replaces Sort::Sorted (vinum_cpp/src/operators/sort/sort.cpp:15-63), i.e. Arrow's
// SortIndices (stable; NaN after every number, NULL after NaN, in BOTH directions;
// -0.0 == +0.0) followed by Take.
//
// Each sort key is mapped to an order-preserving 64-bit integer (DESC = complement of
// the numeric code, so ties keep their input order), then sorted with a stable LSD
// radix sort over 8-bit digits: one histogram pass finds the digits that actually
// vary (constant digits are skipped), and every remaining digit is ONE pass over the
// data ("onesweep": per-tile digit counts are chained with a decoupled look-back, so a
// pass reads and writes each (key, row-id) pair exactly once).  Multi-key sorts run
// the keys from last to first on the running permutation.
#include "vk_common.cuh"
#include <cstdlib>

namespace vk {

constexpr int RS_THREADS = 256;
constexpr int RS_MIN_TILE = RS_THREADS * 8;     // smallest tile any geometry uses (sizes the status array)
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef VK_RS_LB
#define VK_RS_LB 2
#endif
constexpr int RS_LB = VK_RS_LB;   // predecessor status words a look-back step reads at once
constexpr uint32_t RS_NULLBIT = 0x80000000u;

constexpr uint64_t KEY_NAN = 0xFFFFFFFFFFFFFFFEULL;
constexpr uint64_t KEY_NULL = 0xFFFFFFFFFFFFFFFFULL;

// ---------------------------------------------------------------- prepare ----
// key'[i] = code(col[perm[i]]), idx[i] = perm[i] (| NULLBIT for NULL integer keys),
// plus the 8 x 256 digit histogram of the codes.
struct PrepParams {
    Col col;
    int desc;
    int int_nulls;            // integer column with a validity bitmap: NULLs ride in idx bit 31
    const uint32_t* perm;     // nullptr: identity
    int64_t n;
    uint64_t* out_key;
    uint32_t* out_idx;
    unsigned long long* hist; // [9][256]; row 8 = NULL-bit digit (2 buckets used)
};

__device__ __forceinline__ uint64_t sort_code(const Col& c, int64_t i, int desc, bool* is_null) {
    *is_null = !col_valid(c, i);
    if (dtype_is_float(c.dtype)) {
        if (*is_null) { *is_null = false; return KEY_NULL; }  // in-band for floats
        double d = load_as_f64_raw(c, i);
        if (d != d) return KEY_NAN;
        if (d == 0.0) d = 0.0;  // -0.0 -> +0.0
        uint64_t o = f64_to_ordered((uint64_t) __double_as_longlong(d));
        return desc ? ~o : o;
    }
    if (*is_null) return 0;
    uint64_t raw = load_as_u64(c, i);
    uint64_t o = dtype_is_signed(c.dtype) ? (raw ^ 0x8000000000000000ULL) : raw;
    return desc ? ~o : o;
}

__global__ void __launch_bounds__(256) sort_prepare_kernel(const __grid_constant__ PrepParams p) {
    __shared__ uint32_t s_hist[9 * 256];
    for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    // uniform trip count so the warp votes below are convergent
    const int64_t iters = (p.n + stride - 1) / stride;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i = it * stride + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
        const bool in = i < p.n;
        uint64_t code = 0;
        uint32_t idx = 0;
        if (in) {
            const int64_t src = p.perm ? (int64_t) (p.perm[i] & ~RS_NULLBIT) : i;
            bool is_null;
            code = sort_code(p.col, src, p.desc, &is_null);
            idx = (uint32_t) src | ((p.int_nulls && is_null) ? RS_NULLBIT : 0u);
            p.out_key[i] = code;
            p.out_idx[i] = idx;
        }
        const unsigned active = __ballot_sync(0xffffffffu, in);
        if (!active) continue;
        const int leader = __ffs(active) - 1;
#pragma unroll
        for (int d = 0; d < 9; ++d) {
            const uint32_t digit = d < 8 ? (uint32_t) (code >> (8 * d)) & 0xffu : (idx >> 31);
            const uint32_t first = __shfl_sync(0xffffffffu, digit, leader);
            // constant digits are the common case (small integers): one add per warp
            if (__all_sync(0xffffffffu, !in || digit == first)) {
                if ((threadIdx.x & 31) == leader) atomicAdd(&s_hist[d * 256 + first], (uint32_t) __popc(active));
            } else if (in) {
                atomicAdd(&s_hist[d * 256 + digit], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 9 * 256; i += blockDim.x)
        if (s_hist[i]) atomicAdd(p.hist + i, (unsigned long long) s_hist[i]);
}

// Same contract for the common case -- 8-byte key column without a validity bitmap -- with U rows per
// thread per round: the U (gathered) loads are issued together, and the histogram work of a round
// runs while the next round's loads are in flight.  sort_prepare_kernel has one load in flight per
// thread at half occupancy (long scoreboard 32 %, profiles/r01_sort_prepare_ncu_full.md).
// Measured with U = 4: C4 8.22 ms against 8.56 (profiles/r02_variants.md); option SORT_PREP=0 selects
// sort_prepare_kernel.
template <int U>
__global__ void __launch_bounds__(256) sort_prepare8_kernel(const __grid_constant__ PrepParams p) {
    __shared__ uint32_t s_hist[8 * 256];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t iters = (p.n + stride * U - 1) / (stride * U);  // uniform trip count: convergent votes
    const uint64_t* data = reinterpret_cast<const uint64_t*>(p.col.data);
    const bool is_float = p.col.dtype == VK_F64, is_signed = p.col.dtype == VK_I64;
    const uint64_t flip = p.desc ? ~0ULL : 0ULL;
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i0 = it * stride * U + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t src[U];
        uint64_t code[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            src[u] = i < p.n ? (p.perm ? (p.perm[i] & ~RS_NULLBIT) : (uint32_t) i) : 0u;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) code[u] = (i0 + u * stride) < p.n ? data[src[u]] : 0ULL;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            const bool in = i < p.n;
            uint64_t c = code[u];
            if (is_float) {
                const double d = __longlong_as_double((long long) c);
                if (d != d) c = KEY_NAN;
                else {
                    if (d == 0.0) c = 0;  // -0.0 -> +0.0
                    c = f64_to_ordered(c) ^ flip;
                }
            } else {
                c = (is_signed ? (c ^ 0x8000000000000000ULL) : c) ^ flip;
            }
            if (in) {
                p.out_key[i] = c;
                p.out_idx[i] = src[u];
            }
            const unsigned active = __ballot_sync(0xffffffffu, in);
            if (!active) continue;
            const int leader = __ffs(active) - 1;
            // Which digits are the same in every active lane of the warp?  One OR-reduction of the differences
            // to the leader's code answers that for all 8 digits (two REDUX.OR); the per-digit
            // shuffle + vote of round 1 made this kernel instruction-bound (234 instructions per key).
            const uint64_t first = __shfl_sync(0xffffffffu, c, leader);
            const uint64_t diff = in ? (c ^ first) : 0ULL;
            const uint32_t dlo = __reduce_or_sync(0xffffffffu, (uint32_t) diff);
            const uint32_t dhi = __reduce_or_sync(0xffffffffu, (uint32_t) (diff >> 32));
#pragma unroll
            for (int d = 0; d < 8; ++d) {
                const uint32_t digit = (uint32_t) (c >> (8 * d)) & 0xffu;
                const uint32_t varies = ((d < 4 ? dlo : dhi) >> (8 * (d & 3))) & 0xffu;
                if (varies == 0) {
                    // constant digits are the common case (small integers): one add per warp
                    if ((threadIdx.x & 31) == leader) atomicAdd(&s_hist[d * 256 + digit], (uint32_t) __popc(active));
                } else if (in) {
                    atomicAdd(&s_hist[d * 256 + digit], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x)
        if (s_hist[i]) atomicAdd(p.hist + i, (unsigned long long) s_hist[i]);
    // row 8 of the histogram (the NULL-bit digit): no NULLs here, every row is in bucket 0
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.hist + 8 * 256, (unsigned long long) p.n);
}

// exclusive scan of each digit's 256 buckets: hist -> bucket start offsets (in place)
__global__ void __launch_bounds__(256) sort_scan_kernel(unsigned long long* hist) {
    __shared__ unsigned long long s[256];
    const int d = blockIdx.x, t = threadIdx.x;
    unsigned long long v = hist[d * 256 + t];
    s[t] = v;
    __syncthreads();
    for (int off = 1; off < 256; off <<= 1) {
        unsigned long long a = t >= off ? s[t - off] : 0;
        __syncthreads();
        s[t] += a;
        __syncthreads();
    }
    hist[d * 256 + t] = s[t] - v;
}

// -------------------------------------------------------------- radix pass ----
constexpr uint64_t RS_FLAG_SHIFT = 62;
constexpr uint64_t RS_AGG = 1ULL << RS_FLAG_SHIFT;
constexpr uint64_t RS_PREFIX = 2ULL << RS_FLAG_SHIFT;
constexpr uint64_t RS_VALUE_MASK = (1ULL << RS_FLAG_SHIFT) - 1;

struct PassParams {
    const uint64_t* in_key;
    const uint32_t* in_idx;
    uint64_t* out_key;
    uint32_t* out_idx;
    int64_t n;
    int shift;                          // 0..56, or 64 for the NULL-bit pass
    const unsigned long long* bucket_start;  // [256] for this digit
    unsigned long long* ticket;
    unsigned long long* status;         // [tiles][256]
    int64_t* out_final;                 // LAST pass: the permutation (vk_sort_indices' out_indices)
    uint64_t* out_sorted;               // LAST pass, optional: the sorted values of the key column itself
    const uint64_t* src;                // LAST pass + out_sorted: the 8-byte key column (for the codes that do not invert)
    int src_kind;                       // 0 uint64, 1 int64, 2 float64
    int desc;
};

__device__ __forceinline__ uint32_t pass_digit(uint64_t key, uint32_t idx, int shift) {
    return shift == 64 ? (idx >> 31) : (uint32_t) (key >> shift) & 0xffu;
}

// Lanes of the warp whose 8-bit digit equals this lane's.  MATCH.ANY runs on the ADU pipe at
// ~64 issue cycles per warp and was the pass kernel's binding pipe (72 % busy, profiles/); eight
// ballots (one per digit bit) and eight LOP3 do the same on the ALU / vote path.
__device__ __forceinline__ unsigned digit_peers_ballot(uint32_t d, bool in, int nbits) {
    unsigned peers = __ballot_sync(0xffffffffu, in);
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        if (b < nbits) {
            const bool bit = (d >> b) & 1u;
            const unsigned vote = __ballot_sync(0xffffffffu, in && bit);
            peers &= bit ? vote : ~vote;
        }
    }
    return peers;
}

// LAST: the final pass of the sort writes the permutation itself (row ids widened to int64 into
// PassParams::out_final, NULL flag dropped) and no keys -- nothing reads them any more -- which saves the
// 8 B/row key write and the separate widen pass (4 B/row read + 8 B/row write): C4 7.90 ms against 8.22
// (profiles/r02_variants.md).  When the caller also wants the first key column in sorted order
// (`SELECT f3 ... ORDER BY f3`), the same pass writes it too: the code inverts to the value except for
// NaN payloads and the sign of zero, and those rows fetch the original through their row id -- a random
// 8-byte gather over the whole column (take_kernel: 2.6 ms at 1e8 rows) becomes 8 B/row of extra writes.
template <int RS_ITEMS, int MINB, bool BALLOT, bool LAST = false>
__global__ void __launch_bounds__(RS_THREADS, MINB) sort_pass_kernel(const __grid_constant__ PassParams p) {
    constexpr int RS_TILE = RS_THREADS * RS_ITEMS;
    extern __shared__ __align__(16) uint8_t rs_smem[];
    uint64_t* s_key = reinterpret_cast<uint64_t*>(rs_smem);
    uint32_t* s_idx = reinterpret_cast<uint32_t*>(rs_smem + (size_t) RS_TILE * 8);
    __shared__ uint32_t s_wcnt[RS_WARPS][256];   // per-warp digit counters -> warp offsets
    __shared__ uint32_t s_lb[256];               // tile-local bucket start
    __shared__ unsigned long long s_gbase[256];  // global position of the bucket's first item of this tile
    __shared__ int64_t s_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (int64_t) atomicAdd(p.ticket, 1ULL);
    for (int i = tid; i < RS_WARPS * 256; i += RS_THREADS) (&s_wcnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t base = tile * RS_TILE;
    const int tile_n = (int) ((p.n - base) < RS_TILE ? (p.n - base) : RS_TILE);
    const unsigned lt = lanemask_lt();

    // ---- load (warp-striped: item k of lane l is element warp*512 + k*32 + l) ----
    uint64_t key[RS_ITEMS];
    uint32_t idx[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int li = warp * (32 * RS_ITEMS) + k * 32 + lane;
        if (li < tile_n) {
            key[k] = p.in_key[base + li];
            idx[k] = p.in_idx[base + li];
        } else {
            key[k] = 0;
            idx[k] = 0;
        }
    }
    // ---- stable rank inside the warp's 512 items ----
    // (Round 2 measured three other forms of this loop on one box -- digit-counter atomics pipelined four
    // items deep, MATCH.ANY for every other item, both -- all within 1 % of this one: the pass is not bound
    // by the ranking, profiles/r02_tuning.md.)
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int li = warp * (32 * RS_ITEMS) + k * 32 + lane;
        const bool in = li < tile_n;
        const uint32_t d = pass_digit(key[k], idx[k], p.shift);
        unsigned peers;
        if constexpr (BALLOT) {
            peers = digit_peers_ballot(d, in, p.shift == 64 ? 1 : 8);
            if (!in) peers = 1u << lane;
        } else {
            peers = __match_any_sync(0xffffffffu, in ? d : (0x100u | lane));
        }
        const int leader = __ffs(peers) - 1;
        uint32_t c = 0;
        if (in && lane == leader) {
            c = s_wcnt[warp][d];
            s_wcnt[warp][d] = c + __popc(peers);
        }
        c = __shfl_sync(0xffffffffu, c, leader);
        rank[k] = c + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // ---- per bucket (thread b <-> bucket b): scan over warps, publish, look back ----
    {
        const int b = tid;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            uint32_t c = s_wcnt[w][b];
            s_wcnt[w][b] = run;
            run += c;
        }
        const uint64_t total = run;
        unsigned long long* st = p.status + tile * 256 + b;
        uint64_t excl = 0;
        if (tile == 0) {
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(st), "l"(RS_PREFIX | total) : "memory");
        } else {
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(st), "l"(RS_AGG | total) : "memory");
            // The predecessors' status words are read RS_LB (2: measured 7.73 ms at C4 against 7.89 with 4, 8.40 with 8
            // and 8.1 one at a time) at a time: walking back one tile per dependent L2
            // round trip cost ~24 serial loads per bucket (19 % of all stall samples sat on this loop,
            // profiles/r02_sort_pass_ncu_full.md) because ~20 tiles are between "aggregate published" and
            // "prefix published" at any moment; independent loads overlap those round trips.
            int64_t look = tile - 1;
            bool done = false;
            while (!done) {
                unsigned long long v[RS_LB];
#pragma unroll
                for (int j = 0; j < RS_LB; ++j) {
                    const int64_t t = look - j;
                    v[j] = RS_PREFIX;   // before tile 0: prefix 0
                    if (t >= 0) asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v[j]) : "l"(p.status + t * 256 + b) : "memory");
                }
#pragma unroll
                for (int j = 0; j < RS_LB; ++j) {
                    if (!done) {
                        while ((v[j] >> RS_FLAG_SHIFT) == 0)   // not published yet: poll this one
                            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v[j]) : "l"(p.status + (look - j) * 256 + b) : "memory");
                        excl += v[j] & RS_VALUE_MASK;
                        done = (v[j] >> RS_FLAG_SHIFT) == 2;
                    }
                }
                look -= RS_LB;
            }
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(st), "l"(RS_PREFIX | (excl + total)) : "memory");
        }
        s_gbase[b] = p.bucket_start[b] + excl;
        // tile-local exclusive scan over buckets
        s_lb[b] = (uint32_t) total;
    }
    __syncthreads();
    if (warp == 0) {
        // 256 counts, 8 per lane
        uint32_t v[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[j] = s_lb[lane * 8 + j]; sum += v[j]; }
        uint32_t inc = sum;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, dlt);
            if (lane >= dlt) inc += t;
        }
        uint32_t run = inc - sum;
#pragma unroll
        for (int j = 0; j < 8; ++j) { s_lb[lane * 8 + j] = run; run += v[j]; }
    }
    __syncthreads();

    // ---- reorder inside shared memory, then write bucket runs ----
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int li = warp * (32 * RS_ITEMS) + k * 32 + lane;
        if (li < tile_n) {
            const uint32_t d = pass_digit(key[k], idx[k], p.shift);
            const uint32_t pos = s_lb[d] + s_wcnt[warp][d] + rank[k];
            s_key[pos] = key[k];
            s_idx[pos] = idx[k];
        }
    }
    __syncthreads();
    for (int i = tid; i < tile_n; i += RS_THREADS) {
        const uint64_t kx = s_key[i];
        const uint32_t ix = s_idx[i];
        const uint32_t d = pass_digit(kx, ix, p.shift);
        const unsigned long long dst = s_gbase[d] + (uint32_t) (i - s_lb[d]);
        if constexpr (LAST) {
            const uint32_t row = ix & ~RS_NULLBIT;
            p.out_final[dst] = (int64_t) row;
            if (p.out_sorted != nullptr) {
                const uint64_t o = p.desc ? ~kx : kx;
                uint64_t v;
                if (p.src_kind == 2) {
                    // NaN payloads and -0.0 are not in the code: fetch those rows' originals
                    v = ordered_to_f64(o);
                    if (kx >= KEY_NAN || (v << 1) == 0) v = p.src[row];
                } else {
                    v = p.src_kind == 1 ? (o ^ 0x8000000000000000ULL) : o;
                }
                p.out_sorted[dst] = v;
            }
        } else {
            p.out_key[dst] = kx;
            p.out_idx[dst] = ix;
        }
    }
}

__global__ void __launch_bounds__(256) sort_widen_kernel(const uint32_t* idx, int64_t n, int64_t* out) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = (int64_t) (idx[i] & ~RS_NULLBIT);
}

// ------------------------------------------------------------------- take ----
struct TakeParams {
    Col col;
    const int64_t* indices;
    int64_t n;
    void* out;
    uint8_t* out_valid;
};
__global__ void __launch_bounds__(256) take_kernel(const __grid_constant__ TakeParams p) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int es = dtype_size(p.col.dtype);
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += stride) {
        const int64_t src = p.indices[i];
        if (es == 8) reinterpret_cast<uint64_t*>(p.out)[i] = reinterpret_cast<const uint64_t*>(p.col.data)[src];
        else if (es == 4) reinterpret_cast<uint32_t*>(p.out)[i] = reinterpret_cast<const uint32_t*>(p.col.data)[src];
        else if (es == 2) reinterpret_cast<uint16_t*>(p.out)[i] = reinterpret_cast<const uint16_t*>(p.col.data)[src];
        else reinterpret_cast<uint8_t*>(p.out)[i] = p.col.data[src];
        if (p.out_valid) p.out_valid[i] = col_valid(p.col, src);
    }
}

// ------------------------------------------------------------------ top-k ----
// MSD radix select over sort_code(): level l histograms byte (7 - l) of the codes whose higher
// bytes equal `prefix`; the host walks the 256 counts to the bucket that holds the k-th row.
struct TopkParams {
    Col col;
    int desc;
    int64_t n;
    int level;                 // 0..7
    uint64_t prefix;           // the `level` bytes already fixed (right-aligned)
    unsigned long long* hist;  // [256]
};
__global__ void __launch_bounds__(256) topk_hist_kernel(const __grid_constant__ TopkParams p) {
    __shared__ uint32_t s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const int shift = 56 - 8 * p.level;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t iters = (p.n + stride - 1) / stride;  // uniform trip count: convergent votes
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i = it * stride + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
        bool hit = false;
        uint32_t digit = 0;
        if (i < p.n) {
            bool is_null;
            const uint64_t code = sort_code(p.col, i, p.desc, &is_null);
            hit = p.level == 0 || (code >> (shift + 8)) == p.prefix;
            digit = (uint32_t) (code >> shift) & 0xffu;
        }
        const unsigned active = __ballot_sync(0xffffffffu, hit);
        if (!active) continue;
        const int leader = __ffs(active) - 1;
        const uint32_t first = __shfl_sync(0xffffffffu, digit, leader);
        if (__all_sync(0xffffffffu, !hit || digit == first)) {  // one bucket per warp: the common case high up
            if ((threadIdx.x & 31) == leader) atomicAdd(&s_h[first], (uint32_t) __popc(active));
        } else if (hit) {
            atomicAdd(&s_h[digit], 1u);
        }
    }
    __syncthreads();
    if (s_h[threadIdx.x]) atomicAdd(p.hist + threadIdx.x, (unsigned long long) s_h[threadIdx.x]);
}

struct TopkCollectParams {
    Col col;
    int desc;
    int64_t n;
    int shift;                     // candidates: (code >> shift) <= bound
    uint64_t bound;
    int64_t* out;
    unsigned long long* counter;
    unsigned long long capacity;
};
__global__ void __launch_bounds__(256) topk_collect_kernel(const __grid_constant__ TopkCollectParams p) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t iters = (p.n + stride - 1) / stride;
    const unsigned lt = lanemask_lt();
    for (int64_t it = 0; it < iters; ++it) {
        const int64_t i = it * stride + (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
        bool hit = false;
        if (i < p.n) {
            bool is_null;
            hit = (sort_code(p.col, i, p.desc, &is_null) >> p.shift) <= p.bound;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (!m) continue;
        const int leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if ((threadIdx.x & 31) == leader) base = atomicAdd(p.counter, (unsigned long long) __popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        const unsigned long long pos = base + __popc(m & lt);
        if (hit && pos < p.capacity) p.out[pos] = i;
    }
}

static unsigned grid_rows(int64_t n, int per_sm = 8) {
    int64_t need = (n + 255) / 256, cap = (int64_t) sm_count() * per_sm;
    if (need < 1) need = 1;
    return (unsigned) (need < cap ? need : cap);
}

struct SortScratch {
    uint64_t* key[2];
    uint32_t* idx[2];
    unsigned long long* hist;    // [9][256]
    unsigned long long* ticket;  // 1
    unsigned long long* status;  // [tiles][256]
    size_t status_bytes;
};

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

static size_t carve(int64_t n, void* base, SortScratch* s) {
    const int64_t tiles = (n + RS_MIN_TILE - 1) / RS_MIN_TILE;  // enough status rows for any tile geometry
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base ? static_cast<uint8_t*>(base) + off : nullptr;
        off += align_up(bytes, 256);
        return p;
    };
    uint64_t* k0 = (uint64_t*) take((size_t) n * 8);
    uint64_t* k1 = (uint64_t*) take((size_t) n * 8);
    uint32_t* i0 = (uint32_t*) take((size_t) n * 4);
    uint32_t* i1 = (uint32_t*) take((size_t) n * 4);
    unsigned long long* hist = (unsigned long long*) take(9 * 256 * 8);
    unsigned long long* ticket = (unsigned long long*) take(256);
    size_t sb = (size_t) (tiles > 0 ? tiles : 1) * 256 * 8;
    unsigned long long* status = (unsigned long long*) take(sb);
    if (s) {
        s->key[0] = k0; s->key[1] = k1; s->idx[0] = i0; s->idx[1] = i1;
        s->hist = hist; s->ticket = ticket; s->status = status; s->status_bytes = sb;
    }
    return off;
}

}  // namespace vk

using namespace vk;

extern "C" {

uint64_t vk_sort_scratch_bytes(int64_t n_rows) {
    if (n_rows < 1) n_rows = 1;
    return carve(n_rows, nullptr, nullptr);
}

static int sort_indices_impl(const VkColumn* keys, const int32_t* orders, int n_keys, int64_t n_rows,
                             int64_t* out_indices, void* out_key0_sorted, void* scratch, VkStream stream) {
    VK_REQUIRE(keys && orders && n_keys >= 1, "vk_sort_indices: at least one sort key is required");
    VK_REQUIRE(n_rows >= 0, "vk_sort_indices: negative n_rows");
    if (n_rows == 0) return VK_OK;
    VK_REQUIRE(out_indices && scratch, "vk_sort_indices: NULL buffer");
    VK_REQUIRE(n_rows < (int64_t) 0x7fffffffLL, "vk_sort_indices: at most 2^31-1 rows per call");
    for (int k = 0; k < n_keys; ++k) {
        VK_REQUIRE(dtype_valid(keys[k].dtype), "vk_sort_indices: bad key dtype");
        VK_REQUIRE(keys[k].dtype != VK_BOOL8, "Sorting by boolean column is not supported yet.");  // algebra.py:191-201
        VK_REQUIRE(keys[k].length == n_rows, "vk_sort_indices: key length != n_rows");
        VK_REQUIRE(orders[k] == VK_ASC || orders[k] == VK_DESC, "vk_sort_indices: bad sort order");
    }
    if (out_key0_sorted != nullptr) {
        VK_REQUIRE(keys[0].validity == nullptr && !keys[0].nulls_as_nan &&
                   (keys[0].dtype == VK_F64 || keys[0].dtype == VK_I64 || keys[0].dtype == VK_U64),
                   "vk_sort_indices_keys: the first key must be a plain 8-byte column without NULLs");
    }
    cudaStream_t s = (cudaStream_t) stream;
    SortScratch sc;
    carve(n_rows, scratch, &sc);
    // tile geometry (measured, profiles/r01_tuning.md): 16 keys per thread, compiled for 3 CTAs per SM
    constexpr int items = 16;
    void (*pass_kernel)(PassParams) = sort_pass_kernel<items, 3, true>;
    void (*last_kernel)(PassParams) = sort_pass_kernel<items, 3, true, true>;
    const int tile_keys = RS_THREADS * items;
    const int64_t tiles = (n_rows + tile_keys - 1) / tile_keys;
    // only the status rows of this geometry's tiles are cleared per pass (the array is sized for the smallest tile)
    const size_t status_used = (size_t) tiles * 256 * sizeof(unsigned long long);
    const size_t pass_smem = (size_t) tile_keys * 12;
    VK_CUDA(cudaFuncSetAttribute(pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pass_smem));
    VK_CUDA(cudaFuncSetAttribute(last_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) pass_smem));
    const bool fuse_last = opt(OPT_SORT_FUSE_LAST) != 0;
    const int prep = (int) opt(OPT_SORT_PREP);  // rows per thread per round of the 8-byte fast path (0: general kernel)
    bool wrote_final = false;
    int cur = 0;            // buffers holding the running (key', idx)
    unsigned long long h_hist[9 * 256];

    for (int k = n_keys - 1; k >= 0; --k) {
        // ---- codes of this key in the running order + digit histograms ----
        PrepParams pp{};
        pp.col = make_col(keys[k]);
        pp.desc = orders[k] == VK_DESC;
        pp.int_nulls = keys[k].validity != nullptr && !dtype_is_float(keys[k].dtype);
        pp.perm = k < n_keys - 1 ? sc.idx[cur] : nullptr;
        pp.n = n_rows;
        const int nxt = cur ^ 1;
        pp.out_key = sc.key[nxt];
        pp.out_idx = sc.idx[nxt];
        pp.hist = sc.hist;
        VK_CUDA(cudaMemsetAsync(sc.hist, 0, 9 * 256 * 8, s));
        const bool plain8 = keys[k].validity == nullptr && !keys[k].nulls_as_nan &&
                            (keys[k].dtype == VK_F64 || keys[k].dtype == VK_I64 || keys[k].dtype == VK_U64);
        if (prep >= 4 && plain8) sort_prepare8_kernel<4><<<grid_rows(n_rows, 8), 256, 0, s>>>(pp);
        else if (prep >= 2 && plain8) sort_prepare8_kernel<2><<<grid_rows(n_rows, 8), 256, 0, s>>>(pp);
        else sort_prepare_kernel<<<grid_rows(n_rows, 4), 256, 0, s>>>(pp);
        VK_CHECK_LAUNCH("sort_prepare_kernel");
        cur = nxt;
        VK_CUDA(cudaMemcpyAsync(h_hist, sc.hist, sizeof(h_hist), cudaMemcpyDeviceToHost, s));
        VK_CUDA(cudaStreamSynchronize(s));
        sort_scan_kernel<<<9, 256, 0, s>>>(sc.hist);
        VK_CHECK_LAUNCH("sort_scan_kernel");
        // ---- one pass per digit that actually varies ----
        bool varies[9];
        int last_d = -1;
        for (int d = 0; d < 9; ++d) {
            varies[d] = true;
            for (int b = 0; b < 256; ++b)
                if (h_hist[d * 256 + b] == (unsigned long long) n_rows) { varies[d] = false; break; }
            if (varies[d]) last_d = d;
        }
        for (int d = 0; d < 9; ++d) {
            if (!varies[d]) continue;
            const bool is_final = (fuse_last || out_key0_sorted != nullptr) && k == 0 && d == last_d;
            PassParams ps{};
            ps.in_key = sc.key[cur];
            ps.in_idx = sc.idx[cur];
            ps.out_key = sc.key[cur ^ 1];
            ps.out_idx = sc.idx[cur ^ 1];
            ps.n = n_rows;
            ps.shift = d < 8 ? 8 * d : 64;
            ps.bucket_start = sc.hist + d * 256;
            ps.ticket = sc.ticket;
            ps.status = sc.status;
            VK_CUDA(cudaMemsetAsync(sc.ticket, 0, 8, s));
            VK_CUDA(cudaMemsetAsync(sc.status, 0, status_used, s));
            if (is_final) {
                ps.out_final = out_indices;
                ps.out_sorted = reinterpret_cast<uint64_t*>(out_key0_sorted);
                ps.src = reinterpret_cast<const uint64_t*>(pp.col.data);
                ps.src_kind = keys[0].dtype == VK_F64 ? 2 : (keys[0].dtype == VK_I64 ? 1 : 0);
                ps.desc = pp.desc;
                last_kernel<<<(unsigned) tiles, RS_THREADS, pass_smem, s>>>(ps);
                wrote_final = true;
            } else {
                pass_kernel<<<(unsigned) tiles, RS_THREADS, pass_smem, s>>>(ps);
            }
            VK_CHECK_LAUNCH("sort_pass_kernel");
            cur ^= 1;
        }
    }
    if (wrote_final) return VK_OK;
    sort_widen_kernel<<<grid_rows(n_rows), 256, 0, s>>>(sc.idx[cur], n_rows, out_indices);
    VK_CHECK_LAUNCH("sort_widen_kernel");
    if (out_key0_sorted != nullptr) {
        // the first key has no varying digit (a constant column): nothing was fused, gather it
        TakeParams tp{make_col(keys[0]), out_indices, n_rows, out_key0_sorted, nullptr};
        take_kernel<<<grid_rows(n_rows), 256, 0, s>>>(tp);
        VK_CHECK_LAUNCH("take_kernel");
    }
    return VK_OK;
}

int vk_sort_indices(const VkColumn* keys, const int32_t* orders, int n_keys, int64_t n_rows,
                    int64_t* out_indices, void* scratch, VkStream stream) {
    return sort_indices_impl(keys, orders, n_keys, n_rows, out_indices, nullptr, scratch, stream);
}

int vk_sort_indices_keys(const VkColumn* keys, const int32_t* orders, int n_keys, int64_t n_rows,
                         int64_t* out_indices, void* out_key0_sorted, void* scratch, VkStream stream) {
    VK_REQUIRE(out_key0_sorted, "vk_sort_indices_keys: out_key0_sorted is NULL");
    return sort_indices_impl(keys, orders, n_keys, n_rows, out_indices, out_key0_sorted, scratch, stream);
}

uint64_t vk_topk_scratch_bytes(void) { return 257 * sizeof(unsigned long long); }

int vk_topk_candidates(const VkColumn* key, int32_t order, int64_t n_rows, int64_t k, int64_t max_candidates,
                       int64_t* out_rows, int64_t* out_count, void* scratch, VkStream stream) {
    VK_REQUIRE(key && out_count, "vk_topk_candidates: NULL argument");
    VK_REQUIRE(n_rows >= 0 && k >= 0 && max_candidates >= 0, "vk_topk_candidates: negative size");
    VK_REQUIRE(dtype_valid(key->dtype), "vk_topk_candidates: bad key dtype");
    VK_REQUIRE(key->dtype != VK_BOOL8, "Sorting by boolean column is not supported yet.");  // algebra.py:191-201
    VK_REQUIRE(key->length == n_rows, "vk_topk_candidates: key length != n_rows");
    VK_REQUIRE(order == VK_ASC || order == VK_DESC, "vk_topk_candidates: bad sort order");
    *out_count = -1;
    // integer NULLs sort last outside the 64-bit code (bit 31 of the row id in the full sort)
    if (key->validity != nullptr && !dtype_is_float(key->dtype)) return VK_OK;
    if (k < 1 || k > n_rows / 8 || n_rows >= (int64_t) 0x7fffffffLL) return VK_OK;
    VK_REQUIRE(out_rows && scratch, "vk_topk_candidates: NULL buffer");
    cudaStream_t s = (cudaStream_t) stream;
    unsigned long long* d_hist = reinterpret_cast<unsigned long long*>(scratch);
    unsigned long long* d_counter = d_hist + 256;
    unsigned long long h[256];
    TopkParams tp{};
    tp.col = make_col(*key);
    tp.desc = order == VK_DESC;
    tp.n = n_rows;
    tp.hist = d_hist;
    uint64_t prefix = 0;
    int64_t below = 0, krem = k, eq = 0;
    int level = 0;
    const int64_t enough = k * 4 > (1 << 16) ? k * 4 : (1 << 16);  // stop refining below this many candidates
    for (;; ++level) {
        tp.level = level;
        tp.prefix = prefix;
        VK_CUDA(cudaMemsetAsync(d_hist, 0, 256 * sizeof(unsigned long long), s));
        topk_hist_kernel<<<grid_rows(n_rows, 8), 256, 0, s>>>(tp);
        VK_CHECK_LAUNCH("topk_hist_kernel");
        VK_CUDA(cudaMemcpyAsync(h, d_hist, sizeof(h), cudaMemcpyDeviceToHost, s));
        VK_CUDA(cudaStreamSynchronize(s));
        int64_t cum = 0;
        int b = 0;
        for (; b < 256; ++b) {
            if (cum + (int64_t) h[b] >= krem) break;
            cum += (int64_t) h[b];
        }
        if (b == 256) return fail(VK_ERR_STATE, "vk_topk_candidates: histogram does not cover k");
        below += cum;
        krem -= cum;
        eq = (int64_t) h[b];
        prefix = (prefix << 8) | (uint64_t) b;
        if (level == 7 || below + eq <= enough) break;
    }
    const int64_t count = below + eq;
    if (count > max_candidates) return VK_OK;  // too many ties at the cut: the full sort is the better plan
    TopkCollectParams cp{};
    cp.col = tp.col;
    cp.desc = tp.desc;
    cp.n = n_rows;
    cp.shift = 56 - 8 * level;
    cp.bound = prefix;
    cp.out = out_rows;
    cp.counter = d_counter;
    cp.capacity = (unsigned long long) max_candidates;
    VK_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(unsigned long long), s));
    topk_collect_kernel<<<grid_rows(n_rows, 8), 256, 0, s>>>(cp);
    VK_CHECK_LAUNCH("topk_collect_kernel");
    unsigned long long written = 0;
    VK_CUDA(cudaMemcpyAsync(&written, d_counter, sizeof(written), cudaMemcpyDeviceToHost, s));
    VK_CUDA(cudaStreamSynchronize(s));
    if ((int64_t) written != count) return fail(VK_ERR_STATE, "vk_topk_candidates: collected rows != histogram count");
    *out_count = count;
    return VK_OK;
}

int vk_take(const VkColumn* col, const int64_t* indices, int64_t n_indices, void* out, uint8_t* out_valid_bytes,
            VkStream stream) {
    VK_REQUIRE(col && n_indices >= 0, "vk_take: bad argument");
    if (n_indices == 0) return VK_OK;
    VK_REQUIRE(indices && out, "vk_take: NULL buffer");
    VK_REQUIRE(dtype_valid(col->dtype), "vk_take: bad dtype");
    VK_REQUIRE(col->validity == nullptr || out_valid_bytes, "vk_take: column has validity but no out_valid_bytes");
    TakeParams p{make_col(*col), indices, n_indices, out, col->validity ? out_valid_bytes : nullptr};
    take_kernel<<<grid_rows(n_indices), 256, 0, (cudaStream_t) stream>>>(p);
    VK_CHECK_LAUNCH("take_kernel");
    return VK_OK;
}

}  // extern "C"
