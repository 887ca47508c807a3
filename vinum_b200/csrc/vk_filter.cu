// comparison -> mask, boolean mask algebra, and order-preserving stream compaction
// (SURVEY 8a rows a1, a2, a4).  All HBM-bound byte/integer work: coalesced vector
// loads, warp-ballot ranks, a block prefix and a single-pass decoupled look-back so
// every input byte is read exactly once.
#include "vk_common.cuh"
#include <cstdlib>
#include "vk_pred.cuh"

namespace vk {

// ============================================================ compare -> mask
struct CmpParams {
    Col lhs;
    Col rhs;          // column-column compare
    int has_rhs_col;
    int op;
    int domain;
    PredScalar lo;    // scalar compare / BETWEEN low
    PredScalar hi;    // BETWEEN high
    int mode;         // 0 cmp scalar, 1 cmp column, 2 between, 3 not between
};

template <int DOM>
__device__ __forceinline__ bool cmp_row(const CmpParams& p, int64_t i) {
    typename DomT<DOM>::type x = load_dom<DOM>(p.lhs, i);
    if (p.mode == 0) return apply_cmp(p.op, x, scalar_dom<DOM>(p.lo));
    if (p.mode == 1) return apply_cmp(p.op, x, load_dom<DOM>(p.rhs, i));
    typename DomT<DOM>::type lo = scalar_dom<DOM>(p.lo), hi = scalar_dom<DOM>(p.hi);
    if (p.mode == 2) return (x >= lo) & (x <= hi);   // expressions.py:44
    return (x < lo) | (x > hi);                      // expressions.py:47
}

template <int DOM>
__global__ void __launch_bounds__(256) compare_kernel(CmpParams p, int64_t n, uint8_t* __restrict__ out) {
    // each thread produces 4 consecutive mask bytes (one 32-bit store)
    int64_t nquads = (n + 3) >> 2;
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; q < nquads; q += stride) {
        int64_t i = q << 2;
        if (i + 3 < n && ((reinterpret_cast<uintptr_t>(out) & 3) == 0)) {
            uint32_t w = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) w |= (uint32_t) cmp_row<DOM>(p, i + k) << (8 * k);
            reinterpret_cast<uint32_t*>(out)[q] = w;
        } else {
            for (int k = 0; k < 4 && i + k < n; ++k) out[i + k] = cmp_row<DOM>(p, i + k);
        }
    }
}

// Plain 8-byte column(s) compared in their own domain (no validity, no NaN view, 16-byte aligned):
// U lane-contiguous row pairs per thread, the 16-byte loads of a round issued together, two mask
// bytes stored per pair.  compare_kernel dispatches on dtype and mode row by row, which leaves one
// 8-byte load in flight per thread (SASS) -- 3.58 TB/s at 1e8 rows; this kernel with U = 2: 4.08 TB/s, with
// U = 4: 2.91 TB/s (profiles/r01_variants.md).  VINUM_B200_CMP_FAST=0 selects compare_kernel.
template <int DOM, int U>
__global__ void __launch_bounds__(256) compare8_kernel(CmpParams p, int64_t n, uint8_t* __restrict__ out) {
    using T = typename DomT<DOM>::type;
    const int64_t npairs = n >> 1;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const T lo = scalar_dom<DOM>(p.lo), hi = scalar_dom<DOM>(p.hi);
    auto eval = [&](uint64_t xb, uint64_t yb) -> uint32_t {
        const T x = bits_dom<DOM>(xb);
        if (p.mode == 0) return apply_cmp(p.op, x, lo);
        if (p.mode == 1) return apply_cmp(p.op, x, bits_dom<DOM>(yb));
        if (p.mode == 2) return (x >= lo) & (x <= hi);
        return (x < lo) | (x > hi);
    };
    for (int64_t q0 = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; q0 < npairs; q0 += stride * U) {
        uint4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = q0 + u * stride;
            a[u] = b[u] = make_uint4(0, 0, 0, 0);
            if (q < npairs) {
                a[u] = ldg_stream16(p.lhs.data + q * 16);
                if (p.mode == 1) b[u] = ldg_stream16(p.rhs.data + q * 16);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = q0 + u * stride;
            if (q < npairs) {
                const uint32_t r0 = eval(((uint64_t) a[u].y << 32) | a[u].x, ((uint64_t) b[u].y << 32) | b[u].x);
                const uint32_t r1 = eval(((uint64_t) a[u].w << 32) | a[u].z, ((uint64_t) b[u].w << 32) | b[u].z);
                reinterpret_cast<uint16_t*>(out)[q] = (uint16_t) (r0 | (r1 << 8));
            }
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = cmp_row<DOM>(p, n - 1);
}

struct IsinParams {
    Col x;
    const uint64_t* values;  // device, already converted to the domain's bit pattern
    int n_values;
    int negate;
};

template <int DOM>
__global__ void __launch_bounds__(256) isin_kernel(IsinParams p, int64_t n, uint8_t* __restrict__ out) {
    extern __shared__ uint64_t s_vals[];
    for (int i = threadIdx.x; i < p.n_values; i += blockDim.x) s_vals[i] = p.values[i];
    __syncthreads();
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        typename DomT<DOM>::type x = load_dom<DOM>(p.x, i);
        bool hit = false;
        for (int k = 0; k < p.n_values; ++k) hit |= (x == bits_dom<DOM>(s_vals[k]));
        out[i] = hit ^ (bool) p.negate;
    }
}

// ================================================================ mask algebra
__global__ void __launch_bounds__(256) mask_combine_kernel(int op, const uint8_t* __restrict__ a,
                                                           const uint8_t* __restrict__ b, int64_t n,
                                                           uint8_t* __restrict__ out) {
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    int64_t i0 = (int64_t) blockIdx.x * blockDim.x + threadIdx.x;
    bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                 reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    int64_t nvec = vec ? (n >> 4) : 0;
    const uint4 ones = make_uint4(0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u);
    for (int64_t v = i0; v < nvec; v += stride) {
        uint4 x = reinterpret_cast<const uint4*>(a)[v];
        uint4 r;
        if (op == VK_MASK_NOT) {
            r = make_uint4(x.x ^ ones.x, x.y ^ ones.y, x.z ^ ones.z, x.w ^ ones.w);
        } else {
            uint4 y = reinterpret_cast<const uint4*>(b)[v];
            if (op == VK_MASK_AND) r = make_uint4(x.x & y.x, x.y & y.y, x.z & y.z, x.w & y.w);
            else r = make_uint4(x.x | y.x, x.y | y.y, x.z | y.z, x.w | y.w);
        }
        reinterpret_cast<uint4*>(out)[v] = r;
    }
    for (int64_t i = (nvec << 4) + i0; i < n; i += stride) {
        uint8_t x = a[i];
        out[i] = op == VK_MASK_NOT ? (x ^ 1) : (op == VK_MASK_AND ? (x & b[i]) : (x | b[i]));
    }
}

__global__ void __launch_bounds__(256) is_null_kernel(Col x, int want_valid, int64_t n, uint8_t* __restrict__ out) {
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = col_valid(x, i) == (bool) want_valid;
}

// 8 mask bytes -> 1 bitmap byte (LSB first).  Each thread packs 32 rows.
__global__ void __launch_bounds__(256) mask_to_bits_kernel(const uint8_t* __restrict__ mask, int64_t n,
                                                           uint8_t* __restrict__ bits) {
    int64_t nbytes = (n + 7) >> 3;
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t b = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; b < nbytes; b += stride) {
        int64_t i = b << 3;
        uint32_t v = 0;
        if (i + 7 < n && ((reinterpret_cast<uintptr_t>(mask) & 7) == 0)) {
            uint64_t w = reinterpret_cast<const uint64_t*>(mask)[b];
            // gather bit 0 of every byte
            w &= 0x0101010101010101ULL;
            v = (uint32_t) ((w * 0x0102040810204080ULL) >> 56);
        } else {
            for (int k = 0; k < 8 && i + k < n; ++k) v |= (uint32_t) (mask[i + k] & 1) << k;
        }
        bits[b] = (uint8_t) v;
    }
}
__global__ void __launch_bounds__(256) bits_to_mask_kernel(const uint8_t* __restrict__ bits, int64_t bit_offset,
                                                           int64_t n, uint8_t* __restrict__ mask) {
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        int64_t b = bit_offset + i;
        mask[i] = (bits[b >> 3] >> (b & 7)) & 1;
    }
}

// ========================================================= stream compaction
// One CTA owns one tile of 512 * ITERS rows (ticketed, so predecessors are always running or
// done).  Rows are mapped lane-contiguously in pairs -- row = it*512 + tid*2 + e --
// so an 8-byte column is read with one 16-byte load per lane and the rank of a row
// inside its tile is (rows selected by lower (it, warp)) + (ballot rank).
constexpr int FT_THREADS = 256;
constexpr int FT_MIN_TILE = FT_THREADS * 2 * 4;  // smallest tile geometry (2048 rows): sizes the status array
constexpr int FT_MAX_COLS = 12;

constexpr uint64_t ST_FLAG_SHIFT = 62;
constexpr uint64_t ST_AGG = 1ULL << ST_FLAG_SHIFT;
constexpr uint64_t ST_PREFIX = 2ULL << ST_FLAG_SHIFT;
constexpr uint64_t ST_VALUE_MASK = (1ULL << ST_FLAG_SHIFT) - 1;

struct FilterParams {
    Pred pred;
    int64_t n;
    int n_cols;
    Col cols[FT_MAX_COLS];
    void* out_data[FT_MAX_COLS];
    uint8_t* out_valid[FT_MAX_COLS];
    int64_t* out_rows;
    unsigned long long* ticket;  // scratch[0]
    unsigned long long* status;  // scratch[1..]
    int64_t num_tiles;
    int pf;  // bulk-prefetch the tile's payload columns into L2 while the predicate / look-back run (1: at tile start, 2: after phase 1)
};

__device__ __forceinline__ void l2_prefetch_span(const uint8_t* begin, const uint8_t* end) {
    const uint64_t a = reinterpret_cast<uint64_t>(begin) & ~(uint64_t) 15;
    const uint64_t e = reinterpret_cast<uint64_t>(end) & ~(uint64_t) 15;
    if (e > a) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a), "r"((uint32_t) (e - a)) : "memory");
}

__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// PK: predicate kind (vk_pred.cuh); ITERS: row pairs per thread (tile = 512 * ITERS rows).
// STAGE (8-byte CMP predicates only): the 16 bytes of the predicate column that phase 1 loads per row
// pair are parked in shared memory (one conflict-free STS.128 per pair, 16 KB per 2048-row tile), and an
// output column that IS the predicate column (`SELECT * ... WHERE f0 > c`) is scattered from there
// instead of being read from DRAM a second time: C2 moves 4.8 GB instead of 5.6 GB.  Keeping the pairs
// in registers instead costs 16 registers and with them a quarter of the resident tiles (measured
// -30 % in round 1); shared memory is otherwise unused by this kernel.
template <int PK, int ITERS, bool STAGE>
__global__ void __launch_bounds__(FT_THREADS, 8) filter_kernel(const __grid_constant__ FilterParams p) {
    static_assert(!STAGE || PK == PK_F64_VEC || PK == PK_I64_VEC, "staging needs an 8-byte compare predicate");
    constexpr int TILE = FT_THREADS * 2 * ITERS;
    constexpr int NCNT = ITERS * (FT_THREADS / 32);   // (iter, warp) counts: 32 or 64
    constexpr int PER_LANE = NCNT / 32;
    static_assert(NCNT % 32 == 0, "whole counts per lane");
    __shared__ int64_t s_tile;
    __shared__ uint32_t s_cnt[NCNT];
    __shared__ int64_t s_excl;
    __shared__ __align__(16) uint4 s_pred[STAGE ? ITERS * FT_THREADS : 1];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_tile = (int64_t) atomicAdd(p.ticket, 1ULL);
    __syncthreads();
    const int64_t tile = s_tile;
    const int64_t base = tile * TILE;

    // The payload columns are only read after the look-back; start moving this tile's slice of
    // each of them into L2 now (one bulk prefetch per column, no registers, no shared memory).
    if (p.pf == 1 && tid < p.n_cols) {
        const Col col = p.cols[tid];
        const int es = dtype_size(col.dtype);
        const int64_t rows = p.n - base < TILE ? p.n - base : TILE;
        // (the predicate's own column is being loaded by phase 1 right now: no second request)
        if (!(p.pred.kind == VK_PRED_CMP && col.data == p.pred.col.data))
            l2_prefetch_span(col.data + base * es, col.data + (base + rows) * es);
    }

    // ---- phase 1: evaluate the predicate once, keep the flags in registers ----
    uint32_t flags = 0;
    uint32_t lane_off[ITERS];
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int64_t r0 = base + it * (FT_THREADS * 2) + tid * 2;
        bool f0, f1;
        if constexpr (STAGE) {
            f0 = f1 = false;
            uint4 q = make_uint4(0, 0, 0, 0);
            if (r0 + 1 < p.n) {
                q = ldg_stream16(p.pred.col.data + r0 * 8);
            } else if (r0 < p.n) {  // last, unpaired row of the batch
                const uint2 h = *reinterpret_cast<const uint2*>(p.pred.col.data + r0 * 8);
                q.x = h.x;
                q.y = h.y;
            }
            s_pred[it * FT_THREADS + tid] = q;
            if constexpr (PK == PK_F64_VEC) {
                const double c = __longlong_as_double((long long) p.pred.scalar.bits);
                if (r0 < p.n) f0 = apply_cmp(p.pred.op, __hiloint2double(q.y, q.x), c);
                if (r0 + 1 < p.n) f1 = apply_cmp(p.pred.op, __hiloint2double(q.w, q.z), c);
            } else {
                const int64_t c = (int64_t) p.pred.scalar.bits;
                if (r0 < p.n) f0 = apply_cmp(p.pred.op, (int64_t) (((uint64_t) q.y << 32) | q.x), c);
                if (r0 + 1 < p.n) f1 = apply_cmp(p.pred.op, (int64_t) (((uint64_t) q.w << 32) | q.z), c);
            }
        } else {
            pred_pair<PK>(p.pred, r0, p.n, f0, f1);
        }
        unsigned b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
        lane_off[it] = __popc(b0 & lt) + __popc(b1 & lt);
        flags |= ((uint32_t) f0 << (2 * it)) | ((uint32_t) f1 << (2 * it + 1));
        if (lane == 0) s_cnt[it * (FT_THREADS / 32) + warp] = __popc(b0) + __popc(b1);
    }
    // (FILTER_PF=2) the same prefetch issued only now, after the predicate loads: the slices sit in L2 for a
    // shorter time -- 1184 resident tiles x 64 KB of prefetched payload is most of the L2
    if (p.pf == 2 && tid < p.n_cols) {
        const Col col = p.cols[tid];
        const int es = dtype_size(col.dtype);
        const int64_t rows = p.n - base < TILE ? p.n - base : TILE;
        if (!(p.pred.kind == VK_PRED_CMP && col.data == p.pred.col.data))
            l2_prefetch_span(col.data + base * es, col.data + (base + rows) * es);
    }
    __syncthreads();

    // ---- block scan of the (iter, warp) counts + decoupled look-back (warp 0) ----
    if (warp == 0) {
        uint32_t c[PER_LANE], mine = 0;
#pragma unroll
        for (int e = 0; e < PER_LANE; ++e) {
            c[e] = s_cnt[lane * PER_LANE + e];
            mine += c[e];
        }
        uint32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        uint32_t run = inc - mine;  // exclusive offset of this lane's first (iter, warp) entry inside the tile
#pragma unroll
        for (int e = 0; e < PER_LANE; ++e) {
            s_cnt[lane * PER_LANE + e] = run;
            run += c[e];
        }
        uint64_t total = __shfl_sync(0xffffffffu, inc, 31);
        uint64_t excl = 0;
        if (tile == 0) {
            if (lane == 0) st_status(p.status, ST_PREFIX | total);
        } else {
            if (lane == 0) st_status(p.status + tile, ST_AGG | total);
            int64_t look = tile - 1;
            while (true) {
                int64_t idx = look - lane;
                unsigned long long st = ST_PREFIX;  // virtual tile -1: prefix 0
                if (idx >= 0) {
                    do { st = ld_status(p.status + idx); } while ((st >> ST_FLAG_SHIFT) == 0);
                }
                unsigned is_prefix = __ballot_sync(0xffffffffu, (st >> ST_FLAG_SHIFT) == 2 || idx < 0);
                int first = is_prefix ? __ffs(is_prefix) - 1 : 32;
                uint64_t v = (lane <= first) ? (st & ST_VALUE_MASK) : 0;
                if (idx < 0) v = 0;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                excl += v;
                if (is_prefix) break;
                look -= 32;
            }
            if (lane == 0) st_status(p.status + tile, ST_PREFIX | (excl + total));
        }
        if (lane == 0) {
            s_excl = (int64_t) excl;
            if (tile == p.num_tiles - 1) *p.out_rows = (int64_t) (excl + total);
        }
    }
    __syncthreads();
    const int64_t tile_excl = s_excl;

    // ---- phase 2: scatter every column; a warp writes one contiguous run per iter ----
    for (int c = 0; c < p.n_cols; ++c) {
        const Col col = p.cols[c];
        const int es = dtype_size(col.dtype);
        uint8_t* outv = p.out_valid[c];
        const bool vec16 = (es == 8) && ((reinterpret_cast<uintptr_t>(col.data) & 15) == 0);
        const bool staged = STAGE && es == 8 && col.data == p.pred.col.data;
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
            const uint32_t f = (flags >> (2 * it)) & 3u;
            if (f == 0) continue;
            const int64_t r0 = base + it * (FT_THREADS * 2) + tid * 2;
            int64_t pos = tile_excl + s_cnt[it * (FT_THREADS / 32) + warp] + lane_off[it];
            if (es == 8) {
                uint64_t v0, v1;
                if (staged) {
                    const uint4 q = s_pred[STAGE ? it * FT_THREADS + tid : 0];
                    v0 = ((uint64_t) q.y << 32) | q.x;
                    v1 = ((uint64_t) q.w << 32) | q.z;
                } else if (vec16 && r0 + 1 < p.n) {
                    uint4 q = ldg_stream16(col.data + r0 * 8);
                    v0 = ((uint64_t) q.y << 32) | q.x;
                    v1 = ((uint64_t) q.w << 32) | q.z;
                } else {
                    v0 = (f & 1) ? reinterpret_cast<const uint64_t*>(col.data)[r0] : 0;
                    v1 = (f & 2) ? reinterpret_cast<const uint64_t*>(col.data)[r0 + 1] : 0;
                }
                uint64_t* o = reinterpret_cast<uint64_t*>(p.out_data[c]);
                if (f & 1) o[pos++] = v0;
                if (f & 2) o[pos] = v1;
            } else {
                int64_t q = pos;
                if (f & 1) store_from_u64(p.out_data[c], col.dtype, q++, load_as_u64(col, r0));
                if (f & 2) store_from_u64(p.out_data[c], col.dtype, q, load_as_u64(col, r0 + 1));
            }
            if (outv != nullptr) {
                int64_t q = tile_excl + s_cnt[it * (FT_THREADS / 32) + warp] + lane_off[it];
                if (f & 1) outv[q++] = col_valid(col, r0);
                if (f & 2) outv[q] = col_valid(col, r0 + 1);
            }
        }
    }
}

static int grid_for(int64_t work_items, int per_sm = 8) {
    int64_t need = (work_items + 255) / 256;
    int64_t cap = (int64_t) sm_count() * per_sm;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}

}  // namespace vk

using namespace vk;

extern "C" {

static int launch_compare(const CmpParams& p, int64_t n, uint8_t* out, VkStream stream) {
    if (n == 0) return VK_OK;
    VK_REQUIRE(out, "compare: out_mask is NULL");
    int g = grid_for((n + 3) / 4);
    cudaStream_t s = (cudaStream_t) stream;
    const int fast = (int) opt(OPT_CMP_FAST);  // 2: compare8_kernel (two row pairs per thread), 0: compare_kernel
    auto plain8 = [&](const Col& c) {
        const int want = p.domain == DOM_F64 ? VK_F64 : (p.domain == DOM_I64 ? VK_I64 : (p.domain == DOM_U64 ? VK_U64 : -1));
        return c.dtype == want && c.validity == nullptr && !c.nan_nulls && (reinterpret_cast<uintptr_t>(c.data) & 15) == 0;
    };
    if (fast >= 2 && plain8(p.lhs) && (p.mode != 1 || plain8(p.rhs)) && (reinterpret_cast<uintptr_t>(out) & 1) == 0) {
        const int gp = grid_for((n / 2 + 3) / 4);
#define VK_CMP8_GO(DOM)                                                             \
        compare8_kernel<DOM, 2><<<gp, 256, 0, s>>>(p, n, out)
        switch (p.domain) {
            case DOM_I64: VK_CMP8_GO(DOM_I64); break;
            case DOM_U64: VK_CMP8_GO(DOM_U64); break;
            default: VK_CMP8_GO(DOM_F64); break;
        }
#undef VK_CMP8_GO
        VK_CHECK_LAUNCH("compare8_kernel");
        return VK_OK;
    }
    switch (p.domain) {
        case DOM_I64: compare_kernel<DOM_I64><<<g, 256, 0, s>>>(p, n, out); break;
        case DOM_U64: compare_kernel<DOM_U64><<<g, 256, 0, s>>>(p, n, out); break;
        case DOM_F32: compare_kernel<DOM_F32><<<g, 256, 0, s>>>(p, n, out); break;
        default: compare_kernel<DOM_F64><<<g, 256, 0, s>>>(p, n, out); break;
    }
    VK_CHECK_LAUNCH("compare_kernel");
    return VK_OK;
}

int vk_compare_scalar(const VkColumn* lhs, int op, const VkScalar* rhs, uint8_t* out_mask, VkStream stream) {
    VK_REQUIRE(lhs && rhs, "vk_compare_scalar: NULL argument");
    VK_REQUIRE(op >= VK_EQ && op <= VK_LE, "vk_compare_scalar: bad op");
    VK_REQUIRE(dtype_valid(lhs->dtype), "vk_compare_scalar: bad dtype");
    CmpParams p{};
    p.lhs = make_col(*lhs);
    p.op = op;
    p.mode = 0;
    p.domain = pick_domain(*lhs, rhs->dtype);
    int rc = convert_scalar(*rhs, p.domain, &p.lo);
    if (rc != VK_OK) return rc;
    return launch_compare(p, lhs->length, out_mask, stream);
}

int vk_compare_columns(const VkColumn* lhs, int op, const VkColumn* rhs, uint8_t* out_mask, VkStream stream) {
    VK_REQUIRE(lhs && rhs, "vk_compare_columns: NULL argument");
    VK_REQUIRE(op >= VK_EQ && op <= VK_LE, "vk_compare_columns: bad op");
    VK_REQUIRE(lhs->length == rhs->length, "vk_compare_columns: length mismatch");
    VK_REQUIRE(dtype_valid(lhs->dtype) && dtype_valid(rhs->dtype), "vk_compare_columns: bad dtype");
    CmpParams p{};
    p.lhs = make_col(*lhs);
    p.rhs = make_col(*rhs);
    p.has_rhs_col = 1;
    p.op = op;
    p.mode = 1;
    int d = pick_domain_cols(*lhs, *rhs);
    if (d < 0) return fail(VK_ERR_UNSUPPORTED, "vk_compare_columns: mixed signed/unsigned 64-bit compare");
    p.domain = d;
    return launch_compare(p, lhs->length, out_mask, stream);
}

int vk_between_scalar(const VkColumn* x, const VkScalar* lo, const VkScalar* hi, int negate,
                      uint8_t* out_mask, VkStream stream) {
    VK_REQUIRE(x && lo && hi, "vk_between_scalar: NULL argument");
    VK_REQUIRE(dtype_valid(x->dtype), "vk_between_scalar: bad dtype");
    CmpParams p{};
    p.lhs = make_col(*x);
    p.mode = negate ? 3 : 2;
    int any_float = (lo->dtype == VK_F64 || hi->dtype == VK_F64) ? VK_F64 : lo->dtype;
    p.domain = pick_domain(*x, any_float);
    int rc = convert_scalar(*lo, p.domain, &p.lo);
    if (rc != VK_OK) return rc;
    rc = convert_scalar(*hi, p.domain, &p.hi);
    if (rc != VK_OK) return rc;
    return launch_compare(p, x->length, out_mask, stream);
}

int vk_isin_scalars(const VkColumn* x, const VkScalar* host_values, int n_values, int negate,
                    uint8_t* out_mask, VkStream stream) {
    VK_REQUIRE(x && (host_values || n_values == 0), "vk_isin_scalars: NULL argument");
    VK_REQUIRE(n_values >= 0 && n_values <= 4096, "vk_isin_scalars: 0..4096 values supported");
    VK_REQUIRE(dtype_valid(x->dtype), "vk_isin_scalars: bad dtype");
    int64_t n = x->length;
    if (n == 0) return VK_OK;
    int any = VK_I64;
    for (int i = 0; i < n_values; ++i)
        if (host_values[i].dtype == VK_F64) any = VK_F64;
        else if (host_values[i].dtype == VK_U64 && any != VK_F64) any = VK_U64;
    int dom = pick_domain(*x, any);
    uint64_t host_bits[4096];
    int kept = 0;
    for (int i = 0; i < n_values; ++i) {
        PredScalar ps;
        int rc = convert_scalar(host_values[i], dom, &ps);
        if (rc == VK_ERR_UNSUPPORTED) continue;  // value not representable in the column's domain: never equal
        if (rc != VK_OK) return rc;
        host_bits[kept++] = ps.bits;
    }
    uint64_t* dvals = nullptr;
    cudaStream_t s = (cudaStream_t) stream;
    VK_CUDA(cudaMallocAsync((void**) &dvals, sizeof(uint64_t) * (kept ? kept : 1), s));
    if (kept) VK_CUDA(cudaMemcpyAsync(dvals, host_bits, sizeof(uint64_t) * kept, cudaMemcpyHostToDevice, s));
    IsinParams p{make_col(*x), dvals, kept, negate};
    int g = grid_for(n);
    size_t smem = sizeof(uint64_t) * (kept ? kept : 1);
    switch (dom) {
        case DOM_I64: isin_kernel<DOM_I64><<<g, 256, smem, s>>>(p, n, out_mask); break;
        case DOM_U64: isin_kernel<DOM_U64><<<g, 256, smem, s>>>(p, n, out_mask); break;
        case DOM_F32: isin_kernel<DOM_F32><<<g, 256, smem, s>>>(p, n, out_mask); break;
        default: isin_kernel<DOM_F64><<<g, 256, smem, s>>>(p, n, out_mask); break;
    }
    VK_CHECK_LAUNCH("isin_kernel");
    VK_CUDA(cudaFreeAsync(dvals, s));
    return VK_OK;
}

int vk_mask_combine(int op, const uint8_t* a, const uint8_t* b, int64_t n, uint8_t* out_mask, VkStream stream) {
    VK_REQUIRE(op >= VK_MASK_AND && op <= VK_MASK_NOT, "vk_mask_combine: bad op");
    if (n == 0) return VK_OK;
    VK_REQUIRE(a && out_mask && (b || op == VK_MASK_NOT), "vk_mask_combine: NULL argument");
    if (op == VK_MASK_NOT) b = a;
    mask_combine_kernel<<<grid_for((n + 15) / 16), 256, 0, (cudaStream_t) stream>>>(op, a, b, n, out_mask);
    VK_CHECK_LAUNCH("mask_combine_kernel");
    return VK_OK;
}

int vk_is_null(const VkColumn* x, int want_valid, uint8_t* out_mask, VkStream stream) {
    VK_REQUIRE(x, "vk_is_null: NULL argument");
    if (x->length == 0) return VK_OK;
    is_null_kernel<<<grid_for(x->length), 256, 0, (cudaStream_t) stream>>>(make_col(*x), want_valid, x->length, out_mask);
    VK_CHECK_LAUNCH("is_null_kernel");
    return VK_OK;
}

int vk_mask_to_bits(const uint8_t* mask, int64_t n, uint8_t* out_bits, VkStream stream) {
    if (n == 0) return VK_OK;
    VK_REQUIRE(mask && out_bits, "vk_mask_to_bits: NULL argument");
    mask_to_bits_kernel<<<grid_for((n + 7) / 8), 256, 0, (cudaStream_t) stream>>>(mask, n, out_bits);
    VK_CHECK_LAUNCH("mask_to_bits_kernel");
    return VK_OK;
}
int vk_bits_to_mask(const uint8_t* bits, int64_t bit_offset, int64_t n, uint8_t* out_mask, VkStream stream) {
    if (n == 0) return VK_OK;
    VK_REQUIRE(bits && out_mask, "vk_bits_to_mask: NULL argument");
    bits_to_mask_kernel<<<grid_for(n), 256, 0, (cudaStream_t) stream>>>(bits, bit_offset, n, out_mask);
    VK_CHECK_LAUNCH("bits_to_mask_kernel");
    return VK_OK;
}

uint64_t vk_filter_scratch_bytes(int64_t n_rows) {
    int64_t tiles = (n_rows + FT_MIN_TILE - 1) / FT_MIN_TILE;
    if (tiles < 1) tiles = 1;
    return (uint64_t) (tiles + 1) * sizeof(unsigned long long);
}

int vk_filter(const VkPredicate* pred, int64_t n_rows, const VkColumn* cols, int n_cols,
              void* const* out_data, uint8_t* const* out_valid_bytes, int64_t* out_rows,
              void* scratch, VkStream stream) {
    VK_REQUIRE(pred && out_rows && scratch, "vk_filter: NULL argument");
    VK_REQUIRE(n_rows >= 0 && n_cols >= 0, "vk_filter: negative size");
    VK_REQUIRE(n_cols == 0 || (cols && out_data), "vk_filter: NULL column arrays");
    VK_REQUIRE(pred->kind == VK_PRED_MASK || pred->kind == VK_PRED_CMP || pred->kind == VK_PRED_EXPR,
               "vk_filter: predicate kind must be MASK, CMP or EXPR");
    cudaStream_t s = (cudaStream_t) stream;
    if (n_rows == 0) {
        VK_CUDA(cudaMemsetAsync(out_rows, 0, sizeof(int64_t), s));
        return VK_OK;
    }
    Pred dp;
    int pk = 0;
    int rc = make_pred(*pred, n_rows, &dp, &pk);
    if (rc != VK_OK) return rc;
    for (int c = 0; c < n_cols; ++c) {
        VK_REQUIRE(dtype_valid(cols[c].dtype), "vk_filter: bad column dtype");
        VK_REQUIRE(cols[c].length == n_rows, "vk_filter: column length != n_rows");
        VK_REQUIRE(out_data[c], "vk_filter: NULL output buffer");
        VK_REQUIRE(cols[c].validity == nullptr || (out_valid_bytes && out_valid_bytes[c]),
                   "vk_filter: column has validity but no out_valid_bytes buffer");
    }
    // geometry (measured, profiles/r02_variants.md): 2048-row tiles, 8 CTAs per SM
    constexpr int iters = 4;   // 4096-row tiles measured 12 % slower (profiles/r02_variants.md)
    const int tile_rows = FT_THREADS * 2 * iters;
    const int64_t tiles = (n_rows + tile_rows - 1) / tile_rows;
    // columns are processed FT_MAX_COLS at a time; each pass re-evaluates the predicate
    int c0 = 0;
    do {
        FilterParams p{};
        p.pred = dp;
        p.n = n_rows;
        p.n_cols = (n_cols - c0 < FT_MAX_COLS) ? n_cols - c0 : FT_MAX_COLS;
        bool pred_is_output = false;
        for (int c = 0; c < p.n_cols; ++c) {
            p.cols[c] = make_col(cols[c0 + c]);
            p.out_data[c] = out_data[c0 + c];
            p.out_valid[c] = (cols[c0 + c].validity && out_valid_bytes) ? out_valid_bytes[c0 + c] : nullptr;
            if ((pk == PK_F64_VEC || pk == PK_I64_VEC) && p.cols[c].data == dp.col.data && dtype_size(p.cols[c].dtype) == 8)
                pred_is_output = true;
        }
        p.out_rows = out_rows;
        p.ticket = reinterpret_cast<unsigned long long*>(scratch);
        p.status = p.ticket + 1;
        p.num_tiles = tiles;
        p.pf = (int) opt(OPT_FILTER_PF);
        const bool stage = pred_is_output && opt(OPT_FILTER_STAGE) != 0;
        VK_CUDA(cudaMemsetAsync(scratch, 0, vk_filter_scratch_bytes(n_rows), s));
#define VK_FILTER_GO(PK, STAGEABLE)                                                                   \
        do {                                                                                              \
            if (STAGEABLE && stage) filter_kernel<PK, 4, STAGEABLE><<<(unsigned) tiles, FT_THREADS, 0, s>>>(p); \
            else filter_kernel<PK, 4, false><<<(unsigned) tiles, FT_THREADS, 0, s>>>(p);                  \
        } while (0)
        switch (pk) {
            case PK_MASK: VK_FILTER_GO(PK_MASK, false); break;
            case PK_F64_VEC: VK_FILTER_GO(PK_F64_VEC, true); break;
            case PK_I64_VEC: VK_FILTER_GO(PK_I64_VEC, true); break;
            default: VK_FILTER_GO(PK_GENERIC, false); break;
        }
#undef VK_FILTER_GO
        VK_CHECK_LAUNCH("filter_kernel");
        c0 += p.n_cols;
    } while (c0 < n_cols);
    return VK_OK;
}

}  // extern "C"
