// Shared device/host helpers for the vinum_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <atomic>
#include "../../include/vinum_b200.h"

namespace vk {

// ---------------------------------------------------------------- errors ----
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
int cuda_fail(cudaError_t e, const char* what);
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define VK_CUDA(call)                                                  \
    do {                                                               \
        cudaError_t _e = (call);                                       \
        if (_e != cudaSuccess) return ::vk::cuda_fail(_e, #call);      \
    } while (0)

#define VK_CHECK_LAUNCH(name)                                          \
    do {                                                               \
        ::vk::count_launch();                                          \
        cudaError_t _e = cudaGetLastError();                           \
        if (_e != cudaSuccess) return ::vk::cuda_fail(_e, name);       \
    } while (0)

#define VK_REQUIRE(cond, msg)                                          \
    do {                                                               \
        if (!(cond)) return ::vk::fail(VK_ERR_ARG, msg);               \
    } while (0)

// ---------------------------------------------------------------- options ----
// Kernel-selection knobs.  Each has a compiled-in default, may be preset by the
// environment variable VINUM_B200_<NAME> (read once, at first use) and can be changed at any time
// with vk_set_option(), so that one process can exercise every path (tests/test_gpu_paths.py).
enum Opt {
    OPT_FILTER_STAGE = 0,   // 1: an output column that is the predicate column is scattered from shared memory
    OPT_FILTER_PF,          // bulk-prefetch of a tile's payload slices into L2: 0 off, 1 at tile start, 2 after the predicate loads
    OPT_CMP_FAST,           // compare8_kernel row pairs per thread (0: compare_kernel)
    OPT_ARITH_FAST,         // arith8_kernel row pairs per thread (0: arith_kernel)
    OPT_ONEGROUP_FAST,      // agg_onegroup8_kernel row pairs per thread (0: agg_onegroup_kernel)
    OPT_SORT_FUSE_LAST,     // 1: the last radix pass writes the int64 permutation itself
    OPT_SORT_PREP,          // sort_prepare8_kernel loads per thread (0: sort_prepare_kernel)
    OPT_AGG_LOG2S,          // agg_fast_kernel: log2 of the CTA key table slots (hash mode)
    OPT_AGG_PF,             // agg_fast_kernel: L2 prefetch distance in tiles (-1: automatic)
    OPT_AGG_WARPS,          // agg_fast_kernel: warps per CTA (0: automatic)
    OPT_AGG_DIRECT,         // 1: direct (key - base) group ids when the key range allows, 0: always hash
    OPT_AGG_DICT,           // 1: hash mode looks keys up in a host-built read-only cuckoo dictionary
    OPT_AGG_ENTRY,          // agg_fast_kernel, COUNT + SUM(f64) entry: 0 = 16-byte entries, 2 = split entries (SUM array + {tag | COUNT} array), 1 = by selectivity
    OPT_AGG_HOT,            // hot-group step (>= 6 lanes of a warp on one entry -> one butterfly-reduced update): 1 = when the learning launch saw a key with >= 30 % of the rows, 2 = always, 0 = never
    OPT_AGG_NOFAST,         // 1: never use the shared-memory aggregate kernel
    OPT_AGG_WIDE,           // global-table kernel on single-key tables: 1 = agg_wide_kernel (four rows per thread, lock-step probing), 0 = the one-row agg_general_kernel
    OPT_AGG_PARTITION,      // tables beyond the L2: scatter (key, row) into buckets = table slices + update slice by slice: 1 = when the touched table exceeds 96 MB, 2 = always (single key), 0 = never
    OPT_AGG_LEARN_LOG2,     // log2 rows of the learning launch
    OPT_LIST_LOG2,          // log2 of the replay-list capacity cap (entries)
    OPT_DEBUG,              // 1: trace the aggregate's host decisions to stderr
    OPT_INGEST_STAGED,      // 1: pageable host memory goes through the pinned bounce-buffer pool (vk_ingest.cu)
    OPT_INGEST_THREADS,     // worker threads of that pool (0: automatic; read when the pool starts)
    OPT_INGEST_PIECE_KB,    // bytes per bounce copy, in KB
    OPT_COUNT
};
int64_t opt(int id);

int sm_count();           // SMs of the current device (cached per device)
int max_smem_optin();     // max opt-in dynamic shared memory per block

// ---------------------------------------------------------------- dtypes ----
__host__ __device__ inline int dtype_size(int dt) {
    switch (dt) {
        case VK_I8: case VK_U8: case VK_BOOL8: return 1;
        case VK_I16: case VK_U16: return 2;
        case VK_I32: case VK_U32: case VK_F32: return 4;
        case VK_I64: case VK_U64: case VK_F64: return 8;
        default: return 0;
    }
}
__host__ __device__ inline bool dtype_is_float(int dt) { return dt == VK_F32 || dt == VK_F64; }
__host__ __device__ inline bool dtype_is_signed(int dt) { return dt >= VK_I8 && dt <= VK_I64; }
__host__ __device__ inline bool dtype_is_unsigned(int dt) {
    return (dt >= VK_U8 && dt <= VK_U64) || dt == VK_BOOL8;
}
inline bool dtype_valid(int dt) { return dt >= VK_I8 && dt <= VK_BOOL8; }

// Device-side column view (copy of VkColumn with the offset folded in).
struct Col {
    const uint8_t* data;      // already advanced by offset * elsize
    const uint8_t* validity;  // bitmap base (NOT advanced) or nullptr
    int64_t bit_offset;       // bit index of row 0 in `validity`
    int32_t dtype;
    int32_t nan_nulls;        // NULL -> NaN, column read as f64
};

inline Col make_col(const VkColumn& c) {
    Col r;
    r.data = static_cast<const uint8_t*>(c.data) + c.offset * dtype_size(c.dtype);
    r.validity = c.validity;
    r.bit_offset = c.offset;
    r.dtype = c.dtype;
    r.nan_nulls = c.nulls_as_nan;
    return r;
}

__device__ __forceinline__ bool col_valid(const Col& c, int64_t i) {
    if (c.validity == nullptr) return true;
    int64_t b = c.bit_offset + i;
    return (c.validity[b >> 3] >> (b & 7)) & 1;
}

// Raw element widened to 64 bits: integers sign/zero-extend (== NextAsUInt64's
// static_cast<uint64_t>, array_iterators.h:215-217), floats bit-cast with the upper
// bytes zero (FloatArrayIter::floatToInt, array_iterators.h:243-247).
__device__ __forceinline__ uint64_t load_as_u64(const Col& c, int64_t i) {
    switch (c.dtype) {
        case VK_I8: return (uint64_t)(int64_t) reinterpret_cast<const int8_t*>(c.data)[i];
        case VK_I16: return (uint64_t)(int64_t) reinterpret_cast<const int16_t*>(c.data)[i];
        case VK_I32: return (uint64_t)(int64_t) reinterpret_cast<const int32_t*>(c.data)[i];
        case VK_I64: case VK_U64: case VK_F64:
            return reinterpret_cast<const uint64_t*>(c.data)[i];
        case VK_U8: case VK_BOOL8: return reinterpret_cast<const uint8_t*>(c.data)[i];
        case VK_U16: return reinterpret_cast<const uint16_t*>(c.data)[i];
        case VK_U32: case VK_F32: return reinterpret_cast<const uint32_t*>(c.data)[i];
        default: return 0;
    }
}
__device__ __forceinline__ int64_t load_as_i64(const Col& c, int64_t i) {
    return (int64_t) load_as_u64(c, i);  // valid for integer dtypes
}
__device__ __forceinline__ double load_as_f64_raw(const Col& c, int64_t i) {
    switch (c.dtype) {
        case VK_F64: return reinterpret_cast<const double*>(c.data)[i];
        case VK_F32: return (double) reinterpret_cast<const float*>(c.data)[i];
        case VK_U64: return (double) reinterpret_cast<const uint64_t*>(c.data)[i];
        case VK_U8: case VK_U16: case VK_U32: case VK_BOOL8:
            return (double) load_as_u64(c, i);
        default: return (double) (int64_t) load_as_u64(c, i);
    }
}
// NumPy-view semantics of the reference: NULL -> NaN (record_batch.py:100-125).
__device__ __forceinline__ double load_as_f64(const Col& c, int64_t i) {
    if (c.nan_nulls && !col_valid(c, i)) return __longlong_as_double(0x7ff8000000000000LL);
    return load_as_f64_raw(c, i);
}

__device__ __forceinline__ void store_from_u64(void* out, int dtype, int64_t i, uint64_t v) {
    switch (dtype) {
        case VK_I8: case VK_U8: case VK_BOOL8: reinterpret_cast<uint8_t*>(out)[i] = (uint8_t) v; break;
        case VK_I16: case VK_U16: reinterpret_cast<uint16_t*>(out)[i] = (uint16_t) v; break;
        case VK_I32: case VK_U32: case VK_F32: reinterpret_cast<uint32_t*>(out)[i] = (uint32_t) v; break;
        default: reinterpret_cast<uint64_t*>(out)[i] = v; break;
    }
}

// Streaming 16-byte load that does not pollute L1 (data is touched once).
__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ldg_stream8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// splitmix64 finaliser; also used as the hash mixer of the group-by tables.
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

// IEEE double -> order-preserving u64 (total order: -NaN < -inf < ... < +inf < +NaN).
__host__ __device__ __forceinline__ uint64_t f64_to_ordered(uint64_t bits) {
    return (bits & 0x8000000000000000ULL) ? ~bits : (bits | 0x8000000000000000ULL);
}
__host__ __device__ __forceinline__ uint64_t ordered_to_f64(uint64_t k) {
    return (k & 0x8000000000000000ULL) ? (k & 0x7fffffffffffffffULL) : ~k;
}

}  // namespace vk
