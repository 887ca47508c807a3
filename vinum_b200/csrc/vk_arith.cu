// Arithmetic projection kernels: NumPy ufunc semantics of
// vinum/core/expressions.py:13-24 evaluated in one coalesced pass.  Compiled with
// -fmad=false: each result is a single correctly-rounded IEEE operation, as in NumPy.
#include "vk_common.cuh"
#include <cstdlib>
#include "vk_pred.cuh"

namespace vk {

struct Operand {
    Col col;
    int is_col;
    uint64_t bits;  // scalar converted to the compute class
};

struct ArithParams {
    int op;
    int out_dtype;
    Operand a, b;
};

enum ComputeClass { CC_I64 = 0, CC_U64 = 1, CC_F64 = 2, CC_F32 = 3 };

template <int CC> struct CCT;
template <> struct CCT<CC_I64> { using type = int64_t; };
template <> struct CCT<CC_U64> { using type = uint64_t; };
template <> struct CCT<CC_F64> { using type = double; };
template <> struct CCT<CC_F32> { using type = float; };

template <int CC>
__device__ __forceinline__ typename CCT<CC>::type load_operand(const Operand& o, int64_t i) {
    if constexpr (CC == CC_I64) return o.is_col ? (int64_t) load_as_u64(o.col, i) : (int64_t) o.bits;
    else if constexpr (CC == CC_U64) return o.is_col ? load_as_u64(o.col, i) : o.bits;
    else if constexpr (CC == CC_F64) return o.is_col ? load_as_f64(o.col, i) : __longlong_as_double((long long) o.bits);
    else return o.is_col ? load_dom<DOM_F32>(o.col, i) : __uint_as_float((uint32_t) o.bits);
}

// np.mod for floats (npy_divmod): C fmod, then move the result to the divisor's sign.
template <typename F>
__device__ __forceinline__ F np_fmod(F a, F b) {
    F m = fmod(a, b);
    if (b == (F) 0) return m;  // NaN
    if (m != (F) 0) {
        if ((b < (F) 0) != (m < (F) 0)) m += b;
    } else {
        m = copysign((F) 0, b);
    }
    return m;
}

template <int CC>
__device__ __forceinline__ typename CCT<CC>::type arith_apply(int op, typename CCT<CC>::type x,
                                                              typename CCT<CC>::type y) {
    using T = typename CCT<CC>::type;
    if constexpr (CC == CC_F64 || CC == CC_F32) {
        switch (op) {
            case VK_ADD: return x + y;
            case VK_SUB: return x - y;
            case VK_MUL: return x * y;
            case VK_DIV: return x / y;
            case VK_MOD: return np_fmod<T>(x, y);
            case VK_NEG: return -x;
            default: return x;
        }
    } else {
        switch (op) {
            case VK_ADD: return (T) ((uint64_t) x + (uint64_t) y);
            case VK_SUB: return (T) ((uint64_t) x - (uint64_t) y);
            case VK_MUL: return (T) ((uint64_t) x * (uint64_t) y);
            case VK_MOD: {
                if (y == 0) return 0;  // NumPy: integer x % 0 == 0 (with a warning)
                if constexpr (CC == CC_I64) {
                    if (y == -1) return 0;  // avoids INT64_MIN % -1 trap
                    int64_t m = x % y;
                    if (m != 0 && ((m < 0) != (y < 0))) m += y;  // floor-mod: sign of the divisor
                    return m;
                } else {
                    return x % y;
                }
            }
            case VK_BITAND: return x & y;
            case VK_BITOR: return x | y;
            case VK_BITXOR: return x ^ y;
            case VK_NEG: return (T) (0 - (uint64_t) x);
            case VK_BITNOT: return ~x;
            default: return x;
        }
    }
}

template <int CC>
__global__ void __launch_bounds__(256) arith_kernel(ArithParams p, int64_t n, void* __restrict__ out) {
    using T = typename CCT<CC>::type;
    int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T x = load_operand<CC>(p.a, i);
        T y = (p.op >= VK_NEG) ? x : load_operand<CC>(p.b, i);
        T r = arith_apply<CC>(p.op, x, y);
        if constexpr (CC == CC_F64) reinterpret_cast<double*>(out)[i] = r;
        else if constexpr (CC == CC_F32) reinterpret_cast<float*>(out)[i] = r;
        else {
            if (p.out_dtype == VK_BOOL8 && p.op == VK_BITNOT) r = (T) (((uint64_t) x) ^ 1ULL);
            store_from_u64(out, p.out_dtype, i, (uint64_t) r);
        }
    }
}

// 8-byte operands and result in the compute class's own representation (int64 / uint64 / float64
// columns without the NaN view, or scalars): U lane-contiguous row pairs per thread, 16-byte loads
// issued together, 16-byte stores.  arith_kernel goes through the per-row dtype dispatch with one
// 8-byte load per operand in flight.  Measured at 1e8 rows (profiles/r01_variants.md): f64 a + b 0.396 ms =
// 6.06 TB/s (0.93 of the peak) against 0.500 ms; VINUM_B200_ARITH_FAST=0 selects arith_kernel.
template <int CC, int U>
__global__ void __launch_bounds__(256) arith8_kernel(ArithParams p, int64_t n, void* __restrict__ out) {
    using T = typename CCT<CC>::type;
    auto from_bits = [](uint64_t b) -> T {
        if constexpr (CC == CC_F64) return __longlong_as_double((long long) b);
        else return (T) b;
    };
    auto to_bits = [](T v) -> uint64_t {
        if constexpr (CC == CC_F64) return (uint64_t) __double_as_longlong(v);
        else return (uint64_t) v;
    };
    const int64_t npairs = n >> 1;
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const bool unary = p.op >= VK_NEG;
    const bool acol = p.a.is_col, bcol = !unary && p.b.is_col;
    const uint32_t alo = (uint32_t) p.a.bits, ahi = (uint32_t) (p.a.bits >> 32);
    const uint32_t blo = (uint32_t) p.b.bits, bhi = (uint32_t) (p.b.bits >> 32);
    for (int64_t q0 = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; q0 < npairs; q0 += stride * U) {
        uint4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = q0 + u * stride;
            a[u] = make_uint4(alo, ahi, alo, ahi);
            b[u] = make_uint4(blo, bhi, blo, bhi);
            if (q < npairs) {
                if (acol) a[u] = ldg_stream16(p.a.col.data + q * 16);
                if (bcol) b[u] = ldg_stream16(p.b.col.data + q * 16);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t q = q0 + u * stride;
            if (q < npairs) {
                const T x0 = from_bits(((uint64_t) a[u].y << 32) | a[u].x), x1 = from_bits(((uint64_t) a[u].w << 32) | a[u].z);
                const T y0 = unary ? x0 : from_bits(((uint64_t) b[u].y << 32) | b[u].x);
                const T y1 = unary ? x1 : from_bits(((uint64_t) b[u].w << 32) | b[u].z);
                const uint64_t r0 = to_bits(arith_apply<CC>(p.op, x0, y0)), r1 = to_bits(arith_apply<CC>(p.op, x1, y1));
                reinterpret_cast<uint4*>(out)[q] = make_uint4((uint32_t) r0, (uint32_t) (r0 >> 32), (uint32_t) r1, (uint32_t) (r1 >> 32));
            }
        }
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // last, unpaired row
        const T x = load_operand<CC>(p.a, n - 1);
        const T y = unary ? x : load_operand<CC>(p.b, n - 1);
        reinterpret_cast<uint64_t*>(out)[n - 1] = to_bits(arith_apply<CC>(p.op, x, y));
    }
}

static int make_operand(const VkColumn* col, const VkScalar* sc, int cc, int64_t n, Operand* o, const char* side) {
    if (col) {
        if (!dtype_valid(col->dtype)) return fail(VK_ERR_ARG, std::string("vk_arith: bad dtype on ") + side);
        if (col->length != n) return fail(VK_ERR_ARG, std::string("vk_arith: length mismatch on ") + side);
        if ((cc == CC_I64 || cc == CC_U64) && (dtype_is_float(col->dtype) || col->nulls_as_nan))
            return fail(VK_ERR_ARG, "vk_arith: float operand with integer result dtype");
        o->col = make_col(*col);
        o->is_col = 1;
        return VK_OK;
    }
    o->is_col = 0;
    switch (cc) {
        case CC_I64: case CC_U64:
            if (sc->dtype == VK_F64) return fail(VK_ERR_ARG, "vk_arith: float scalar with integer result dtype");
            o->bits = sc->v.u;
            return VK_OK;
        case CC_F64: {
            double d = sc->dtype == VK_F64 ? sc->v.f : (sc->dtype == VK_I64 ? (double) sc->v.i : (double) sc->v.u);
            memcpy(&o->bits, &d, 8);
            return VK_OK;
        }
        default: {
            double d = sc->dtype == VK_F64 ? sc->v.f : (sc->dtype == VK_I64 ? (double) sc->v.i : (double) sc->v.u);
            float f = (float) d;
            uint32_t b;
            memcpy(&b, &f, 4);
            o->bits = b;
            return VK_OK;
        }
    }
}

}  // namespace vk

using namespace vk;

extern "C" int vk_arith(int op, const VkColumn* lhs_col, const VkScalar* lhs_scalar, const VkColumn* rhs_col,
                        const VkScalar* rhs_scalar, int64_t n_rows, int out_dtype, void* out, VkStream stream) {
    VK_REQUIRE(op >= VK_ADD && op <= VK_BITNOT, "vk_arith: bad op");
    VK_REQUIRE((lhs_col != nullptr) != (lhs_scalar != nullptr), "vk_arith: lhs must be exactly one of column/scalar");
    const bool unary = op >= VK_NEG;
    if (unary) VK_REQUIRE(!rhs_col && !rhs_scalar, "vk_arith: unary op takes no rhs");
    else VK_REQUIRE((rhs_col != nullptr) != (rhs_scalar != nullptr), "vk_arith: rhs must be exactly one of column/scalar");
    VK_REQUIRE(dtype_valid(out_dtype), "vk_arith: bad out_dtype");
    VK_REQUIRE(n_rows >= 0, "vk_arith: negative n_rows");
    if (n_rows == 0) return VK_OK;
    VK_REQUIRE(out, "vk_arith: out is NULL");
    int cc;
    if (out_dtype == VK_F64) cc = CC_F64;
    else if (out_dtype == VK_F32) cc = CC_F32;
    else if (out_dtype == VK_U64) cc = CC_U64;
    else cc = CC_I64;
    if (cc == CC_F64 || cc == CC_F32)
        VK_REQUIRE(op <= VK_MOD || op == VK_NEG, "vk_arith: bitwise op with float result dtype");
    if (cc == CC_I64 || cc == CC_U64) VK_REQUIRE(op != VK_DIV, "vk_arith: '/' is true division (float result)");
    ArithParams p{};
    p.op = op;
    p.out_dtype = out_dtype;
    int rc = make_operand(lhs_col, lhs_scalar, cc, n_rows, &p.a, "lhs");
    if (rc != VK_OK) return rc;
    if (!unary) {
        rc = make_operand(rhs_col, rhs_scalar, cc, n_rows, &p.b, "rhs");
        if (rc != VK_OK) return rc;
    }
    int64_t need = (n_rows + 255) / 256, cap = (int64_t) sm_count() * 8;
    int g = (int) (need < cap ? need : cap);
    cudaStream_t s = (cudaStream_t) stream;
    const int fast = (int) opt(OPT_ARITH_FAST);  // row pairs per thread of arith8_kernel (0: off; 4 measured: 6.06 vs 4.80 TB/s on f64 a + b)
    auto raw8 = [&](const Operand& o) {
        if (!o.is_col) return true;
        const bool same_class = cc == CC_F64 ? o.col.dtype == VK_F64 : (o.col.dtype == VK_I64 || o.col.dtype == VK_U64);
        return same_class && !o.col.nan_nulls && (reinterpret_cast<uintptr_t>(o.col.data) & 15) == 0;
    };
    const int out8 = cc == CC_F64 ? VK_F64 : (cc == CC_I64 ? VK_I64 : VK_U64);
    if (fast >= 2 && cc != CC_F32 && out_dtype == out8 && raw8(p.a) && (unary || raw8(p.b)) &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        int64_t need8 = (n_rows / 2 + 256 * 4 - 1) / (256 * 4);
        const int g8 = (int) (need8 < 1 ? 1 : (need8 < cap ? need8 : cap));
#define VK_ARITH8_GO(CC)                                                                 \
        do {                                                                             \
            if (fast >= 4) arith8_kernel<CC, 4><<<g8, 256, 0, s>>>(p, n_rows, out);      \
            else arith8_kernel<CC, 2><<<g8, 256, 0, s>>>(p, n_rows, out);                \
        } while (0)
        switch (cc) {
            case CC_I64: VK_ARITH8_GO(CC_I64); break;
            case CC_U64: VK_ARITH8_GO(CC_U64); break;
            default: VK_ARITH8_GO(CC_F64); break;
        }
#undef VK_ARITH8_GO
        VK_CHECK_LAUNCH("arith8_kernel");
        return VK_OK;
    }
    switch (cc) {
        case CC_I64: arith_kernel<CC_I64><<<g, 256, 0, s>>>(p, n_rows, out); break;
        case CC_U64: arith_kernel<CC_U64><<<g, 256, 0, s>>>(p, n_rows, out); break;
        case CC_F32: arith_kernel<CC_F32><<<g, 256, 0, s>>>(p, n_rows, out); break;
        default: arith_kernel<CC_F64><<<g, 256, 0, s>>>(p, n_rows, out); break;
    }
    VK_CHECK_LAUNCH("arith_kernel");
    return VK_OK;
}

// ---- fused expression chains (vk_expr.cuh) ------------------------------------------------
__global__ void __launch_bounds__(256) expr_eval_kernel(const __grid_constant__ vk::EChain c, int64_t n, uint64_t* __restrict__ out) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    for (int64_t i = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = vk::chain_eval(c, i);
}
// two rows per thread, two mask bytes per store (the mask buffer is 2-byte aligned)
__global__ void __launch_bounds__(256) expr_compare_kernel(const __grid_constant__ vk::ECompare e, int64_t n, uint8_t* __restrict__ out) {
    const int64_t stride = (int64_t) gridDim.x * blockDim.x;
    const int64_t pairs = n >> 1;
    for (int64_t q = (int64_t) blockIdx.x * blockDim.x + threadIdx.x; q < pairs; q += stride) {
        const uint16_t m = (uint16_t) vk::compare_eval(e, 2 * q) | ((uint16_t) vk::compare_eval(e, 2 * q + 1) << 8);
        reinterpret_cast<uint16_t*>(out)[q] = m;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[n - 1] = vk::compare_eval(e, n - 1);
}

extern "C" int vk_expr_eval(const VkExprChain* chain, int64_t n_rows, void* out, int32_t* out_dtype, VkStream stream) {
    VK_REQUIRE(chain && out_dtype && n_rows >= 0, "vk_expr_eval: bad argument");
    EChain c;
    int rc = make_chain(*chain, n_rows, &c);
    if (rc != VK_OK) return rc;
    *out_dtype = c.out_dom == EX_F64 ? VK_F64 : VK_I64;
    if (n_rows == 0) return VK_OK;
    VK_REQUIRE(out, "vk_expr_eval: out is NULL");
    int64_t need = (n_rows + 255) / 256, cap = (int64_t) sm_count() * 8;
    expr_eval_kernel<<<(unsigned) (need < cap ? need : cap), 256, 0, (cudaStream_t) stream>>>(c, n_rows, reinterpret_cast<uint64_t*>(out));
    VK_CHECK_LAUNCH("expr_eval_kernel");
    return VK_OK;
}

extern "C" int vk_expr_compare(const VkExprCompare* cmp, int64_t n_rows, uint8_t* out_mask, VkStream stream) {
    VK_REQUIRE(cmp && n_rows >= 0, "vk_expr_compare: bad argument");
    ECompare e;
    int rc = make_compare(*cmp, n_rows, &e);
    if (rc != VK_OK) return rc;
    if (n_rows == 0) return VK_OK;
    VK_REQUIRE(out_mask && (reinterpret_cast<uintptr_t>(out_mask) & 1) == 0, "vk_expr_compare: out_mask must be 2-byte aligned");
    int64_t need = (n_rows / 2 + 255) / 256, cap = (int64_t) sm_count() * 8;
    if (need < 1) need = 1;
    expr_compare_kernel<<<(unsigned) (need < cap ? need : cap), 256, 0, (cudaStream_t) stream>>>(e, n_rows, out_mask);
    VK_CHECK_LAUNCH("expr_compare_kernel");
    return VK_OK;
}
