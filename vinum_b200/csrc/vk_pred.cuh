// Comparison domains and the in-register predicate shared by the compare, filter and
// fused filter->aggregate kernels.
//
// The reference evaluates comparisons with NumPy on zero-copy views
// (vinum/core/expressions.py:30-36), so the comparison domain follows NumPy's
// promotion: integer column vs Python int -> the column's integer type (compared here
// in 64 bits, which is value-preserving); anything involving a float -> float64,
// except float32 column vs Python float -> float32 (weak scalar); a column with NULLs
// is first materialised as float with NaN (vinum/arrow/record_batch.py:100-125).
#pragma once
#include "vk_common.cuh"
#include "vk_expr.cuh"
#include <cmath>

namespace vk {

enum Domain { DOM_I64 = 0, DOM_U64 = 1, DOM_F64 = 2, DOM_F32 = 3 };

template <int DOM> struct DomT;
template <> struct DomT<DOM_I64> { using type = int64_t; };
template <> struct DomT<DOM_U64> { using type = uint64_t; };
template <> struct DomT<DOM_F64> { using type = double; };
template <> struct DomT<DOM_F32> { using type = float; };

struct PredScalar {
    uint64_t bits;  // value in the domain's representation (f32 in the low 32 bits)
};

template <int DOM>
__device__ __forceinline__ typename DomT<DOM>::type bits_dom(uint64_t b) {
    if constexpr (DOM == DOM_I64) return (int64_t) b;
    else if constexpr (DOM == DOM_U64) return b;
    else if constexpr (DOM == DOM_F64) return __longlong_as_double((long long) b);
    else return __uint_as_float((uint32_t) b);
}
template <int DOM>
__device__ __forceinline__ typename DomT<DOM>::type scalar_dom(const PredScalar& s) {
    return bits_dom<DOM>(s.bits);
}

template <int DOM>
__device__ __forceinline__ typename DomT<DOM>::type load_dom(const Col& c, int64_t i) {
    if constexpr (DOM == DOM_I64) return (int64_t) load_as_u64(c, i);
    else if constexpr (DOM == DOM_U64) return load_as_u64(c, i);
    else if constexpr (DOM == DOM_F64) return load_as_f64(c, i);
    else {
        if (c.nan_nulls && !col_valid(c, i)) return __uint_as_float(0x7fc00000u);
        if (c.dtype == VK_F32) return reinterpret_cast<const float*>(c.data)[i];
        return (float) (int64_t) load_as_u64(c, i);  // int8/int16 operands of an f32 expression
    }
}

template <typename T>
__device__ __forceinline__ bool apply_cmp(int op, T x, T y) {
    switch (op) {
        case VK_EQ: return x == y;
        case VK_NE: return x != y;
        case VK_GT: return x > y;
        case VK_GE: return x >= y;
        case VK_LT: return x < y;
        default: return x <= y;
    }
}

// ---- host-side domain selection -------------------------------------------------
inline int pick_domain(const VkColumn& c, int scalar_dtype) {
    if (c.dtype == VK_F32) return (scalar_dtype == VK_F64 || scalar_dtype == VK_I64 || scalar_dtype == VK_U64) ? DOM_F32 : DOM_F32;
    if (c.dtype == VK_F64 || scalar_dtype == VK_F64 || c.nulls_as_nan) return DOM_F64;
    if (dtype_is_unsigned(c.dtype) && c.dtype == VK_U64) return DOM_U64;
    return DOM_I64;  // every other integer type is value-preserved in int64
}
inline int pick_domain_cols(const VkColumn& a, const VkColumn& b) {
    bool fa = dtype_is_float(a.dtype) || a.nulls_as_nan, fb = dtype_is_float(b.dtype) || b.nulls_as_nan;
    if (fa || fb) {
        // float32 survives only against float32 / int8 / int16 / uint8 / uint16 (NumPy promotion)
        auto small = [](const VkColumn& c) {
            return c.dtype == VK_F32 || (!c.nulls_as_nan && dtype_size(c.dtype) <= 2 && !dtype_is_float(c.dtype));
        };
        if ((a.dtype == VK_F32 || b.dtype == VK_F32) && small(a) && small(b)) return DOM_F32;
        return DOM_F64;
    }
    bool ua = a.dtype == VK_U64, ub = b.dtype == VK_U64;
    if (ua && ub) return DOM_U64;
    if (ua != ub) {
        // uint64 vs a signed type: NumPy compares exactly; neither 64-bit domain can
        bool other_signed = ua ? dtype_is_signed(b.dtype) : dtype_is_signed(a.dtype);
        if (other_signed) return -1;
        return DOM_U64;
    }
    return DOM_I64;
}

// Convert a user scalar into `dom`.  VK_ERR_UNSUPPORTED when it is not representable
// (the caller routes such expressions to the host path).
inline int convert_scalar(const VkScalar& s, int dom, PredScalar* out) {
    switch (dom) {
        case DOM_I64:
            if (s.dtype == VK_I64) { out->bits = (uint64_t) s.v.i; return VK_OK; }
            if (s.dtype == VK_U64 && s.v.u <= (uint64_t) INT64_MAX) { out->bits = s.v.u; return VK_OK; }
            return fail(VK_ERR_UNSUPPORTED, "scalar not representable as int64");
        case DOM_U64:
            if (s.dtype == VK_U64) { out->bits = s.v.u; return VK_OK; }
            if (s.dtype == VK_I64 && s.v.i >= 0) { out->bits = (uint64_t) s.v.i; return VK_OK; }
            return fail(VK_ERR_UNSUPPORTED, "scalar not representable as uint64");
        case DOM_F64: {
            double d = s.dtype == VK_F64 ? s.v.f : (s.dtype == VK_I64 ? (double) s.v.i : (double) s.v.u);
            memcpy(&out->bits, &d, 8);
            return VK_OK;
        }
        case DOM_F32: {
            double d = s.dtype == VK_F64 ? s.v.f : (s.dtype == VK_I64 ? (double) s.v.i : (double) s.v.u);
            float f = (float) d;
            uint32_t b;
            memcpy(&b, &f, 4);
            out->bits = b;
            return VK_OK;
        }
        default:
            return fail(VK_ERR_ARG, "bad comparison domain");
    }
}

// ---- device predicate -------------------------------------------------------------
struct Pred {
    int32_t kind;  // VkPredKind
    int32_t op;
    int32_t domain;
    int32_t _pad;
    const uint8_t* mask;
    Col col;
    PredScalar scalar;
    ECompare expr;   // VK_PRED_EXPR: <chain> <cmp> <chain> evaluated in registers (vk_expr.cuh)
};

// Predicate kernel specialisations.
enum PredKernelKind {
    PK_NONE = 0,     // no predicate: every row passes
    PK_MASK = 1,     // byte mask
    PK_F64_VEC = 2,  // float64 column, no validity, 16-byte aligned: one LDG.128 per row pair
    PK_I64_VEC = 3,  // int64 column, likewise
    PK_GENERIC = 4   // any dtype / domain / validity through load_dom
};

template <int DOM>
__device__ __forceinline__ bool pred_row_dom(const Pred& p, int64_t i) {
    return apply_cmp(p.op, load_dom<DOM>(p.col, i), scalar_dom<DOM>(p.scalar));
}
__device__ __forceinline__ bool pred_row_generic(const Pred& p, int64_t i) {
    if (p.kind == VK_PRED_EXPR) return compare_eval(p.expr, i);
    switch (p.domain) {
        case DOM_I64: return pred_row_dom<DOM_I64>(p, i);
        case DOM_U64: return pred_row_dom<DOM_U64>(p, i);
        case DOM_F32: return pred_row_dom<DOM_F32>(p, i);
        default: return pred_row_dom<DOM_F64>(p, i);
    }
}

// Flags of the row pair (r0, r0+1); r0 is even relative to the batch start.
template <int PK>
__device__ __forceinline__ void pred_pair(const Pred& p, int64_t r0, int64_t n, bool& f0, bool& f1) {
    f0 = f1 = false;
    if (r0 >= n) return;
    const bool two = r0 + 1 < n;
    if constexpr (PK == PK_NONE) {
        f0 = true;
        f1 = two;
    } else if constexpr (PK == PK_MASK) {
        if (two) {
            uint16_t m = *reinterpret_cast<const uint16_t*>(p.mask + r0);  // r0 even, cudaMalloc-aligned base
            f0 = m & 0xff;
            f1 = m >> 8;
        } else {
            f0 = p.mask[r0];
        }
    } else if constexpr (PK == PK_F64_VEC) {
        if (two) {
            uint4 q = ldg_stream16(p.col.data + r0 * 8);
            double a = __hiloint2double(q.y, q.x), b = __hiloint2double(q.w, q.z);
            double c = __longlong_as_double((long long) p.scalar.bits);
            f0 = apply_cmp(p.op, a, c);
            f1 = apply_cmp(p.op, b, c);
        } else {
            f0 = apply_cmp(p.op, reinterpret_cast<const double*>(p.col.data)[r0],
                           __longlong_as_double((long long) p.scalar.bits));
        }
    } else if constexpr (PK == PK_I64_VEC) {
        if (two) {
            uint4 q = ldg_stream16(p.col.data + r0 * 8);
            int64_t a = (int64_t) (((uint64_t) q.y << 32) | q.x), b = (int64_t) (((uint64_t) q.w << 32) | q.z);
            f0 = apply_cmp(p.op, a, (int64_t) p.scalar.bits);
            f1 = apply_cmp(p.op, b, (int64_t) p.scalar.bits);
        } else {
            f0 = apply_cmp(p.op, reinterpret_cast<const int64_t*>(p.col.data)[r0], (int64_t) p.scalar.bits);
        }
    } else {
        f0 = pred_row_generic(p, r0);
        if (two) f1 = pred_row_generic(p, r0 + 1);
    }
}

// Build the device predicate and pick the kernel specialisation.
inline int make_pred(const VkPredicate& in, int64_t n_rows, Pred* out, int* out_pk) {
    Pred p{};
    p.kind = in.kind;
    if (in.kind == VK_PRED_NONE) {
        *out = p;
        *out_pk = PK_NONE;
        return VK_OK;
    }
    if (in.kind == VK_PRED_MASK) {
        if (!in.mask) return fail(VK_ERR_ARG, "predicate: mask is NULL");
        if (reinterpret_cast<uintptr_t>(in.mask) & 1) return fail(VK_ERR_ARG, "predicate: mask must be 2-byte aligned");
        p.mask = in.mask;
        *out = p;
        *out_pk = PK_MASK;
        return VK_OK;
    }
    if (in.kind == VK_PRED_EXPR) {
        if (!in.expr) return fail(VK_ERR_ARG, "predicate: expr is NULL");
        const int rc = make_compare(*in.expr, n_rows, &p.expr);
        if (rc != VK_OK) return rc;
        *out = p;
        *out_pk = PK_GENERIC;
        return VK_OK;
    }
    if (in.kind != VK_PRED_CMP) return fail(VK_ERR_ARG, "predicate: unknown kind");
    if (in.op < VK_EQ || in.op > VK_LE) return fail(VK_ERR_ARG, "predicate: bad comparison op");
    if (!dtype_valid(in.column.dtype)) return fail(VK_ERR_ARG, "predicate: bad column dtype");
    if (in.column.length != n_rows) return fail(VK_ERR_ARG, "predicate: column length != n_rows");
    p.op = in.op;
    p.col = make_col(in.column);
    p.domain = pick_domain(in.column, in.scalar.dtype);
    int rc = convert_scalar(in.scalar, p.domain, &p.scalar);
    if (rc != VK_OK) return rc;
    const bool plain = in.column.validity == nullptr && !in.column.nulls_as_nan &&
                       (reinterpret_cast<uintptr_t>(p.col.data) & 15) == 0;
    if (plain && in.column.dtype == VK_F64 && p.domain == DOM_F64) *out_pk = PK_F64_VEC;
    else if (plain && in.column.dtype == VK_I64 && p.domain == DOM_I64) *out_pk = PK_I64_VEC;
    else *out_pk = PK_GENERIC;
    *out = p;
    return VK_OK;
}

}  // namespace vk
