// Fused expression chains.
//
// The reference evaluates an expression tree node by node (VectorizedExpression.evaluate,
// vinum/core/base.py:105-125): `WHERE a * 10 > b` is three NumPy calls and two materialised 8 B/row
// intermediates.  Here a left-deep chain  t0 <op1> t1 <op2> t2 ...  over plain int64 / float64 columns
// and scalars (and a comparison of two such chains) is ONE pass in registers: every step is the same
// single IEEE / wrapping integer operation NumPy performs (-fmad=false), with NumPy's promotion applied
// step by step -- int64 op int64 stays int64 (wraps) except `/`, anything with a float64 is float64 --
// so results are bit-identical to the node-by-node evaluation.  The chain is also a predicate kind
// (VK_PRED_EXPR): vk_filter and vk_agg_update evaluate it in their own row loops, nothing is
// materialised at all.
#pragma once
#include "vk_common.cuh"
#include <cmath>

namespace vk {

constexpr int EX_I64 = 0, EX_F64 = 1;

struct ETerm {
    const uint8_t* data;   // column values (8-byte elements, offset folded in) or nullptr for a scalar
    uint64_t bits;         // scalar in its own domain
    uint8_t op;            // VkArithOp: acc = acc <op> term (unused for the first term)
    uint8_t tdom;          // domain of the term itself
    uint8_t cdom;          // domain the step computes in (after promotion)
    uint8_t _pad;
};
struct EChain {
    int32_t n;
    int32_t out_dom;
    ETerm t[VK_EXPR_MAX_TERMS];
};
struct ECompare {
    EChain lhs, rhs;
    int32_t op;            // VkCmpOp
    int32_t dom;           // EX_I64: both sides int64, else float64
};

// np.mod for floats (npy_divmod): C fmod, then move the result to the divisor's sign.
__device__ __forceinline__ double ex_fmod(double a, double b) {
    double m = fmod(a, b);
    if (b == 0.0) return m;  // NaN
    if (m != 0.0) {
        if ((b < 0.0) != (m < 0.0)) m += b;
    } else {
        m = copysign(0.0, b);
    }
    return m;
}
__device__ __forceinline__ double ex_f64(int op, double x, double y) {
    switch (op) {
        case VK_ADD: return x + y;
        case VK_SUB: return x - y;
        case VK_MUL: return x * y;
        case VK_DIV: return x / y;
        default: return ex_fmod(x, y);   // VK_MOD
    }
}
__device__ __forceinline__ int64_t ex_i64(int op, int64_t x, int64_t y) {
    switch (op) {
        case VK_ADD: return (int64_t) ((uint64_t) x + (uint64_t) y);
        case VK_SUB: return (int64_t) ((uint64_t) x - (uint64_t) y);
        case VK_MUL: return (int64_t) ((uint64_t) x * (uint64_t) y);
        case VK_MOD: {
            if (y == 0 || y == -1) return 0;   // NumPy: x % 0 == 0; INT64_MIN % -1 would trap
            int64_t m = x % y;
            if (m != 0 && ((m < 0) != (y < 0))) m += y;  // floor-mod: sign of the divisor
            return m;
        }
        case VK_BITAND: return x & y;
        case VK_BITOR: return x | y;
        default: return x ^ y;   // VK_BITXOR
    }
}
__device__ __forceinline__ double ex_as_f64(uint64_t bits, int dom) {
    return dom == EX_F64 ? __longlong_as_double((long long) bits) : (double) (int64_t) bits;
}
__device__ __forceinline__ uint64_t ex_term(const ETerm& t, int64_t i) {
    return t.data != nullptr ? reinterpret_cast<const uint64_t*>(t.data)[i] : t.bits;
}
// Value of the chain at row i, in c.out_dom.  Every branch is uniform across the grid.
__device__ __forceinline__ uint64_t chain_eval(const EChain& c, int64_t i) {
    uint64_t acc = ex_term(c.t[0], i);
    int dom = c.t[0].tdom;
#pragma unroll
    for (int k = 1; k < VK_EXPR_MAX_TERMS; ++k) {
        if (k < c.n) {
            const ETerm& t = c.t[k];
            const uint64_t y = ex_term(t, i);
            if (t.cdom == EX_F64) {
                acc = (uint64_t) __double_as_longlong(ex_f64(t.op, ex_as_f64(acc, dom), ex_as_f64(y, t.tdom)));
                dom = EX_F64;
            } else {
                acc = (uint64_t) ex_i64(t.op, (int64_t) acc, (int64_t) y);
            }
        }
    }
    return acc;
}
__device__ __forceinline__ bool ex_cmp_i64(int op, int64_t x, int64_t y) {
    switch (op) {
        case VK_EQ: return x == y;
        case VK_NE: return x != y;
        case VK_GT: return x > y;
        case VK_GE: return x >= y;
        case VK_LT: return x < y;
        default: return x <= y;
    }
}
__device__ __forceinline__ bool ex_cmp_f64(int op, double x, double y) {
    switch (op) {
        case VK_EQ: return x == y;
        case VK_NE: return x != y;
        case VK_GT: return x > y;
        case VK_GE: return x >= y;
        case VK_LT: return x < y;
        default: return x <= y;
    }
}
__device__ __forceinline__ bool compare_eval(const ECompare& e, int64_t i) {
    const uint64_t a = chain_eval(e.lhs, i), b = chain_eval(e.rhs, i);
    if (e.dom == EX_I64) return ex_cmp_i64(e.op, (int64_t) a, (int64_t) b);
    return ex_cmp_f64(e.op, ex_as_f64(a, e.lhs.out_dom), ex_as_f64(b, e.rhs.out_dom));
}

// ---- host: public structs -> device form (validation + NumPy promotion) -------------------
inline int make_chain(const VkExprChain& in, int64_t n_rows, EChain* out) {
    if (in.n_terms < 1 || in.n_terms > VK_EXPR_MAX_TERMS) return fail(VK_ERR_ARG, "expression chain: 1..4 terms");
    EChain c{};
    c.n = in.n_terms;
    int dom = EX_I64;
    for (int k = 0; k < in.n_terms; ++k) {
        const VkExprTerm& s = in.terms[k];
        ETerm t{};
        if (s.is_column) {
            const VkColumn& col = s.column;
            if (col.validity != nullptr || col.nulls_as_nan) return fail(VK_ERR_UNSUPPORTED, "expression chain: column with NULLs");
            if (col.dtype != VK_I64 && col.dtype != VK_F64) return fail(VK_ERR_UNSUPPORTED, "expression chain: int64 / float64 columns only");
            if (col.length != n_rows) return fail(VK_ERR_ARG, "expression chain: column length != n_rows");
            if (!col.data) return fail(VK_ERR_ARG, "expression chain: NULL column data");
            t.data = static_cast<const uint8_t*>(col.data) + col.offset * 8;
            t.tdom = col.dtype == VK_F64 ? EX_F64 : EX_I64;
        } else {
            if (s.scalar.dtype == VK_F64) { t.tdom = EX_F64; memcpy(&t.bits, &s.scalar.v.f, 8); }
            else if (s.scalar.dtype == VK_I64) { t.tdom = EX_I64; t.bits = (uint64_t) s.scalar.v.i; }
            else return fail(VK_ERR_UNSUPPORTED, "expression chain: int64 / float64 scalars only");
        }
        if (k == 0) {
            dom = t.tdom;
        } else {
            if (s.op < VK_ADD || s.op > VK_BITXOR) return fail(VK_ERR_ARG, "expression chain: binary arithmetic ops only");
            const bool any_float = dom == EX_F64 || t.tdom == EX_F64;
            if (s.op >= VK_BITAND && any_float) return fail(VK_ERR_UNSUPPORTED, "expression chain: bitwise op on a float");
            t.op = (uint8_t) s.op;
            t.cdom = (any_float || s.op == VK_DIV) ? EX_F64 : EX_I64;   // `/` is true division: float64 even for ints
            dom = t.cdom;
        }
        c.t[k] = t;
    }
    c.out_dom = dom;
    *out = c;
    return VK_OK;
}
inline int make_compare(const VkExprCompare& in, int64_t n_rows, ECompare* out) {
    ECompare e{};
    int rc = make_chain(in.lhs, n_rows, &e.lhs);
    if (rc != VK_OK) return rc;
    rc = make_chain(in.rhs, n_rows, &e.rhs);
    if (rc != VK_OK) return rc;
    if (in.op < VK_EQ || in.op > VK_LE) return fail(VK_ERR_ARG, "expression compare: bad comparison op");
    e.op = in.op;
    e.dom = (e.lhs.out_dom == EX_I64 && e.rhs.out_dom == EX_I64) ? EX_I64 : EX_F64;
    *out = e;
    return VK_OK;
}

}  // namespace vk
