// Hash group-by aggregate: shared declarations (table layout, accumulator kinds,
// device-side find-or-insert).  See vk_hashagg.cu for the kernels and DESIGN.md for
// the layout rationale.
#pragma once
#include "vk_common.cuh"
#include "vk_pred.cuh"

namespace vk {

// Accumulator kinds.  Every accumulator is one or two 64-bit lanes whose identity is
// all-zero bits, so a table is initialised with a single memset:
//   SUM_F64   lo = f64 bits, atomicAdd(double)                 (SumFunc<..,double_t>, agg_funcs.h:280-317)
//   SUM_I64   lo = wrapping 64-bit sum                         (SumFunc<Int8/16/32,...>)
//   SUM_I128  lo,hi = two's complement 128-bit sum             (SumOverflowFunc, agg_funcs.h:319-435; hugeint AddInPlace huge_int.cpp:281-300)
//   MAXORD    lo = max over an order-preserving u64 transform  (MinMaxFunc, agg_funcs.h:164-216); MIN stores the complement
enum AccKind {
    ACC_NONE = 0,   // COUNT(*): the per-group row counter is the result
    ACC_COUNT = 1,  // COUNT(col): row counter minus the per-function NULL counter
    ACC_SUM_F64 = 2,
    ACC_SUM_I64 = 3,
    ACC_SUM_I128 = 4,
    ACC_MAXORD = 5
};
// Order transforms for MAXORD
enum OrdKind { ORD_S64 = 0, ORD_U64 = 1, ORD_F64 = 2 };

struct FuncSpec {
    int32_t func;      // VkAggFunc
    int32_t in_dtype;  // VkDType
    int32_t acc;       // AccKind
    int32_t ord;       // OrdKind (MAXORD)
    int32_t is_min;    // MAXORD: store ~ord(v)
    int32_t in_unsigned;  // SUM_I128/SUM_I64: zero-extend (uint) instead of sign-extend
};

// Global (HBM/L2-resident) open-addressing table, structure-of-arrays.
//
// Two slot protocols:
//  * single-key tables (`single` = 1, the SingleNumericalHashAggregate case): the 64-bit
//    normalised key IS the slot tag.  A slot is claimed AND published by one 64-bit CAS
//    from GT_EMPTY to the key, so lookups never wait on another thread (wait-free
//    readers, lock-free inserts).  The key value GT_EMPTY itself and the NULL key live in
//    two dedicated slots behind the table (indices capacity and capacity+1), flagged in
//    `state[0..1]`.
//  * multi-key tables: a per-slot state word (0 empty, 1 being written, 2 ready) guards
//    the key tuple; a reader that meets a slot in state 1 backs off with __nanosleep so
//    the writer (possibly a lane of the same warp) is scheduled.
constexpr uint64_t GT_EMPTY = 0xFFFFFFFFFFFFFFFFULL;

struct GTable {
    int64_t capacity;      // power of two (hashed slots)
    int64_t max_groups;    // insertion limit (load factor bound)
    int32_t n_keys;
    int32_t n_funcs;
    int32_t single;        // 1: single-key protocol
    int32_t _pad;
    uint32_t* state;       // multi: [capacity] slot states; single: [2] flags of the special slots
    uint64_t* keys;        // multi: [n_keys][capacity]; single: [capacity + 2], GT_EMPTY = free
    uint32_t* knull;       // multi: [capacity] bit k set when key k is NULL; single: unused
    uint64_t* count_star;  // [slots]
    uint64_t* acc_lo[VK_AGG_MAX_FUNCS];
    uint64_t* acc_hi[VK_AGG_MAX_FUNCS];  // only for SUM_I128
    uint64_t* nnull[VK_AGG_MAX_FUNCS];   // NULL inputs seen per group
    unsigned long long* num_groups;      // device counter
};

__host__ __device__ __forceinline__ int64_t gt_total_slots(const GTable& t) { return t.capacity + (t.single ? 2 : 0); }

__device__ __forceinline__ uint64_t hash_keys(const uint64_t* kv, uint32_t nullmask, int n_keys) {
    uint64_t h = 0x243F6A8885A308D3ULL ^ nullmask;
    for (int k = 0; k < n_keys; ++k) h = splitmix64(h ^ ((nullmask >> k) & 1 ? 0 : kv[k]));
    return h;
}
__device__ __forceinline__ uint64_t hash_key1(uint64_t key) { return splitmix64(0x243F6A8885A308D3ULL ^ key); }

__device__ __forceinline__ uint32_t ld_state(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_state_release(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_key_relaxed(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Occupancy / key of slot s (finalize, rehash, export walk every slot).
__device__ __forceinline__ bool gt_slot_occupied(const GTable& t, int64_t s) {
    if (!t.single) return t.state[s] == 2;
    if (s < t.capacity) return t.keys[s] != GT_EMPTY;
    return t.state[s - t.capacity] != 0;
}
__device__ __forceinline__ void gt_slot_key(const GTable& t, int64_t s, uint64_t* kv, uint32_t* nullmask) {
    if (!t.single) {
        *nullmask = t.n_keys ? t.knull[s] : 0;
        for (int k = 0; k < t.n_keys; ++k) kv[k] = t.keys[(int64_t) k * t.capacity + s];
    } else if (s < t.capacity) {
        kv[0] = t.keys[s];
        *nullmask = 0;
    } else if (s == t.capacity) {
        kv[0] = GT_EMPTY;
        *nullmask = 0;
    } else {
        kv[0] = 0;
        *nullmask = 1;
    }
}

// Single-key protocol.  Returns the slot, or -1 when the table holds `limit` groups.
__device__ __forceinline__ int64_t gt1_find_or_insert(const GTable& t, uint64_t key, bool is_null, uint64_t hash,
                                                      int64_t limit) {
    if (is_null || key == GT_EMPTY) {
        const int which = is_null ? 1 : 0;
        if (ld_state(t.state + which) == 0u) {
            if (*reinterpret_cast<volatile unsigned long long*>(t.num_groups) >= (unsigned long long) limit) return -1;
            if (atomicCAS(t.state + which, 0u, 2u) == 0u) atomicAdd(t.num_groups, 1ULL);
        }
        return t.capacity + which;
    }
    const uint64_t mask = (uint64_t) t.capacity - 1;
    uint64_t slot = hash & mask;
    for (int64_t probes = 0; probes < t.capacity; ++probes) {
        uint64_t k = ld_key_relaxed(t.keys + slot);
        if (k == key) return (int64_t) slot;
        if (k == GT_EMPTY) {
            if (*reinterpret_cast<volatile unsigned long long*>(t.num_groups) >= (unsigned long long) limit) return -1;
            unsigned long long old = atomicCAS(reinterpret_cast<unsigned long long*>(t.keys + slot),
                                               (unsigned long long) GT_EMPTY, (unsigned long long) key);
            if (old == GT_EMPTY) {
                atomicAdd(t.num_groups, 1ULL);
                return (int64_t) slot;
            }
            if (old == key) return (int64_t) slot;
        }
        slot = (slot + 1) & mask;
    }
    return -1;
}

// Find the slot of (kv, nullmask) or claim a new one.  Returns -1 when the table has
// reached `limit` groups (the caller records the row for replay after a grow).
template <int NK>
__device__ __forceinline__ int64_t gt_find_or_insert(const GTable& t, const uint64_t* kv, uint32_t nullmask,
                                                     uint64_t hash, int64_t limit) {
    if (NK == 1 || t.single) return gt1_find_or_insert(t, kv[0], nullmask & 1u, hash, limit);
    const int n_keys = t.n_keys;
    const uint64_t mask = (uint64_t) t.capacity - 1;
    uint64_t slot = hash & mask;
    unsigned backoff = 8;
    for (int64_t probes = 0; probes < t.capacity;) {
        uint32_t s = ld_state(t.state + slot);
        if (s == 2) {
            bool eq = t.knull[slot] == nullmask;
            for (int k = 0; eq && k < n_keys; ++k)
                eq = ((nullmask >> k) & 1) || t.keys[(int64_t) k * t.capacity + slot] == kv[k];
            if (eq) return (int64_t) slot;
            slot = (slot + 1) & mask;
            ++probes;
        } else if (s == 0) {
            if (*reinterpret_cast<volatile unsigned long long*>(t.num_groups) >= (unsigned long long) limit)
                return -1;
            if (atomicCAS(t.state + slot, 0u, 1u) == 0u) {
                for (int k = 0; k < n_keys; ++k) t.keys[(int64_t) k * t.capacity + slot] = kv[k];
                t.knull[slot] = nullmask;
                st_state_release(t.state + slot, 2u);
                atomicAdd(t.num_groups, 1ULL);
                return (int64_t) slot;
            }
        } else {
            // being written by another thread: yield so that the writer (possibly a lane of
            // this warp) runs, then re-read the same slot
            __nanosleep(backoff);
            if (backoff < 256) backoff <<= 1;
        }
    }
    return -1;
}

// ---- accumulator updates on the global table -------------------------------------
__device__ __forceinline__ void acc_add_i128(uint64_t* lo, uint64_t* hi, uint64_t v_lo, uint64_t v_hi) {
    unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(lo), (unsigned long long) v_lo);
    uint64_t carry = (old + v_lo) < old ? 1ULL : 0ULL;
    uint64_t add_hi = v_hi + carry;
    if (add_hi) atomicAdd(reinterpret_cast<unsigned long long*>(hi), (unsigned long long) add_hi);
}

__device__ __forceinline__ uint64_t ord_transform(int ord, int is_min, uint64_t raw) {
    uint64_t o = ord == ORD_S64 ? (raw ^ 0x8000000000000000ULL) : (ord == ORD_U64 ? raw : f64_to_ordered(raw));
    return is_min ? ~o : o;
}
__device__ __forceinline__ uint64_t ord_inverse(int ord, int is_min, uint64_t stored) {
    uint64_t o = is_min ? ~stored : stored;
    return ord == ORD_S64 ? (o ^ 0x8000000000000000ULL) : (ord == ORD_U64 ? o : ordered_to_f64(o));
}

// Value of row i of `col` in the 64-bit representation the accumulator expects.
__device__ __forceinline__ uint64_t acc_load(const FuncSpec& f, const Col& col, int64_t i) {
    switch (f.acc) {
        case ACC_SUM_F64: return (uint64_t) __double_as_longlong(load_as_f64_raw(col, i));
        case ACC_MAXORD:
            if (f.ord == ORD_F64) return (uint64_t) __double_as_longlong(load_as_f64_raw(col, i));
            return load_as_u64(col, i);
        default: return load_as_u64(col, i);  // sign/zero-extended by dtype
    }
}

__device__ __forceinline__ void acc_update_global(const GTable& t, int fi, const FuncSpec& f, int64_t slot, uint64_t v) {
    switch (f.acc) {
        case ACC_SUM_F64:
            atomicAdd(reinterpret_cast<double*>(t.acc_lo[fi] + slot), __longlong_as_double((long long) v));
            break;
        case ACC_SUM_I64:
            atomicAdd(reinterpret_cast<unsigned long long*>(t.acc_lo[fi] + slot), (unsigned long long) v);
            break;
        case ACC_SUM_I128:
            acc_add_i128(t.acc_lo[fi] + slot, t.acc_hi[fi] + slot, v,
                         (!f.in_unsigned && (int64_t) v < 0) ? ~0ULL : 0ULL);
            break;
        case ACC_MAXORD:
            atomicMax(reinterpret_cast<unsigned long long*>(t.acc_lo[fi] + slot),
                      (unsigned long long) ord_transform(f.ord, f.is_min, v));
            break;
        default: break;
    }
}

// Rows that could not be inserted (table at its load limit) are appended here and replayed
// by the host after a grow.
struct ReplayList {
    uint32_t* rows;                 // row ids relative to the chunk start
    unsigned long long* count;      // appended so far
    unsigned long long* lost;       // rows that did not fit (must stay 0)
    unsigned long long* spilled;    // rows the fast kernel sent to the global path (statistics)
    unsigned long long* selected;   // rows that passed the fused predicate in the CTA tables (statistics, cumulative)
    uint64_t capacity;
};

__device__ __forceinline__ void replay_append(const ReplayList& l, int64_t row) {
    unsigned long long pos = atomicAdd(l.count, 1ULL);
    if (pos < l.capacity) l.rows[pos] = (uint32_t) row;
    else atomicAdd(l.lost, 1ULL);
}

}  // namespace vk
