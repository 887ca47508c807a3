/*
 * vinum_b200 -- C ABI of the B200-native (sm_100a) physical operators that replace
 * Vinum's CPU operators on its one data-parallel hot path:
 *   comparison/filter -> projection/arithmetic -> hash group-by aggregate -> sort
 * over Arrow column chunks resident in HBM.
 *
 * This header is the drop-in boundary.  Everything is `extern "C"`, plain pointers
 * and sizes; no C++/torch/Arrow types cross it.  Every entry point cites the
 * reference interface (path:line under the reference checkout) it replaces.
 *
 * Conventions
 *   - every function returns VK_OK (0) or a negative VK_ERR_* code; the message is
 *     available from vk_last_error() (thread-local).  Nothing aborts or throws
 *     (reference: `.ValueOrDie()` aborts at vinum/core/vinum_lib.cpp:62-63 and
 *     RAISE_ON_ARROW_FAILURE throws at vinum_cpp/src/common/util.h:4-11).
 *   - all `data` / `validity` / `mask` / output pointers are DEVICE pointers owned by
 *     the caller unless a parameter is explicitly named `host_*`.
 *   - a column is described exactly like an Arrow primitive array: values buffer,
 *     optional validity bitmap (LSB-first, 1 = valid), element offset, length.
 *   - one CUDA stream per call (`VkStream` is a cudaStream_t; NULL = legacy default
 *     stream).  Calls are asynchronous with respect to the host unless documented
 *     otherwise.  Objects are thread-compatible, not thread-safe (same as the
 *     reference's operators, which run under the GIL).
 */
#ifndef VINUM_B200_H
#define VINUM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VK_ABI_VERSION 1

#define VK_OK 0
#define VK_ERR_CUDA (-1)        /* a CUDA runtime call or kernel failed             */
#define VK_ERR_ARG (-2)         /* invalid argument                                  */
#define VK_ERR_UNSUPPORTED (-3) /* dtype / op combination not implemented on device  */
#define VK_ERR_OOM (-4)         /* device or pinned-host allocation failed           */
#define VK_ERR_STATE (-5)       /* object used in the wrong state                    */

typedef void* VkStream; /* cudaStream_t */
typedef void* VkEvent;  /* cudaEvent_t  */

/* Physical element types.  Arrow date32/time32 map to VK_I32; date64/time64/
 * timestamp/duration map to VK_I64 (the reference treats them the same way through
 * NumericArrayIter<T>, vinum_cpp/src/common/array_iterators.cpp:7-46). */
typedef enum VkDType {
    VK_I8 = 1, VK_I16 = 2, VK_I32 = 3, VK_I64 = 4,
    VK_U8 = 5, VK_U16 = 6, VK_U32 = 7, VK_U64 = 8,
    VK_F32 = 9, VK_F64 = 10,
    VK_BOOL8 = 11 /* one byte per row, 0/1: a NumPy `bool_` mask (expressions.py:30-36) */
} VkDType;

/* Device view of one Arrow primitive array (replaces arrow::NumericArray<T> as seen
 * by NumericArrayIter::SetArray, vinum_cpp/src/common/array_iterators.h:180-187). */
typedef struct VkColumn {
    const void* data;        /* values buffer (element 0 of the buffer, NOT of the slice) */
    const uint8_t* validity; /* Arrow validity bitmap or NULL when there are no nulls      */
    int64_t offset;          /* slice offset in elements (applies to data and validity)    */
    int64_t length;          /* number of rows                                             */
    int32_t dtype;           /* VkDType                                                    */
    int32_t nulls_as_nan;    /* 1: load through the reference's NumPy view semantics:
                                NULL -> NaN and the column is read as float64
                                (vinum/arrow/record_batch.py:100-125)                      */
} VkColumn;

typedef struct VkScalar {
    int32_t dtype; /* VK_I64, VK_U64 or VK_F64 */
    int32_t _pad;
    union { int64_t i; uint64_t u; double f; } v;
} VkScalar;

/* ---------------------------------------------------------------- runtime -- */
int vk_abi_version(void);
const char* vk_last_error(void);
int vk_device_count(int* out_n);
int vk_set_device(int device);
int vk_get_device(int* out_device);
int vk_device_info(int device, int* out_sm_count, int* out_cc_major, int* out_cc_minor,
                   uint64_t* out_total_bytes, uint64_t* out_free_bytes);
int vk_malloc(void** out_ptr, uint64_t bytes, VkStream stream);   /* stream-ordered pool */
int vk_free(void* ptr, VkStream stream);
int vk_host_alloc(void** out_ptr, uint64_t bytes);                /* pinned host memory  */
int vk_host_free(void* ptr);
int vk_host_register(void* ptr, uint64_t bytes);                  /* pin an Arrow buffer */
int vk_host_unregister(void* ptr);
int vk_memcpy_h2d(void* dst, const void* host_src, uint64_t bytes, VkStream stream);
/* Host -> device copy of PAGEABLE memory through a pool of pinned bounce buffers filled by worker
 * threads (vk_ingest.cu): returns when every piece is queued on `stream`, not when it has arrived.
 * vk_memcpy_h2d_auto picks: pinned / registered source or < 1 MB -> plain cudaMemcpyAsync, else staged.
 * This is how a host pyarrow.Table reaches the operators (the reference slices it in place,
 * vinum_cpp/src/operators/table_batch_reader.cpp:5-16). */
int vk_memcpy_h2d_staged(void* dst, const void* host_src, uint64_t bytes, VkStream stream);
int vk_memcpy_h2d_auto(void* dst, const void* host_src, uint64_t bytes, VkStream stream);
int vk_ingest_threads(void);   /* worker threads of the bounce-buffer pool (starts it) */
int vk_memcpy_d2h(void* host_dst, const void* src, uint64_t bytes, VkStream stream);
int vk_memcpy_d2d(void* dst, const void* src, uint64_t bytes, VkStream stream);
int vk_memset(void* dst, int byte, uint64_t bytes, VkStream stream);
int vk_stream_create(VkStream* out_stream);
int vk_stream_destroy(VkStream stream);
int vk_stream_sync(VkStream stream);
int vk_device_sync(void);
int vk_event_create(VkEvent* out_event);
int vk_event_destroy(VkEvent event);
int vk_event_record(VkEvent event, VkStream stream);
int vk_event_sync(VkEvent event);
int vk_stream_wait_event(VkStream stream, VkEvent event);  /* cudaStreamWaitEvent */
int vk_event_elapsed_ms(VkEvent start, VkEvent stop, float* out_ms);
/* number of kernels this library has launched since load (bench.py `gpu_launches`) */
uint64_t vk_launch_count(void);
/* Kernel-selection knobs (INTEGRATION.md "knobs"): compiled-in default, preset by the environment
 * variable VINUM_B200_<NAME> at first use, changed at run time here.  Unknown name -> VK_ERR_ARG.
 * They select between kernels that compute the same result; the reference has no counterpart. */
int vk_set_option(const char* name, int64_t value);
int vk_get_option(const char* name, int64_t* out_value);
int vk_reset_options(void);   /* forget every vk_set_option: back to environment / defaults */

/* ------------------------------------------------- synthetic table (8d) ---- */
/* Deterministic, shard-regenerable columns: value(row r, column c) is a pure
 * function of splitmix64(r*16 + c + seed*0x9E3779B97F4A7C15); the same formulas
 * are restated in NumPy in vinum_b200/datagen.py, so any row range is bit-identical
 * on host and device.  `kind` is a VkGenKind. */
typedef enum VkGenKind {
    VK_GEN_I0 = 0,  /* int64  u % 1000                       */
    VK_GEN_I1 = 1,  /* int64  uniform in [-2^40, 2^40)       */
    VK_GEN_I2 = 2,  /* int64  global row id                  */
    VK_GEN_I3 = 3,  /* int64  u % 1000000                    */
    VK_GEN_F0 = 4,  /* f64    uniform [0,1)                  */
    VK_GEN_F1 = 5,  /* f64    uniform [-1000,1000)           */
    VK_GEN_F2 = 6,  /* f64    sum of 4 uniforms - 2          */
    VK_GEN_F3 = 7,  /* f64    uniform [0,1e6)                */
    VK_GEN_K32 = 8  /* int32  u % 1000                       */
} VkGenKind;
int vk_datagen(int kind, uint64_t seed, int64_t row0, int64_t nrows, void* out, VkStream stream);

/* ------------------------------------------------ comparison -> mask (a1) -- */
/* Replaces the NumPy comparison lambdas of vinum/core/expressions.py:30-36.
 * out_mask: one byte per row (NumPy bool_).  NULL rows (when nulls_as_nan) compare
 * as NaN: every op is false except NE (record_batch.py:112-118 + IEEE). */
typedef enum VkCmpOp { VK_EQ = 0, VK_NE = 1, VK_GT = 2, VK_GE = 3, VK_LT = 4, VK_LE = 5 } VkCmpOp;
int vk_compare_scalar(const VkColumn* lhs, int op, const VkScalar* rhs, uint8_t* out_mask,
                      VkStream stream);
int vk_compare_columns(const VkColumn* lhs, int op, const VkColumn* rhs, uint8_t* out_mask,
                       VkStream stream);
/* BETWEEN / NOT BETWEEN, expressions.py:43-48: (x >= lo) & (x <= hi) / (x < lo) | (x > hi) */
int vk_between_scalar(const VkColumn* x, const VkScalar* lo, const VkScalar* hi, int negate,
                      uint8_t* out_mask, VkStream stream);
/* IN / NOT IN over a literal list, expressions.py:39-40 (np.isin).  `host_values`
 * are n_values scalars of one dtype, read on the host during the call. */
int vk_isin_scalars(const VkColumn* x, const VkScalar* host_values, int n_values, int negate,
                    uint8_t* out_mask, VkStream stream);

/* ------------------------------------------- boolean mask combine (a2) ----- */
/* Replaces pc.and_/or_/invert and pc.is_null/is_valid, expressions.py:27-29,37-38.
 * Masks are null-free byte masks on the numeric path. */
typedef enum VkMaskOp { VK_MASK_AND = 0, VK_MASK_OR = 1, VK_MASK_NOT = 2 } VkMaskOp;
int vk_mask_combine(int op, const uint8_t* a, const uint8_t* b /* NULL for NOT */, int64_t n,
                    uint8_t* out_mask, VkStream stream);
int vk_is_null(const VkColumn* x, int want_valid, uint8_t* out_mask, VkStream stream);
/* byte mask <-> Arrow bit-packed boolean (pa.array(np_bool), record_batch.py:86-87) */
int vk_mask_to_bits(const uint8_t* mask, int64_t n, uint8_t* out_bits, VkStream stream);
int vk_bits_to_mask(const uint8_t* bits, int64_t bit_offset, int64_t n, uint8_t* out_mask,
                    VkStream stream);

/* --------------------------------------------- filter / compaction (a4) ---- */
/* Replaces RecordBatch.filter -> pa.RecordBatch.filter(mask) (vinum/arrow/
 * record_batch.py:85-90, called from FilterOperator._kernel, vinum/core/
 * algebra.py:119-123): order-preserving stream compaction of EVERY column.
 *
 * The predicate is either a byte mask (pred_kind = VK_PRED_MASK) or a fused
 * `column <op> scalar` comparison evaluated in registers (VK_PRED_CMP), so the
 * mask is never materialised.  out_cols[i].data must have room for cols[i].length
 * elements; out_valid_bytes[i] (may be NULL when cols[i].validity is NULL)
 * receives one validity byte per selected row.  *out_rows (device int64) receives
 * the number of selected rows.  `scratch` is vk_filter_scratch_bytes(n) bytes. */
typedef enum VkPredKind { VK_PRED_NONE = 0, VK_PRED_MASK = 1, VK_PRED_CMP = 2, VK_PRED_EXPR = 3 } VkPredKind;
/* Fused expression chain (row a3): t0 <op1> t1 <op2> t2 ..., evaluated strictly left to right in
 * registers with NumPy's promotion per step (int64 op int64 wraps and stays int64 except `/`; anything
 * with a float64 is float64) -- what VectorizedExpression.evaluate (vinum/core/base.py:105-125) computes
 * node by node with one materialised array per node.  Terms: int64 / float64 columns without validity,
 * int64 / float64 scalars; ops: VK_ADD..VK_BITXOR (declared below). */
#define VK_EXPR_MAX_TERMS 4
typedef struct VkExprTerm {
    VkColumn column;       /* is_column != 0 */
    VkScalar scalar;       /* is_column == 0 */
    int32_t is_column;
    int32_t op;            /* VkArithOp: acc = acc <op> term; ignored for the first term */
} VkExprTerm;
typedef struct VkExprChain {
    int32_t n_terms;       /* 1 .. VK_EXPR_MAX_TERMS */
    int32_t _pad;
    VkExprTerm terms[VK_EXPR_MAX_TERMS];
} VkExprChain;
typedef struct VkExprCompare {
    VkExprChain lhs;
    VkExprChain rhs;
    int32_t op;            /* VkCmpOp */
    int32_t _pad;
} VkExprCompare;
typedef struct VkPredicate {
    int32_t kind;          /* VkPredKind                                   */
    int32_t op;            /* VkCmpOp (VK_PRED_CMP)                        */
    const uint8_t* mask;   /* VK_PRED_MASK: one byte per row               */
    VkColumn column;       /* VK_PRED_CMP: left-hand side                  */
    VkScalar scalar;       /* VK_PRED_CMP: right-hand side                 */
    const VkExprCompare* expr;  /* VK_PRED_EXPR: <chain> <cmp> <chain>, e.g. WHERE a * 10 > b */
} VkPredicate;
uint64_t vk_filter_scratch_bytes(int64_t n_rows);
int vk_filter(const VkPredicate* pred, int64_t n_rows, const VkColumn* cols, int n_cols,
              void* const* out_data, uint8_t* const* out_valid_bytes, int64_t* out_rows,
              void* scratch, VkStream stream);

/* ------------------------------------------ arithmetic projection (a5) ----- */
/* Replaces the NumPy ufuncs of vinum/core/expressions.py:13-24.  NumPy semantics:
 * `/` is true division (float64 out), `%` is floor-mod, integers wrap.  The caller
 * passes the NumPy result dtype as out_dtype.  Compiled with -fmad=false so float
 * results are bit-identical to NumPy's single IEEE operations. */
typedef enum VkArithOp {
    VK_ADD = 0, VK_SUB = 1, VK_MUL = 2, VK_DIV = 3, VK_MOD = 4,
    VK_BITAND = 5, VK_BITOR = 6, VK_BITXOR = 7,
    VK_NEG = 8, VK_BITNOT = 9
} VkArithOp;
/* lhs/rhs: exactly one of (column, scalar) is non-NULL per side; for unary ops rhs
 * side is entirely NULL. */
int vk_arith(int op, const VkColumn* lhs_col, const VkScalar* lhs_scalar, const VkColumn* rhs_col,
             const VkScalar* rhs_scalar, int64_t n_rows, int out_dtype, void* out, VkStream stream);
/* A whole chain in one pass (`SELECT a * 10 + b`): out = n_rows int64 or float64 values, *out_dtype tells
 * which (VK_I64 / VK_F64, decided by the promotion rules above).  vk_expr_compare: <chain> <cmp> <chain>
 * -> one mask byte per row. */
int vk_expr_eval(const VkExprChain* chain, int64_t n_rows, void* out, int32_t* out_dtype, VkStream stream);
int vk_expr_compare(const VkExprCompare* cmp, int64_t n_rows, uint8_t* out_mask, VkStream stream);

/* ------------------------------------------------ hash aggregate (a8-a14) -- */
/* Replaces BaseAggregate / SingleNumericalHashAggregate / MultiNumericalHashAggregate
 * / OneGroupAggregate (vinum_cpp/src/operators/aggregate/*.cpp) and the aggregate
 * functions of agg_funcs.h.  A VkAgg is the streaming state of ONE operator: created
 * once, fed every batch with vk_agg_update (== BaseAggregate::Next,
 * base_aggregate.cpp:23-45), asked once for the result (== Result, :47-68).
 *
 * Keys are normalised exactly like NextAsUInt64 (array_iterators.h:215-217,239-248):
 * integers sign-extend to u64, floats bit-cast; a NULL key is its own group. */
typedef enum VkAggFunc {
    VK_AGG_COUNT_STAR = 0, VK_AGG_COUNT = 1, VK_AGG_MIN = 2, VK_AGG_MAX = 3,
    VK_AGG_SUM = 4, VK_AGG_AVG = 5
} VkAggFunc;
typedef struct VkAgg VkAgg;
#define VK_AGG_MAX_KEYS 8
#define VK_AGG_MAX_FUNCS 16
/* key_dtypes[n_keys]; funcs[n_funcs] with func_in_dtypes[i] the input column dtype
 * (ignored for COUNT_STAR).  n_keys == 0 is the un-grouped OneGroupAggregate. */
int vk_agg_create(VkAgg** out, int n_keys, const int32_t* key_dtypes, int n_funcs,
                  const int32_t* funcs, const int32_t* func_in_dtypes, int64_t expected_groups);
int vk_agg_destroy(VkAgg* agg);
/* One batch.  keys[n_keys], values[n_funcs] (values[i] ignored for COUNT_STAR; may
 * alias).  `pred` may be NULL (no filter) -- a non-NULL predicate fuses the WHERE
 * clause into the aggregation pass (north-star filter->hash-agg pipeline). */
int vk_agg_update(VkAgg* agg, const VkPredicate* pred, int64_t n_rows, const VkColumn* keys,
                  const VkColumn* values, VkStream stream);
/* Number of groups so far (synchronises the stream). */
int vk_agg_num_groups(VkAgg* agg, int64_t* out_groups, VkStream stream);
/* Finalised, dense result (device buffers, each with room for num_groups rows):
 *   out_keys[k]        : u64 normalised key values
 *   out_key_valid[k]   : one byte per group, 0 for the NULL key
 *   out_count_star     : u64 rows per group
 *   out_vals_lo/hi[f]  : raw 64-bit result lanes per function:
 *       COUNT*          lo = count (u64)
 *       MIN/MAX         lo = value in its input type widened to 64 bits / f64 bits
 *       SUM int8-32     lo = int64/uint64 wrap sum        SUM float  lo = f64 bits
 *       SUM (u)int64    lo,hi = 128-bit two's complement sum
 *       AVG             lo = f64 bits of the average computed with the reference's
 *                       formulas (agg_funcs.h:519-540), or f32 value widened to f64
 *                       for (u)int8/16 inputs (agg_func_factory.cpp:179-196)
 *   out_vals_valid[f]  : one byte per group, 0 when the group had no valid input
 * Group order is unspecified (as in the reference) except that the NULL
 * single-key group is last (single_numerical_hash_aggregate.cpp:54-60). */
int vk_agg_result(VkAgg* agg, int64_t num_groups, uint64_t* const* out_keys,
                  uint8_t* const* out_key_valid, uint64_t* out_count_star,
                  uint64_t* const* out_vals_lo, uint64_t* const* out_vals_hi,
                  uint8_t* const* out_vals_valid, VkStream stream);
/* Multi-GPU exchange: export this rank's partial groups destined to
 * rank `dest` of `n_ranks` (hash(key) mod n_ranks) as a flat u64 record stream and
 * merge a received stream into the local table.  Record layout: vk_agg_record_words. */
int vk_agg_record_words(VkAgg* agg, int* out_words);
int vk_agg_partition_counts(VkAgg* agg, int n_ranks, int64_t* out_counts_dev, VkStream stream);
int vk_agg_export_partials(VkAgg* agg, int n_ranks, const int64_t* offsets_dev /* n_ranks */,
                           uint64_t* out_records, VkStream stream);
int vk_agg_merge_partials(VkAgg* agg, const uint64_t* records, int64_t n_records, VkStream stream);
/* Finalise into ONE device block (one copy + one synchronisation for the caller).  u64 words, C = capacity:
 *   [0] groups in the table (may exceed C: retry with a larger block)  [1] reserved
 *   keys[n_keys][C] | count_star[C] | lo[n_funcs][C] | hi[n_funcs][C] | bytes key_valid[n_keys][C] valid[n_funcs][C]
 * Same values as vk_agg_result (BaseAggregate::Result, base_aggregate.cpp:47-68). */
uint64_t vk_agg_result_packed_bytes(int n_keys, int n_funcs, int64_t capacity);
int vk_agg_result_packed(VkAgg* agg, int64_t capacity, void* out_block, VkStream stream);

/* ---- peer exchange: low-cardinality group-by across the GPUs of one box without a collective ----
 * Every rank owns a window of device memory (cudaMalloc) that its peers map through cudaIpc handles
 * (exchanged once by the host, e.g. torch.distributed.all_gather_object) and write into over NVLink.
 * vk_agg_peer_send: this rank's partial groups -> its slot in the owner's window + release flag, then
 * waits (on the device) for the owner's acknowledgement.  vk_agg_peer_merge (owner): acquires every
 * rank's flag, folds the records into the owner's own table, acknowledges.  `epoch` = 1, 2, 3 ... is
 * the query number, the same on every rank.  After the stream has drained, *vk_peer_decision_ptr(epoch)
 * (a device address; copy it back) is 1 = merged, 2 = a rank had more than cap_groups partial groups,
 * 3 = the owner's table must grow, 4 = a peer did not answer within 4 s; on 2 / 3 every rank falls back
 * to the all-to-all repartition (vk_agg_partition_counts / _export_partials / _merge_partials).
 * The reference has no counterpart (single-threaded, vinum/executor/executor.py:24-31). */
typedef struct VkPeer VkPeer;
int vk_peer_create(VkPeer** out, int rank, int world, int64_t cap_groups, int max_record_words);
int vk_peer_destroy(VkPeer* peer);
int vk_peer_handle(VkPeer* peer, void* out_handle64 /* 64 bytes */);
int vk_peer_open(VkPeer* peer, int peer_rank, const void* handle64);
int vk_peer_attach_local(VkPeer* peer, int peer_rank, VkPeer* other /* same process */);
void* vk_peer_decision_ptr(VkPeer* peer, uint64_t epoch);
int vk_agg_peer_send(VkAgg* agg, VkPeer* peer, int owner_rank, uint64_t epoch, VkStream stream);
int vk_agg_peer_merge(VkAgg* agg, VkPeer* peer, uint64_t epoch, VkStream stream);

/* introspection for tests/bench: which kernel path the last update used
 * (0 = none, 1 = shared-memory table, 2 = global table, 3 = one-group reduction,
 * 4 = partitioned: scatter into buckets that are slices of the global table + update slice by slice) */
int vk_agg_last_path(VkAgg* agg);
/* The estimate behind the sizing of the global table (host arithmetic, no device work): number of distinct keys G
 * if `groups` distinct keys were seen among `selected_rows` rows drawn evenly, g = G (1 - exp(-n / G)); 0 when
 * groups >= selected_rows (no upper bound).  The reference sizes nothing in advance (std::unordered_map grows,
 * single_numerical_hash_aggregate.cpp:15-46); exported for tests. */
double vk_agg_estimate_groups(double selected_rows, double groups);
/* Optional per-launch kernel timing (CUDA events recorded on the launch stream, resolved
 * lazily: no extra synchronisation).  vk_agg_profile_read returns, for `path` (1/2 as
 * above), the summed kernel milliseconds, the number of launches and the rows they
 * processed since profiling was enabled; it waits for the recorded events. */
int vk_agg_profile(VkAgg* agg, int enable);
int vk_agg_profile_read(VkAgg* agg, int path, double* out_ms, int64_t* out_launches, int64_t* out_rows);

/* --------------------------------------------------------- sort (a15-a16) -- */
/* Replaces Sort::Sorted (vinum_cpp/src/operators/sort/sort.cpp:15-63):
 * arrow::compute::SortIndices (stable; NaN then NULL last in both directions;
 * -0.0 == +0.0) followed by Take of every column. */
typedef enum VkSortOrder { VK_ASC = 0, VK_DESC = 1 } VkSortOrder;
uint64_t vk_sort_scratch_bytes(int64_t n_rows);
/* out_indices: n_rows int64 row ids (the permutation SortIndices returns). */
int vk_sort_indices(const VkColumn* keys, const int32_t* orders, int n_keys, int64_t n_rows,
                    int64_t* out_indices, void* scratch, VkStream stream);
/* Same, and also Take of the FIRST key column itself (sort.cpp:40-48 gathers every column, the sort
 * key included): out_key0_sorted receives n_rows 8-byte values, written by the last radix pass from the
 * sorted codes instead of a random gather.  keys[0] must be a plain int64 / uint64 / float64 column
 * without validity; values are bit-identical to Take (NaN payloads and the sign of zero included). */
int vk_sort_indices_keys(const VkColumn* keys, const int32_t* orders, int n_keys, int64_t n_rows,
                         int64_t* out_indices, void* out_key0_sorted, void* scratch, VkStream stream);
/* Take: out[i] = col[indices[i]]; out_valid_bytes may be NULL when no validity. */
int vk_take(const VkColumn* col, const int64_t* indices, int64_t n_indices, void* out,
            uint8_t* out_valid_bytes, VkStream stream);

/* ORDER BY key <order> LIMIT k without the full sort (SliceOperator below SortOperator,
 * vinum/core/algebra.py:204-247 + vinum/planner/planner.py:478-501): an MSD radix select over
 * the key's order-preserving code finds the byte prefix that holds the k-th row, then the ids of
 * every row at or before that prefix are written to out_rows (a superset of the first k rows of
 * the stable sort, in no particular order; the caller sorts just these).  *out_count (HOST) is
 * their number, or -1 when the select does not pay (k > n_rows / 8, an integer key with a
 * validity bitmap, or more than max_candidates candidates) and the caller should sort
 * everything.  Reads the key column once per prefix byte examined (at most 8 times) plus once to
 * collect; synchronises `stream`.  scratch: vk_topk_scratch_bytes() device bytes. */
uint64_t vk_topk_scratch_bytes(void);
int vk_topk_candidates(const VkColumn* key, int32_t order, int64_t n_rows, int64_t k, int64_t max_candidates,
                       int64_t* out_rows, int64_t* out_count, void* scratch, VkStream stream);

#ifdef __cplusplus
}
#endif
#endif /* VINUM_B200_H */
