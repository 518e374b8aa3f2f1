/*
 * dfgpu.h — C ABI of the B200-native (sm_100a) engine for DataFusion 0.6.0's Arrow-batch hot path.
 *
 * This is the drop-in boundary.  The reference (andygrove/datafusion-archive, Rust) has no FFI; its
 * operator "plugin API" is the `Relation` trait plus the `Expr`/`LogicalPlan` IR.  Each entry point
 * below names the reference interface it replaces (file:line relative to the reference root).
 * A Rust `extern "C"` block binding these symbols, and the `impl Relation for Gpu*Relation` that
 * calls them, is shown in INTEGRATION.md.
 *
 * Conventions
 *   - Every function returns 0 on success, a DFGPU_ERR_* code otherwise; the message is available
 *     from dfgpu_last_error() (thread-local, valid until the next call on the same thread).
 *     The shim maps codes onto `ExecutionError` variants (src/execution/error.rs:51-60).
 *   - No CPU fallback exists anywhere behind this ABI: an unsupported dtype / operator returns
 *     DFGPU_ERR_NOT_IMPLEMENTED; a missing GPU returns DFGPU_ERR_CUDA.
 *   - Input buffers are BORROWED for the duration of a call (Arrow `ArrayData` views: values
 *     pointer, length, offset, optional LSB-first validity bitmap).  Outputs are copied into
 *     caller-allocated buffers after a shape query — no cross-allocator frees.
 *   - One host thread per dfgpu_ctx (the reference is `Rc<RefCell<..>>`, i.e. !Send:
 *     src/execution/context.rs:34).  One ctx drives one GPU; multi-GPU = one process (or ctx) per
 *     GPU with row-range partitioning, joined by dfgpu_comm_init() for the partial-aggregate merge.
 */
#ifndef DFGPU_H
#define DFGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFGPU_ABI_VERSION 2

/* ---- error codes (→ ExecutionError variants, src/execution/error.rs:51-60) ---- */
enum {
  DFGPU_OK = 0,
  DFGPU_ERR_GENERAL = 1,         /* ExecutionError::General          */
  DFGPU_ERR_EXECUTION = 2,       /* ExecutionError::ExecutionError   */
  DFGPU_ERR_NOT_IMPLEMENTED = 3, /* ExecutionError::NotImplemented   */
  DFGPU_ERR_INVALID_COLUMN = 4,  /* ExecutionError::InvalidColumn    */
  DFGPU_ERR_INTERNAL = 5,        /* ExecutionError::InternalError    */
  DFGPU_ERR_ARROW = 6,           /* ExecutionError::ArrowError (e.g. DivideByZero, length mismatch) */
  DFGPU_ERR_CUDA = 7,            /* device / driver / NCCL failure   */
  DFGPU_ERR_OOM = 8
};

/* ---- Arrow data types on the path (arrow::datatypes::DataType) ---- */
enum {
  DFGPU_BOOL = 1, /* bit-packed, LSB first */
  DFGPU_INT8 = 2,
  DFGPU_INT16 = 3,
  DFGPU_INT32 = 4,
  DFGPU_INT64 = 5,
  DFGPU_UINT8 = 6,
  DFGPU_UINT16 = 7,
  DFGPU_UINT32 = 8,
  DFGPU_UINT64 = 9,
  DFGPU_FLOAT32 = 10,
  DFGPU_FLOAT64 = 11,
  DFGPU_UTF8 = 12 /* arrow 0.12 BinaryArray: i32 offsets (len+1) + u8 data */
};

/* Borrowed view of one Arrow array (arrow `ArrayData`): element i lives at values[(offset+i)],
 * validity bit i at validity[(offset+i)>>3] >> ((offset+i)&7) & 1 (1 = valid, NULL = all valid).
 * For DFGPU_UTF8 `values` is the byte buffer and `offsets` the i32 offsets buffer. */
typedef struct dfgpu_col {
  int32_t dtype;
  int32_t _pad;
  int64_t len;
  int64_t offset;
  const void* values;
  const uint8_t* validity;
  const int32_t* offsets;
  int64_t values_bytes; /* UTF8 only: size of the byte buffer; 0 otherwise */
} dfgpu_col;

/* ---- expression programs ----
 * An `Expr` tree (src/logicalplan.rs:136-167) is lowered to a postfix program of dfgpu_insn.
 * This replaces the closure tree built by compile_scalar_expr (src/execution/expression.rs:283-505).
 * `dtype` is the column type for COL, the literal type for LIT, the TARGET type for CAST (`col`
 * then carries the source type) and, advisory, the left operand type for binary ops: the engine
 * re-infers operand types itself and rejects mixed-type operands the way the reference does
 * (ExecutionError "math_ops" / "comparison_ops": expression.rs:166,207).  */
enum {
  DFGPU_OP_COL = 1,  /* Expr::Column(col)                         expression.rs:311-315 */
  DFGPU_OP_LIT = 2,  /* Expr::Literal(ScalarValue)                expression.rs:226-243 */
  DFGPU_OP_CAST = 3, /* Expr::Cast{expr,data_type}                expression.rs:246-280,316-378 */
  DFGPU_OP_ADD = 10, /* Operator::Plus     → array_ops::add       expression.rs:466 */
  DFGPU_OP_SUB = 11, /* Operator::Minus    → array_ops::subtract  expression.rs:473 */
  DFGPU_OP_MUL = 12, /* Operator::Multiply → array_ops::multiply  expression.rs:480 */
  DFGPU_OP_DIV = 13, /* Operator::Divide   → array_ops::divide    expression.rs:487 */
  DFGPU_OP_EQ = 20,  /* array_ops::eq      expression.rs:410 */
  DFGPU_OP_NE = 21,  /* array_ops::neq     expression.rs:417 */
  DFGPU_OP_LT = 22,  /* array_ops::lt      expression.rs:424 */
  DFGPU_OP_LE = 23,  /* array_ops::lt_eq   expression.rs:431 */
  DFGPU_OP_GT = 24,  /* array_ops::gt      expression.rs:438 */
  DFGPU_OP_GE = 25,  /* array_ops::gt_eq   expression.rs:445 */
  DFGPU_OP_AND = 30, /* array_ops::and     expression.rs:452 */
  DFGPU_OP_OR = 31   /* array_ops::or      expression.rs:459 */
};

typedef struct dfgpu_insn {
  int32_t op;
  int32_t col;   /* COL: column index; CAST: source dtype */
  int32_t dtype; /* see above */
  int32_t _pad;
  union {
    double f64;
    int64_t i64;
    uint64_t u64;
    float f32;
  } lit;
} dfgpu_insn;

/* ---- aggregates (src/execution/expression.rs:32-39 AggregateType) ---- */
enum { DFGPU_AGG_MIN = 1, DFGPU_AGG_MAX = 2, DFGPU_AGG_SUM = 3, DFGPU_AGG_COUNT = 4 };

/* One aggregate expression: func(arg).  `arg` is a postfix program (exactly one argument, as
 * compile_expr asserts: expression.rs:91).  `out_dtype` is Expr::AggregateFunction.return_type
 * (arg type for MIN/MAX/SUM, UInt64 for COUNT: src/sqlplanner.rs:320-341). */
typedef struct dfgpu_agg {
  int32_t func;
  int32_t arg_len;
  const dfgpu_insn* arg;
  int32_t out_dtype;
  int32_t _pad;
} dfgpu_agg;

typedef struct dfgpu_ctx dfgpu_ctx;       /* one GPU + stream + memory pool                  */
typedef struct dfgpu_batch dfgpu_batch;   /* device-resident RecordBatch (columns in HBM)    */
typedef struct dfgpu_result dfgpu_result; /* device-resident output batch                    */
typedef struct dfgpu_aggstate dfgpu_aggstate; /* device hash table / accumulators of one AggregateRelation */

/* ---- library / context ---- */
int dfgpu_abi_version(void);
const char* dfgpu_last_error(void);
/* Replaces nothing in the reference (it has no device); called once from ExecutionContext::new
 * (src/execution/context.rs:38).  `device` is the CUDA ordinal this ctx owns. */
int dfgpu_init(int device, dfgpu_ctx** out);
int dfgpu_shutdown(dfgpu_ctx* ctx);
int dfgpu_device_count(int* out);
/* Block until all work queued on the ctx stream is complete. */
int dfgpu_sync(dfgpu_ctx* ctx);
/* Pinned host memory for Arrow buffers (so uploads/downloads are single DMA transfers). */
int dfgpu_host_alloc(size_t bytes, void** out);
int dfgpu_host_free(void* p);
/* Device timing on the ctx stream (CUDA events): start / stop→milliseconds. */
int dfgpu_timer_start(dfgpu_ctx* ctx);
int dfgpu_timer_stop(dfgpu_ctx* ctx, float* ms);
/* Write `bytes` (> L2) of scratch to evict L2 between timed iterations. */
int dfgpu_flush_l2(dfgpu_ctx* ctx);
/* Counters: number of engine kernels launched on this ctx since init. */
int dfgpu_kernel_launches(const dfgpu_ctx* ctx, int64_t* out);
/* Per-kernel device timing of the dominant (scan) kernels: CUDA events recorded on the ctx stream
 * immediately around each launch of the filter/project, hash-aggregate and reduce kernels.
 * enable(1) starts a fresh accumulation; get() synchronises and returns the summed kernel time and
 * the number of timed launches since enable. */
int dfgpu_profile_enable(dfgpu_ctx* ctx, int on);
int dfgpu_profile_get(dfgpu_ctx* ctx, double* kernel_ms, int64_t* launches);

/* ---- batches: the RecordBatch handed to Relation::next's consumer (src/execution/relation.rs:27-32) ---- */
/* Copy the Arrow buffers of one RecordBatch into HBM (one cudaMemcpyAsync per buffer; pageable
 * memory is staged through a pinned ring).  All columns must have the same `len`. */
int dfgpu_batch_upload(dfgpu_ctx* ctx, const dfgpu_col* cols, int ncols, dfgpu_batch** out);
int dfgpu_batch_rows(const dfgpu_batch* b, int64_t* nrows);
int dfgpu_batch_free(dfgpu_batch* b);

/* ---- compile_scalar_expr's checks alone (src/execution/expression.rs:283-505) ----
 * Type-check one expression program against a schema (`col_dtypes[i]` = dtype of column i) exactly as
 * dfgpu_filter_project / dfgpu_aggregate_update would before launching anything, without touching a
 * GPU: identical operand dtypes ("math_ops" / "comparison_ops"), Boolean operands for AND / OR, the CAST
 * rules, column indices.  `out_dtype` receives the result type.  Usable on a machine with no device. */
int dfgpu_check_program(const int32_t* col_dtypes, int ncols, const dfgpu_insn* prog, int prog_len, int32_t* out_dtype);

/* ---- FilterRelation + ProjectRelation fused (src/execution/filter.rs:46-110,
 *      src/execution/projection.rs:46-66, wiring at src/execution/context.rs:126-161) ----
 * pred_len == 0: no WHERE clause.  nproj == 0: emit every input column (what FilterRelation alone
 * does: filter.rs:55-57).  Output rows keep input order (filter.rs:86-90).  A projection whose type is
 * Boolean (a comparison or AND / OR: expression.rs:212-224,236-290) comes back as a DFGPU_BOOL column,
 * bit-packed LSB first like arrow's BooleanArray; dfgpu_result_col_bytes reports (nrows + 7) / 8. */
int dfgpu_filter_project(dfgpu_ctx* ctx, const dfgpu_batch* batch, const dfgpu_insn* pred, int pred_len,
                         const dfgpu_insn* const* proj, const int* proj_len, int nproj, dfgpu_result** out);

/* Same operator, host buffers in, host buffers out, for one big RecordBatch: the batch is cut into
 * row-range chunks and upload (H2D), kernel and download (D2H) of successive chunks overlap on three
 * streams; only the referenced columns cross PCIe.  This is what GpuFilterProjectRelation::next calls
 * for large batches.  The result's columns live in pinned host memory owned by the library
 * (dfgpu_result_col_host_ptr for zero-copy, dfgpu_result_copy_col to copy out).
 * Input buffers should be pinned (dfgpu_host_alloc) for the copies to be asynchronous.
 * Batches whose referenced columns include Utf8, Boolean or nullable columns are not chunk-pipelined: the
 * referenced columns are uploaded whole, the resident operator runs, and the result stays in device memory
 * (dfgpu_result_on_host tells which; dfgpu_result_copy_col works for both). */
int dfgpu_filter_project_host(dfgpu_ctx* ctx, const dfgpu_col* cols, int ncols, const dfgpu_insn* pred, int pred_len,
                              const dfgpu_insn* const* proj, const int* proj_len, int nproj, int64_t chunk_rows /*0 = default*/,
                              dfgpu_result** out);

/* ---- AggregateRelation (src/execution/aggregate.rs:38-61, 615-631, 703-952) ----
 * create → update once per input batch (the `while let Some(batch)` loops at aggregate.rs:707,796)
 * → finish (materialise group columns then aggregate columns: aggregate.rs:890-949; with a
 * communicator attached this is also where the partial-aggregate merge happens).
 * nkeys == 0: no GROUP BY (aggregate.rs:703-785).  Keys are postfix programs (group_expr are
 * compiled scalar exprs: context.rs:171-175); integer and Utf8 types only (aggregate.rs:63-76). */
int dfgpu_aggregate_create(dfgpu_ctx* ctx, const dfgpu_insn* const* keys, const int* key_len, int nkeys,
                           const dfgpu_agg* aggs, int naggs, int64_t expected_groups /*0 = unknown*/,
                           dfgpu_aggstate** out);
/* FilterRelation fused under the aggregate: the wiring `Aggregate{input: Selection{expr, ..}}` that
 * ExecutionContext::execute builds for `SELECT .. WHERE .. GROUP BY ..` (src/execution/context.rs:126-139,
 * 162-192; src/sqlplanner.rs:93-96).  The predicate (a Boolean postfix program over the SAME input columns,
 * else "Filter expression did not evaluate to boolean": filter.rs:62-67) is evaluated inside the scan
 * kernel before the probe: one pass over the input, no intermediate batch.  Must be called before the
 * first dfgpu_aggregate_update; pred_len == 0 removes it. */
int dfgpu_aggregate_set_predicate(dfgpu_aggstate* st, const dfgpu_insn* pred, int pred_len);
int dfgpu_aggregate_update(dfgpu_aggstate* st, const dfgpu_batch* batch);
/* Same as update for one big HOST RecordBatch (the `while let Some(batch)` body with the upload inside): the
 * batch is cut into row-range chunks, every H2D copy is queued up front on a copy stream and the scan of
 * chunk c waits only for chunk c's copies, so PCIe and the kernel overlap.  Input buffers should be pinned
 * (dfgpu_host_alloc).  Nullable / Utf8 / small batches take the plain upload path inside. */
int dfgpu_aggregate_update_host(dfgpu_aggstate* st, const dfgpu_col* cols, int ncols, int64_t chunk_rows /*0 = default*/);
int dfgpu_aggregate_finish(dfgpu_aggstate* st, dfgpu_result** out);
int dfgpu_aggregate_free(dfgpu_aggstate* st);

/* ---- results ---- */
int dfgpu_result_shape(const dfgpu_result* r, int64_t* nrows, int* ncols);
int dfgpu_result_col_dtype(const dfgpu_result* r, int i, int32_t* dtype);
/* Utf8 columns: number of data bytes (offsets buffer has nrows+1 entries). */
int dfgpu_result_col_bytes(const dfgpu_result* r, int i, int64_t* nbytes);
/* null_count of column i (aggregate outputs can be null: aggregate.rs:641-643). */
int dfgpu_result_col_nulls(const dfgpu_result* r, int i, int64_t* null_count);
/* Copy column i to host.  dst_values: nrows*width bytes (Utf8: nbytes); dst_validity: ceil(nrows/8)
 * bytes or NULL; dst_offsets: (nrows+1) i32 for Utf8, else NULL. */
int dfgpu_result_copy_col(const dfgpu_result* r, int i, void* dst_values, uint8_t* dst_validity, int32_t* dst_offsets);
/* 1 when the result's columns live in pinned host memory (the chunk-pipelined dfgpu_filter_project_host), else 0. */
int dfgpu_result_on_host(const dfgpu_result* r, int* on_host);
/* Host pointer of column i's values (host-resident results only). */
int dfgpu_result_col_host_ptr(const dfgpu_result* r, int i, const void** hptr);
/* Device pointer of column i's values (for zero-copy consumers on the same GPU). */
int dfgpu_result_col_device_ptr(const dfgpu_result* r, int i, const void** dptr);
int dfgpu_result_free(dfgpu_result* r);

/* ---- multi-GPU: row-range partitioned batches, one ctx (process) per GPU ----
 * The reference has no distribution (ROADMAP.md:36-51 is roadmap only).  A communicator makes
 * dfgpu_aggregate_finish merge the per-rank partial aggregates (NCCL over NVLink/NVSwitch) so that
 * every rank returns the global result.  `nccl_unique_id` is the 128-byte ncclUniqueId created by
 * dfgpu_comm_unique_id on rank 0 and distributed by the host application. */
int dfgpu_comm_unique_id(uint8_t out_id[128]);
int dfgpu_comm_init(dfgpu_ctx* ctx, int rank, int world, const uint8_t nccl_unique_id[128]);
int dfgpu_comm_destroy(dfgpu_ctx* ctx);
/* Number of ranks of the attached communicator (1 = none). */
int dfgpu_comm_world(const dfgpu_ctx* ctx, int64_t* world);

#ifdef __cplusplus
}
#endif
#endif /* DFGPU_H */
