//! `extern "C"` declarations of include/dfgpu.h (ABI version 2) and the mapping of its status codes onto
//! `ExecutionError` (src/execution/error.rs:51-60).
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_int, c_void};

use super::super::error::{ExecutionError, Result};

#[repr(C)] pub struct dfgpu_ctx { _p: [u8; 0] }
#[repr(C)] pub struct dfgpu_batch { _p: [u8; 0] }
#[repr(C)] pub struct dfgpu_result { _p: [u8; 0] }
#[repr(C)] pub struct dfgpu_aggstate { _p: [u8; 0] }

// dfgpu dtype codes (arrow::datatypes::DataType)
pub const DT_BOOL: i32 = 1;
pub const DT_INT8: i32 = 2;
pub const DT_INT16: i32 = 3;
pub const DT_INT32: i32 = 4;
pub const DT_INT64: i32 = 5;
pub const DT_UINT8: i32 = 6;
pub const DT_UINT16: i32 = 7;
pub const DT_UINT32: i32 = 8;
pub const DT_UINT64: i32 = 9;
pub const DT_FLOAT32: i32 = 10;
pub const DT_FLOAT64: i32 = 11;
pub const DT_UTF8: i32 = 12;

// expression opcodes
pub const OP_COL: i32 = 1;
pub const OP_LIT: i32 = 2;
pub const OP_CAST: i32 = 3;
pub const OP_ADD: i32 = 10;
pub const OP_SUB: i32 = 11;
pub const OP_MUL: i32 = 12;
pub const OP_DIV: i32 = 13;
pub const OP_EQ: i32 = 20;
pub const OP_NE: i32 = 21;
pub const OP_LT: i32 = 22;
pub const OP_LE: i32 = 23;
pub const OP_GT: i32 = 24;
pub const OP_GE: i32 = 25;
pub const OP_AND: i32 = 30;
pub const OP_OR: i32 = 31;

// aggregate functions (src/execution/expression.rs:32-39 AggregateType)
pub const AGG_MIN: i32 = 1;
pub const AGG_MAX: i32 = 2;
pub const AGG_SUM: i32 = 3;
pub const AGG_COUNT: i32 = 4;

/// Borrowed view of one Arrow array (dfgpu_col).
#[repr(C)]
pub struct dfgpu_col {
    pub dtype: i32,
    pub _pad: i32,
    pub len: i64,
    pub offset: i64,
    pub values: *const c_void,
    pub validity: *const u8,
    pub offsets: *const i32,
    pub values_bytes: i64,
}

/// One postfix instruction of an expression program (dfgpu_insn; `lit` carries f64 / i64 / u64 / f32 bits).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct dfgpu_insn {
    pub op: i32,
    pub col: i32,
    pub dtype: i32,
    pub _pad: i32,
    pub lit: u64,
}

#[repr(C)]
pub struct dfgpu_agg {
    pub func: i32,
    pub arg_len: i32,
    pub arg: *const dfgpu_insn,
    pub out_dtype: i32,
    pub _pad: i32,
}

extern "C" {
    pub fn dfgpu_abi_version() -> c_int;
    pub fn dfgpu_last_error() -> *const c_char;
    pub fn dfgpu_init(device: c_int, out: *mut *mut dfgpu_ctx) -> c_int;
    pub fn dfgpu_shutdown(ctx: *mut dfgpu_ctx) -> c_int;
    pub fn dfgpu_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn dfgpu_host_free(p: *mut c_void) -> c_int;
    pub fn dfgpu_batch_upload(ctx: *mut dfgpu_ctx, cols: *const dfgpu_col, ncols: c_int, out: *mut *mut dfgpu_batch) -> c_int;
    pub fn dfgpu_batch_free(b: *mut dfgpu_batch) -> c_int;
    pub fn dfgpu_filter_project(
        ctx: *mut dfgpu_ctx, batch: *const dfgpu_batch, pred: *const dfgpu_insn, pred_len: c_int,
        proj: *const *const dfgpu_insn, proj_len: *const c_int, nproj: c_int, out: *mut *mut dfgpu_result,
    ) -> c_int;
    /// host buffers in, pinned host buffers out, chunk-pipelined H2D | kernel | D2H (large batches)
    pub fn dfgpu_filter_project_host(
        ctx: *mut dfgpu_ctx, cols: *const dfgpu_col, ncols: c_int, pred: *const dfgpu_insn, pred_len: c_int,
        proj: *const *const dfgpu_insn, proj_len: *const c_int, nproj: c_int, chunk_rows: i64, out: *mut *mut dfgpu_result,
    ) -> c_int;
    pub fn dfgpu_aggregate_create(
        ctx: *mut dfgpu_ctx, keys: *const *const dfgpu_insn, key_len: *const c_int, nkeys: c_int,
        aggs: *const dfgpu_agg, naggs: c_int, expected_groups: i64, out: *mut *mut dfgpu_aggstate,
    ) -> c_int;
    /// the WHERE clause of a Selection directly under the Aggregate, fused into the scan kernel
    pub fn dfgpu_aggregate_set_predicate(st: *mut dfgpu_aggstate, pred: *const dfgpu_insn, pred_len: c_int) -> c_int;
    pub fn dfgpu_aggregate_update(st: *mut dfgpu_aggstate, batch: *const dfgpu_batch) -> c_int;
    /// one big host RecordBatch: chunked H2D overlapped with the scan
    pub fn dfgpu_aggregate_update_host(st: *mut dfgpu_aggstate, cols: *const dfgpu_col, ncols: c_int, chunk_rows: i64) -> c_int;
    pub fn dfgpu_aggregate_finish(st: *mut dfgpu_aggstate, out: *mut *mut dfgpu_result) -> c_int;
    pub fn dfgpu_aggregate_free(st: *mut dfgpu_aggstate) -> c_int;
    pub fn dfgpu_result_shape(r: *const dfgpu_result, nrows: *mut i64, ncols: *mut c_int) -> c_int;
    pub fn dfgpu_result_col_dtype(r: *const dfgpu_result, i: c_int, dtype: *mut i32) -> c_int;
    pub fn dfgpu_result_col_bytes(r: *const dfgpu_result, i: c_int, nbytes: *mut i64) -> c_int;
    pub fn dfgpu_result_col_nulls(r: *const dfgpu_result, i: c_int, nulls: *mut i64) -> c_int;
    pub fn dfgpu_result_copy_col(r: *const dfgpu_result, i: c_int, dst_values: *mut c_void, dst_validity: *mut u8, dst_offsets: *mut i32) -> c_int;
    pub fn dfgpu_result_on_host(r: *const dfgpu_result, on_host: *mut c_int) -> c_int;
    pub fn dfgpu_result_col_host_ptr(r: *const dfgpu_result, i: c_int, hptr: *mut *const c_void) -> c_int;
    pub fn dfgpu_result_free(r: *mut dfgpu_result) -> c_int;
    pub fn dfgpu_comm_unique_id(out_id: *mut u8) -> c_int;
    pub fn dfgpu_comm_init(ctx: *mut dfgpu_ctx, rank: c_int, world: c_int, id: *const u8) -> c_int;
    pub fn dfgpu_comm_destroy(ctx: *mut dfgpu_ctx) -> c_int;
    /// number of ranks of the attached communicator (1 without one): a rank whose input is empty still has to join
    /// the merge of dfgpu_aggregate_finish (GpuAggregateRelation::next)
    pub fn dfgpu_comm_world(ctx: *const dfgpu_ctx, world: *mut i64) -> c_int;
    // ---- the rest of include/dfgpu.h (diagnostics, timing, zero-copy consumers), declared for completeness ----
    pub fn dfgpu_device_count(out: *mut c_int) -> c_int;
    pub fn dfgpu_sync(ctx: *mut dfgpu_ctx) -> c_int;
    pub fn dfgpu_timer_start(ctx: *mut dfgpu_ctx) -> c_int;
    pub fn dfgpu_timer_stop(ctx: *mut dfgpu_ctx, ms: *mut f32) -> c_int;
    pub fn dfgpu_flush_l2(ctx: *mut dfgpu_ctx) -> c_int;
    pub fn dfgpu_kernel_launches(ctx: *const dfgpu_ctx, out: *mut i64) -> c_int;
    pub fn dfgpu_profile_enable(ctx: *mut dfgpu_ctx, on: c_int) -> c_int;
    pub fn dfgpu_profile_get(ctx: *mut dfgpu_ctx, kernel_ms: *mut f64, launches: *mut i64) -> c_int;
    pub fn dfgpu_batch_rows(b: *const dfgpu_batch, nrows: *mut i64) -> c_int;
    /// type check of one expression program without a device (the checks compile_scalar_expr makes: expression.rs:136-290)
    pub fn dfgpu_check_program(col_dtypes: *const i32, ncols: c_int, prog: *const dfgpu_insn, prog_len: c_int, out_dtype: *mut i32) -> c_int;
    pub fn dfgpu_result_col_device_ptr(r: *const dfgpu_result, i: c_int, dptr: *mut *const c_void) -> c_int;
}

/// nonzero status -> ExecutionError (src/execution/error.rs:51-60)
pub fn check(rc: c_int) -> Result<()> {
    if rc == 0 {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(dfgpu_last_error()) }.to_string_lossy().into_owned();
    Err(match rc {
        1 => ExecutionError::General(msg),
        3 => ExecutionError::NotImplemented(msg),
        4 => ExecutionError::InvalidColumn(msg),
        5 => ExecutionError::InternalError(msg),
        // 2 EXECUTION, 6 ARROW (DivideByZero, length mismatch), 7 CUDA, 8 OOM
        _ => ExecutionError::ExecutionError(msg),
    })
}
