//! `Expr` -> postfix expression program of the C ABI.  Replaces compile_scalar_expr
//! (src/execution/expression.rs:283-505): instead of a closure tree the expression is flattened; type
//! checking (identical operand dtypes, Boolean operands for And / Or, the Cast rules) happens inside the
//! library with the reference's error strings.
use arrow::datatypes::{DataType, Schema};

use super::super::error::{ExecutionError, Result};
use super::super::super::logicalplan::{Expr, Operator, ScalarValue};
use super::ffi::*;

pub fn dtype_code(dt: &DataType) -> Result<i32> {
    Ok(match dt {
        DataType::Boolean => DT_BOOL,
        DataType::Int8 => DT_INT8,
        DataType::Int16 => DT_INT16,
        DataType::Int32 => DT_INT32,
        DataType::Int64 => DT_INT64,
        DataType::UInt8 => DT_UINT8,
        DataType::UInt16 => DT_UINT16,
        DataType::UInt32 => DT_UINT32,
        DataType::UInt64 => DT_UINT64,
        DataType::Float32 => DT_FLOAT32,
        DataType::Float64 => DT_FLOAT64,
        DataType::Utf8 => DT_UTF8,
        other => return Err(ExecutionError::NotImplemented(format!("data type {:?} on the GPU path", other))),
    })
}

fn insn(op: i32, col: i32, dtype: i32, lit: u64) -> dfgpu_insn {
    dfgpu_insn { op, col, dtype, _pad: 0, lit }
}

/// (dtype code, raw bits) of a literal — Expr::Literal(ScalarValue), expression.rs:289-310
fn literal(v: &ScalarValue) -> Result<(i32, u64)> {
    Ok(match v {
        ScalarValue::Int8(x) => (DT_INT8, *x as i64 as u64),
        ScalarValue::Int16(x) => (DT_INT16, *x as i64 as u64),
        ScalarValue::Int32(x) => (DT_INT32, *x as i64 as u64),
        ScalarValue::Int64(x) => (DT_INT64, *x as u64),
        ScalarValue::UInt8(x) => (DT_UINT8, *x as u64),
        ScalarValue::UInt16(x) => (DT_UINT16, *x as u64),
        ScalarValue::UInt32(x) => (DT_UINT32, *x as u64),
        ScalarValue::UInt64(x) => (DT_UINT64, *x),
        ScalarValue::Float32(x) => (DT_FLOAT32, x.to_bits() as u64),
        ScalarValue::Float64(x) => (DT_FLOAT64, x.to_bits()),
        other => return Err(ExecutionError::ExecutionError(format!("No support for literal type {:?}", other))),
    })
}

fn op_code(op: &Operator) -> Result<i32> {
    Ok(match op {
        Operator::Eq => OP_EQ,
        Operator::NotEq => OP_NE,
        Operator::Lt => OP_LT,
        Operator::LtEq => OP_LE,
        Operator::Gt => OP_GT,
        Operator::GtEq => OP_GE,
        Operator::And => OP_AND,
        Operator::Or => OP_OR,
        Operator::Plus => OP_ADD,
        Operator::Minus => OP_SUB,
        Operator::Multiply => OP_MUL,
        Operator::Divide => OP_DIV,
        other => return Err(ExecutionError::ExecutionError(format!("operator: {:?}", other))), // expression.rs:494-497
    })
}

/// `remap[i]` = index of input column i among the columns actually uploaded (pruned to the referenced ones).
pub fn lower(e: &Expr, schema: &Schema, remap: &[Option<usize>], out: &mut Vec<dfgpu_insn>) -> Result<()> {
    match e {
        Expr::Column(i) => {
            let at = remap.get(*i).and_then(|x| *x).ok_or_else(|| ExecutionError::InvalidColumn(format!("column index {} out of range", i)))?;
            out.push(insn(OP_COL, at as i32, dtype_code(schema.field(*i).data_type())?, 0));
        }
        Expr::Literal(v) => {
            let (dt, bits) = literal(v)?;
            out.push(insn(OP_LIT, 0, dt, bits));
        }
        Expr::Cast { expr, data_type } => {
            lower(expr, schema, remap, out)?;
            let src = expr.get_type(schema);
            out.push(insn(OP_CAST, dtype_code(&src)?, dtype_code(data_type)?, 0)); // col carries the source dtype
        }
        Expr::BinaryExpr { left, op, right } => {
            lower(left, schema, remap, out)?;
            lower(right, schema, remap, out)?;
            out.push(insn(op_code(op)?, 0, dtype_code(&left.get_type(schema)).unwrap_or(0), 0));
        }
        other => return Err(ExecutionError::ExecutionError(format!("expression {:?}", other))), // expression.rs:500-503
    }
    Ok(())
}

/// Columns an expression reads (collect_expr, src/sqlplanner.rs:435-458).
pub fn collect_columns(e: &Expr, acc: &mut Vec<usize>) {
    match e {
        Expr::Column(i) => {
            if !acc.contains(i) {
                acc.push(*i)
            }
        }
        Expr::BinaryExpr { left, right, .. } => {
            collect_columns(left, acc);
            collect_columns(right, acc);
        }
        Expr::Cast { expr, .. } | Expr::IsNull(expr) | Expr::IsNotNull(expr) => collect_columns(expr, acc),
        Expr::Sort { expr, .. } => collect_columns(expr, acc),
        Expr::AggregateFunction { args, .. } | Expr::ScalarFunction { args, .. } => args.iter().for_each(|a| collect_columns(a, acc)),
        Expr::Literal(_) => {}
    }
}
