"""Python mirror of the reference's expression IR (src/logicalplan.rs:67-167), used by the test and
bench harness to build the postfix expression programs the C ABI takes (include/dfgpu.h).

The production host layer is the C++ mirror under csrc/host/ (the reference is compiled code);
this module only exists so tests read like the reference's own tests, which build `Expr` by hand
(src/execution/aggregate.rs:971-988).
"""
import struct

from . import _abi as A

_CMP = {A.OP_EQ, A.OP_NE, A.OP_LT, A.OP_LE, A.OP_GT, A.OP_GE}
_BOOLOP = {A.OP_AND, A.OP_OR}
_OPNAME = {
    A.OP_ADD: "Plus", A.OP_SUB: "Minus", A.OP_MUL: "Multiply", A.OP_DIV: "Divide", A.OP_EQ: "Eq", A.OP_NE: "NotEq",
    A.OP_LT: "Lt", A.OP_LE: "LtEq", A.OP_GT: "Gt", A.OP_GE: "GtEq", A.OP_AND: "And", A.OP_OR: "Or",
}


class Expr:
    def _bin(self, op, other):
        return BinaryExpr(self, op, _wrap(other))

    def __add__(self, o): return self._bin(A.OP_ADD, o)
    def __sub__(self, o): return self._bin(A.OP_SUB, o)
    def __mul__(self, o): return self._bin(A.OP_MUL, o)
    def __truediv__(self, o): return self._bin(A.OP_DIV, o)
    def __lt__(self, o): return self._bin(A.OP_LT, o)
    def __le__(self, o): return self._bin(A.OP_LE, o)
    def __gt__(self, o): return self._bin(A.OP_GT, o)
    def __ge__(self, o): return self._bin(A.OP_GE, o)
    def eq(self, o): return self._bin(A.OP_EQ, o)
    def not_eq(self, o): return self._bin(A.OP_NE, o)
    def __and__(self, o): return self._bin(A.OP_AND, o)
    def __or__(self, o): return self._bin(A.OP_OR, o)

    def cast(self, dtype):
        return Cast(self, dtype)

    def program(self, schema_dtypes):
        out = []
        self._emit(schema_dtypes, out)
        return out


class Column(Expr):
    def __init__(self, index):
        self.index = index

    def get_type(self, schema):
        return schema[self.index] if 0 <= self.index < len(schema) else 0

    def _emit(self, schema, out):
        i = A.Insn()
        # an out-of-range index is passed through: the engine reports InvalidColumn
        i.op, i.col, i.dtype = A.OP_COL, self.index, (schema[self.index] if 0 <= self.index < len(schema) else 0)
        out.append(i)

    def __repr__(self):
        return "#%d" % self.index


class Literal(Expr):
    """Expr::Literal(ScalarValue); default typing follows the planner: Python int -> Int64,
    float -> Float64 (src/sqlplanner.rs:214-218)."""

    def __init__(self, value, dtype=None):
        if dtype is None:
            dtype = A.FLOAT64 if isinstance(value, float) else A.INT64
        self.value, self.dtype = value, dtype

    def get_type(self, schema):
        return self.dtype

    def _emit(self, schema, out):
        i = A.Insn()
        i.op, i.dtype = A.OP_LIT, self.dtype
        if self.dtype == A.FLOAT64:
            i.lit.f64 = float(self.value)
        elif self.dtype == A.FLOAT32:
            i.lit.u64 = 0
            i.lit.f32 = float(self.value)
        elif self.dtype in (A.UINT8, A.UINT16, A.UINT32, A.UINT64):
            i.lit.u64 = int(self.value)
        else:
            i.lit.i64 = int(self.value)
        out.append(i)

    def __repr__(self):
        return "%s(%r)" % (A.DTYPE_NAMES[self.dtype], self.value)


class Cast(Expr):
    def __init__(self, expr, dtype):
        self.expr, self.dtype = expr, dtype

    def get_type(self, schema):
        return self.dtype

    def _emit(self, schema, out):
        self.expr._emit(schema, out)
        i = A.Insn()
        i.op, i.dtype, i.col = A.OP_CAST, self.dtype, self.expr.get_type(schema)
        out.append(i)

    def __repr__(self):
        return "CAST(%r AS %s)" % (self.expr, A.DTYPE_NAMES[self.dtype])


class BinaryExpr(Expr):
    def __init__(self, left, op, right):
        self.left, self.op, self.right = left, op, right

    def get_type(self, schema):
        if self.op in _CMP or self.op in _BOOLOP:
            return A.BOOL
        return self.left.get_type(schema)  # op_type = left type (expression.rs:408)

    def _emit(self, schema, out):
        self.left._emit(schema, out)
        self.right._emit(schema, out)
        i = A.Insn()
        i.op = self.op
        i.dtype = self.left.get_type(schema)  # operand type (advisory: the engine re-infers and checks)
        out.append(i)

    def __repr__(self):
        return "%r %s %r" % (self.left, _OPNAME[self.op], self.right)


class AggregateFunction:
    """Expr::AggregateFunction{name,args,return_type} (src/logicalplan.rs:162-166)."""

    _F = {"min": A.AGG_MIN, "max": A.AGG_MAX, "sum": A.AGG_SUM, "count": A.AGG_COUNT}

    def __init__(self, name, arg, return_type=None):
        self.name, self.arg, self.return_type = name, _wrap(arg), return_type

    def lower(self, schema):
        func = self._F[self.name.lower()]
        rt = self.return_type
        if rt is None:
            rt = A.UINT64 if func == A.AGG_COUNT else self.arg.get_type(schema)
        return (func, self.arg.program(schema), rt)


def _wrap(x):
    return x if isinstance(x, Expr) else Literal(x)


def col(i):
    return Column(i)


def lit(v, dtype=None):
    return Literal(v, dtype)


def f64_bits(x):
    return struct.unpack("<Q", struct.pack("<d", x))[0]
