"""The product's expression type checker (dfgpu_check_program: the checks the operators run before any
launch) against the oracle's, on the CPU: random well- and ill-typed expression trees must be accepted /
rejected alike, with the same error class and the same result type.  No GPU."""
import numpy as np
import pytest

import oracle_lib as O
from datafusion_archive_b200 import _abi as A
from datafusion_archive_b200 import engine
from datafusion_archive_b200.expr import BinaryExpr, col, lit

NP = {A.INT8: np.int8, A.INT16: np.int16, A.INT32: np.int32, A.INT64: np.int64, A.UINT8: np.uint8, A.UINT16: np.uint16,
      A.UINT32: np.uint32, A.UINT64: np.uint64, A.FLOAT32: np.float32, A.FLOAT64: np.float64}
OPS = [A.OP_ADD, A.OP_SUB, A.OP_MUL, A.OP_DIV, A.OP_EQ, A.OP_NE, A.OP_LT, A.OP_LE, A.OP_GT, A.OP_GE, A.OP_AND, A.OP_OR]


def gen(rng, schema, depth):
    if depth <= 0 or rng.random() < 0.3:
        if rng.random() < 0.7:
            return col(int(rng.integers(0, len(schema))))
        dt = int(rng.choice(list(NP)))
        return lit(float(rng.integers(1, 5)) if dt in (A.FLOAT32, A.FLOAT64) else int(rng.integers(1, 5)), dt)
    return BinaryExpr(gen(rng, schema, depth - 1), int(rng.choice(OPS)), gen(rng, schema, depth - 1))


def outcome(fn):
    try:
        return ("ok", fn())
    except engine.DfGpuError as e:
        return ("err", e.code)
    except O.OracleError as e:
        return ("err", e.code)


def test_type_checker_agrees_with_the_oracle_on_random_trees():
    rng = np.random.default_rng(99)
    schema = [A.FLOAT64, A.FLOAT64, A.INT64, A.INT32, A.UINT8, A.FLOAT32]
    arrays = [np.array([3], dtype=NP[d]) for d in schema]  # one row, no zero divisors
    agree_ok = agree_err = 0
    for _ in range(1500):
        e = gen(rng, schema, int(rng.integers(0, 4)))
        got = outcome(lambda: engine.check_program(schema, e))

        def run_oracle():
            (c,) = O.filter_project(arrays, None, [e])
            c = c[0] if isinstance(c, tuple) else c
            return A.BOOL if c.dtype == bool else [k for k, v in NP.items() if np.dtype(v) == c.dtype][0]
        exp = outcome(run_oracle)
        # the reference checks operand types inside the closures, at evaluation time, so a data error in
        # a child (DivideByZero) surfaces before the parent's type error; the product checks types first
        data_error_first = got[0] == exp[0] == "err" and exp[1] == A.ERR_ARROW
        assert got == exp or data_error_first, "%r: product %r, oracle %r" % (e, got, exp)
        agree_ok += got[0] == "ok"
        agree_err += got[0] == "err"
    assert agree_ok > 100 and agree_err > 100


def test_reference_error_classes_without_a_gpu():
    schema = [A.INT64, A.FLOAT64, A.UTF8]
    for e, code, msg in [(col(0) + col(1), A.ERR_EXECUTION, "math_ops"), (col(1) > lit(1), A.ERR_EXECUTION, "comparison_ops"),
                         ((col(1) + col(1)) & (col(1) > lit(1.0)), A.ERR_INTERNAL, "boolean_ops"), (col(7), A.ERR_INVALID_COLUMN, "out of range"),
                         ((col(1) + col(1)).cast(A.INT32), A.ERR_GENERAL, "CAST not implemented for expression"),
                         (lit(1.5).cast(A.INT32), A.ERR_NOT_IMPLEMENTED, "CAST from Float64")]:
        with pytest.raises(engine.DfGpuError) as err:
            engine.check_program(schema, e)
        assert err.value.code == code and msg in err.value.msg, (e, err.value.code, err.value.msg)
    assert engine.check_program(schema, col(1) * col(1) < col(1)) == A.BOOL
    assert engine.check_program(schema, col(0).cast(A.INT16)) == A.INT16
    assert engine.check_program(schema, col(2)) == A.UTF8


def test_malformed_programs_are_rejected_not_crashed():
    import ctypes as C
    L = engine.lib()
    L.dfgpu_check_program.argtypes = [C.POINTER(C.c_int32), C.c_int, C.POINTER(A.Insn), C.c_int, C.POINTER(C.c_int32)]
    rng = np.random.default_rng(5)
    dts = (C.c_int32 * 3)(A.FLOAT64, A.INT64, A.UTF8)
    out = C.c_int32()
    rejected = 0
    for _ in range(3000):
        n = int(rng.integers(1, 12))
        prog = (A.Insn * n)()
        for i in range(n):
            prog[i].op = int(rng.choice([A.OP_COL, A.OP_LIT, A.OP_CAST, A.OP_ADD, A.OP_DIV, A.OP_EQ, A.OP_AND, 0, 99, -1]))
            prog[i].col = int(rng.integers(-2, 5))
            prog[i].dtype = int(rng.integers(-1, 15))
            prog[i].lit.u64 = int(rng.integers(0, 2**63))
        rc = L.dfgpu_check_program(dts, 3, prog, n, C.byref(out))
        rejected += rc != 0
        assert 0 <= rc <= 8
    assert rejected > 1000
    assert L.dfgpu_check_program(dts, 3, None, 1, C.byref(out)) == A.ERR_GENERAL
