"""Random expression trees obeying the reference's typing rules (identical operand dtypes for math /
compare, boolean operands for And / Or), for GPU-vs-oracle fuzzing."""
import numpy as np

from datafusion_archive_b200 import _abi as A
from datafusion_archive_b200.expr import BinaryExpr, col, lit

MATH = [A.OP_ADD, A.OP_SUB, A.OP_MUL, A.OP_DIV]
CMP = [A.OP_EQ, A.OP_NE, A.OP_LT, A.OP_LE, A.OP_GT, A.OP_GE]


def gen_numeric(rng, schema, dtype, depth):
    """Expression of type `dtype`."""
    cols = [i for i, d in enumerate(schema) if d == dtype]
    if depth <= 0 or rng.random() < 0.3:
        if cols and rng.random() < 0.75:
            return col(int(rng.choice(cols)))
        if dtype in (A.FLOAT64, A.FLOAT32):
            return lit(float(np.round(rng.random() * 4 - 2, 3)) or 0.5, dtype)
        return lit(int(rng.integers(-5, 6)) or 3, dtype)
    op = int(rng.choice(MATH if dtype in (A.FLOAT64, A.FLOAT32) else MATH[:3]))  # integer division: zero divisors are data dependent
    left = gen_numeric(rng, schema, dtype, depth - 1)
    right = gen_numeric(rng, schema, dtype, depth - 1)
    if op == A.OP_DIV:
        right = lit(float(rng.choice([0.5, 2.0, -3.0, 7.25])), dtype)  # never a zero divisor
    return BinaryExpr(left, op, right)


def gen_bool(rng, schema, depth):
    if depth <= 0 or rng.random() < 0.45:
        dtype = int(rng.choice(sorted(set(schema))))
        return BinaryExpr(gen_numeric(rng, schema, dtype, max(0, depth - 1)), int(rng.choice(CMP)), gen_numeric(rng, schema, dtype, max(0, depth - 1)))
    return BinaryExpr(gen_bool(rng, schema, depth - 1), int(rng.choice([A.OP_AND, A.OP_OR])), gen_bool(rng, schema, depth - 1))


def gen_query(rng, schema, max_depth=3, has_literal_only_ok=False):
    pred = gen_bool(rng, schema, int(rng.integers(0, max_depth + 1))) if rng.random() < 0.85 else None
    nproj = int(rng.integers(1, 4))
    proj = []
    for _ in range(nproj):
        dtype = int(rng.choice(sorted(set(schema))))
        e = gen_numeric(rng, schema, dtype, int(rng.integers(0, max_depth + 1)))
        proj.append(e)
    return pred, proj


def references_column(e):
    from datafusion_archive_b200.expr import Column, Cast
    if isinstance(e, Column):
        return True
    if isinstance(e, BinaryExpr):
        return references_column(e.left) or references_column(e.right)
    if isinstance(e, Cast):
        return references_column(e.expr)
    return False
