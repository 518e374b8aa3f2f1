"""The operators no reference test pins ("parity unpinned") cross-checked against an
independent implementation (numpy) on the same buffers: Minus/Multiply/Divide, Eq/NotEq/LtEq/GtEq, Or,
every non-Float64 dtype, integer wrap-around, NaN / signed zero in MIN/MAX, COUNT, and the arrow 0.12
null rules as SURVEY §8c states them.  CPU only."""
import numpy as np
import pytest

import oracle_lib as O
from datafusion_archive_b200 import _abi as A
from datafusion_archive_b200.expr import AggregateFunction, col, lit

INT_DTYPES = [np.int8, np.int16, np.int32, np.int64, np.uint8, np.uint16, np.uint32, np.uint64]
ALL_DTYPES = INT_DTYPES + [np.float32, np.float64]
DT = {np.int8: A.INT8, np.int16: A.INT16, np.int32: A.INT32, np.int64: A.INT64, np.uint8: A.UINT8, np.uint16: A.UINT16,
      np.uint32: A.UINT32, np.uint64: A.UINT64, np.float32: A.FLOAT32, np.float64: A.FLOAT64}


def rand(dt, n, rng, nonzero=False):
    if np.issubdtype(dt, np.floating):
        x = (rng.random(n) * 200 - 100).astype(dt)
    else:
        info = np.iinfo(dt)
        x = rng.integers(info.min, info.max, n, dtype=dt, endpoint=True)
    if nonzero:
        x[x == 0] = 1
    return x


def nullable(values, valid):
    import pyarrow as pa
    values = np.ascontiguousarray(values)
    bits = np.packbits(np.asarray(valid, dtype=bool), bitorder="little")
    return pa.Array.from_buffers(pa.from_numpy_dtype(values.dtype), len(values), [pa.py_buffer(bits.tobytes()), pa.py_buffer(values.tobytes())])


def unpack(c):
    return c if isinstance(c, tuple) else (c, np.ones(len(c), dtype=bool))


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_arithmetic_matches_numpy_wrapping(dt):
    rng = np.random.default_rng(11)
    a, b = rand(dt, 5000, rng), rand(dt, 5000, rng, nonzero=True)
    if np.issubdtype(dt, np.signedinteger):
        b[(a == np.iinfo(dt).min) & (b == -1)] = 1  # MIN / -1 overflows (Rust panics in every build mode)
    got = O.filter_project([a, b], None, [col(0) + col(1), col(0) - col(1), col(0) * col(1), col(0) / col(1)])
    with np.errstate(over="ignore"):
        exp = [a + b, a - b, a * b]
        if np.issubdtype(dt, np.floating):
            exp.append(a / b)
        else:  # Rust integer division truncates toward zero
            q = np.abs(a.astype(object)) // np.abs(b.astype(object))
            sign = np.where((a.astype(object) < 0) != (b.astype(object) < 0), -1, 1)
            exp.append(np.array([int(x) for x in q * sign], dtype=object).astype(dt))
    for g, e in zip(got, exp):
        assert g.dtype == np.dtype(dt)
        assert np.array_equal(g.view(np.uint8), np.asarray(e, dtype=dt).view(np.uint8))


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_comparisons_match_numpy(dt):
    rng = np.random.default_rng(12)
    a = rand(dt, 4000, rng)
    b = a.copy()
    swap = rng.random(4000) < 0.6
    b[swap] = rand(dt, int(swap.sum()), rng)
    if np.issubdtype(dt, np.floating):
        a[::97] = np.nan
        b[::89] = np.nan
        a[5], b[5] = 0.0, -0.0
    got = O.filter_project([a, b], None, [col(0).eq(col(1)), col(0).not_eq(col(1)), col(0) < col(1), col(0) <= col(1),
                                            col(0) > col(1), col(0) >= col(1)])
    exp = [a == b, a != b, a < b, a <= b, a > b, a >= b]
    for g, e in zip(got, exp):
        assert g.dtype == bool and np.array_equal(g, e)


def test_boolean_connectives_and_predicate_use():
    rng = np.random.default_rng(13)
    a, b, c = rng.random(3000), rng.random(3000), rng.random(3000)
    p = ((col(0) < col(1)) | (col(2) > lit(0.7))) & (col(0) >= lit(0.1))
    m = ((a < b) | (c > 0.7)) & (a >= 0.1)
    (got,) = O.filter_project([a, b, c], None, [p])
    assert np.array_equal(got, m)
    got = O.filter_project([a, b, c], p, [col(0), col(2)])
    assert np.array_equal(got[0], a[m]) and np.array_equal(got[1], c[m])


def test_divide_by_zero_is_an_error_for_ints_and_floats():
    for dt in (np.int32, np.float64):
        a, b = np.array([1, 2, 3], dtype=dt), np.array([1, 0, 1], dtype=dt)
        with pytest.raises(O.OracleError) as e:
            O.filter_project([a, b], None, [col(0) / col(1)])
        assert "DivideByZero" in e.value.msg


@pytest.mark.parametrize("dt", ALL_DTYPES)
def test_aggregates_match_numpy(dt):
    rng = np.random.default_rng(14)
    n = 20000
    k = rng.integers(0, 37, n, dtype=np.int32)
    v = rand(dt, n, rng) if np.issubdtype(dt, np.floating) else rng.integers(0, 100, n).astype(dt)
    aggs = [AggregateFunction(f, col(1)) for f in ("min", "max", "sum", "count")]
    got = O.aggregate([k, v], [col(0)], aggs)
    o = np.argsort(got[0])
    keys = got[0][o]
    assert np.array_equal(keys, np.unique(k))
    for j, key in enumerate(keys):
        sel = v[k == key]
        assert got[1][o][j] == sel.min() and got[2][o][j] == sel.max() and got[4][o][j] == len(sel)
        if np.issubdtype(dt, np.floating):
            acc = dt(0)
            for x in sel:  # strictly in row order (aggregate.rs:277-278)
                acc = dt(acc + x)
            assert got[3][o][j] == acc
        else:
            with np.errstate(over="ignore"):
                assert got[3][o][j] == sel.sum(dtype=dt)  # wrapping
    # no GROUP BY: array_ops::{min,max,sum}
    got = O.aggregate([v], [], [AggregateFunction(f, col(0)) for f in ("min", "max", "sum", "count")])
    assert got[0][0] == v.min() and got[1][0] == v.max() and got[3][0] == n


def test_nan_handling_differs_between_group_by_and_column_reduce():
    v = np.array([2.0, np.nan, 1.0, 3.0])
    k = np.zeros(4, dtype=np.int32)
    # GROUP BY folds with f64::min / f64::max (aggregate.rs:139-140,208-209): NaN is ignored
    got = O.aggregate([k, v], [col(0)], [AggregateFunction("min", col(1)), AggregateFunction("max", col(1))])
    assert got[1][0] == 1.0 and got[2][0] == 3.0
    # without GROUP BY array_ops::{min,max} replace on `m < n` / `m > n`: a NaN never replaces, and nothing replaces a leading NaN
    got = O.aggregate([v], [], [AggregateFunction("min", col(0)), AggregateFunction("max", col(0))])
    assert got[0][0] == 1.0 and got[1][0] == 3.0
    got = O.aggregate([np.array([np.nan, 1.0, 2.0])], [], [AggregateFunction("min", col(0))])
    assert np.isnan(got[0][0])


def test_null_rules_of_array_ops():
    a = nullable(np.array([1.0, 2.0, 3.0, 4.0]), [1, 0, 1, 0])
    b = nullable(np.array([1.0, 5.0, 0.5, 9.0]), [1, 1, 0, 0])
    add, lt, gt, eq, ne = O.filter_project([a, b], None, [col(0) + col(1), col(0) < col(1), col(0) > col(1), col(0).eq(col(1)), col(0).not_eq(col(1))])
    v, m = unpack(add)
    assert list(m) == [True, False, False, False] and v[0] == 2.0          # null if either side is null
    for c in (lt, gt, eq, ne):
        assert not isinstance(c, tuple) or c[1].all()                          # comparisons never yield nulls
    assert list(unpack(lt)[0]) == [False, True, False, True]                  # lt: null on the left -> true
    assert list(unpack(gt)[0]) == [False, False, True, True]                  # gt: null on the right -> true
    assert list(unpack(eq)[0]) == [True, False, False, True]                  # eq compares the Options
    assert list(unpack(ne)[0]) == [False, True, True, False]
    # filter ignores the validity of the columns it gathers (filter.rs:86-90): no nulls come out
    out = O.filter_project([a, b], col(0) < col(1), [col(0), col(1)])
    assert all(not isinstance(c, tuple) for c in out) and list(out[0]) == [2.0, 4.0]
    # min / max / sum skip nulls; an all-null column aggregates to null
    s = O.aggregate([a], [], [AggregateFunction("sum", col(0)), AggregateFunction("min", col(0)), AggregateFunction("count", col(0))])
    assert unpack(s[0])[0][0] == 4.0 and unpack(s[1])[0][0] == 1.0 and unpack(s[2])[0][0] == 2
    none = nullable(np.array([1.0, 2.0]), [0, 0])
    s = O.aggregate([none], [], [AggregateFunction("sum", col(0)), AggregateFunction("max", col(0))])
    assert not unpack(s[0])[1][0] and not unpack(s[1])[1][0]
