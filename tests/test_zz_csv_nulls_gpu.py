"""CSV with empty fields -> nulls -> the GPU operators, through ExecutionContext.sql().  The file is
test/data/null_test.csv of the reference (read by none of its tests): the expectations come from the
oracle on the same nullable arrays.  Runs last: everything here is an extension of the pinned surface."""
import os

import numpy as np
import pytest

import oracle_lib as O
from datafusion_archive_b200 import _abi as A
from datafusion_archive_b200 import host
from datafusion_archive_b200.expr import AggregateFunction, col, lit

pytestmark = pytest.mark.gpu
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")
FIELDS = [("c_int", A.INT32), ("c_float", A.FLOAT64), ("c_string", A.UTF8), ("c_bool", A.BOOL)]


def nullable(values, valid):
    import pyarrow as pa
    values = np.ascontiguousarray(values)
    bits = np.packbits(np.asarray(valid, dtype=bool), bitorder="little")
    return pa.Array.from_buffers(pa.from_numpy_dtype(values.dtype), len(values), [pa.py_buffer(bits.tobytes()), pa.py_buffer(values.tobytes())])


@pytest.fixture()
def ctx():
    c = host.ExecutionContext(0)
    yield c
    c.close()


def columns(rel):
    batches = rel.collect()
    assert len(batches) == 1
    return batches[0]


def unpack(c):
    return c if isinstance(c, tuple) else (np.asarray(c), np.ones(len(c), dtype=bool))


def test_filter_over_a_column_with_nulls(ctx):
    c_int = np.array([1, 2, 3, 4, 5], dtype=np.int32)
    c_float = nullable(np.array([1.1, 2.2, 0.0, 4.4, 6.6]), [1, 1, 0, 1, 1])
    O.set_extensions(filter_all_primitives=True)
    try:
        exp = O.filter_project([c_int, c_float], col(1) > lit(2.0), [col(0), col(1)])
    finally:
        O.set_extensions(filter_all_primitives=False)
    ctx.register_csv("t", os.path.join(DATA, "null_test.csv"), FIELDS, 1024)
    got = columns(ctx.sql("SELECT c_int, c_float FROM t WHERE c_float > 2.0"))
    for g, e in zip(got, exp):
        (gv, gm), (ev, em) = unpack(g), unpack(e)
        assert np.array_equal(gm, em) and np.array_equal(gv[gm], ev[em])
    assert list(unpack(got[0])[0]) == [2, 4, 5]  # gt(null, x) is false: the null row is dropped


def test_aggregates_skip_nulls(ctx):
    c_float = nullable(np.array([1.1, 2.2, 0.0, 4.4, 6.6]), [1, 1, 0, 1, 1])
    exp = O.aggregate([c_float], [], [AggregateFunction("min", col(0)), AggregateFunction("max", col(0)), AggregateFunction("sum", col(0))])
    ctx.register_csv("t", os.path.join(DATA, "null_test.csv"), FIELDS, 1024)
    got = columns(ctx.sql("SELECT MIN(c_float), MAX(c_float), SUM(c_float) FROM t"))
    assert unpack(got[0])[0][0] == unpack(exp[0])[0][0] == 1.1
    assert unpack(got[1])[0][0] == unpack(exp[1])[0][0] == 6.6
    assert abs(unpack(got[2])[0][0] - unpack(exp[2])[0][0]) <= 1e-9 * abs(unpack(exp[2])[0][0])


def test_project_all_columns(ctx):
    # src/execution/projection.rs:83-103: ProjectRelation over people.csv with [Column(0)] -> one column named "id"
    ctx.register_csv("people", os.path.join(DATA, "people.csv"), [("id", A.INT32), ("first_name", A.UTF8)], 1024)
    rel = ctx.sql("SELECT id FROM people")
    assert rel.schema() == [("id", A.INT32)]
    (batch,) = rel.collect()
    assert len(batch) == 1 and batch[0].dtype == np.int32 and batch[0][0] == 1
