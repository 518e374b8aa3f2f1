"""End-to-end through the reference-shaped API: ExecutionContext.sql() -> Relation.next(), from the
reference's own CSV fixtures (tests/sql.rs) and from in-memory Arrow batches.  GPU required."""
import os

import numpy as np
import pytest

import oracle_lib as O
from datafusion_archive_b200 import _abi as A
from datafusion_archive_b200 import host, workloads
from datafusion_archive_b200.expr import AggregateFunction, col, lit

pytestmark = pytest.mark.gpu
DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data")


@pytest.fixture()
def ctx():
    c = host.ExecutionContext(0)
    yield c
    c.close()


def register_cities(ctx):  # tests/sql.rs:79-87
    ctx.register_csv("cities", os.path.join(DATA, "uk_cities.csv"), [("city", A.UTF8), ("lat", A.FLOAT64), ("lng", A.FLOAT64)], 1024)


def result_rows(rel):
    rows = []
    for batch in rel.collect():
        rows.extend(zip(*batch))
    return rows


def test_csv_query_with_predicate_numeric_columns(ctx, golden, fmt_f64):
    # tests/sql.rs:30-37 without the Utf8 column (GPU Utf8 gather: SURVEY §8f-3)
    register_cities(ctx)
    rel = ctx.sql("SELECT lat, lng, lat + lng FROM cities WHERE lat > 51.0 AND lat < 53")
    assert rel.schema() == [("lat", A.FLOAT64), ("lng", A.FLOAT64), ("#1 Plus #2", A.FLOAT64)]  # projection.rs:52-57 names
    got = ["%s\t%s\t%s" % tuple(fmt_f64(x) for x in r) for r in result_rows(rel)]
    assert got == [l.split("\t", 1)[1] for l in golden["csv_query_with_predicate"]["expected"].splitlines()]
    assert rel.next() is None


def test_csv_query_with_predicate_verbatim(ctx, golden, fmt_f64):
    # tests/sql.rs:30-37 exactly as the reference runs it: SQL text in, golden string out
    register_cities(ctx)
    rel = ctx.sql(golden["csv_query_with_predicate"]["sql"])
    s = "".join('"%s"\t%s\t%s\t%s\n' % (a, fmt_f64(b), fmt_f64(c), fmt_f64(d)) for a, b, c, d in result_rows(rel))
    assert s == golden["csv_query_with_predicate"]["expected"]


def test_csv_query_cast(ctx, golden):
    register_cities(ctx)
    rows = result_rows(ctx.sql(golden["csv_query_cast"]["sql"]))
    assert "".join("%d\n" % r[0] for r in rows) == golden["csv_query_cast"]["expected"]


def test_csv_query_group_by_int_min_max(ctx, golden, fmt_f64):
    ctx.register_csv("t1", os.path.join(DATA, "aggregate_test_1.csv"), [("a", A.INT32), ("b", A.FLOAT64)], 1024)
    rows = result_rows(ctx.sql(golden["csv_query_group_by_int_min_max"]["sql"]))
    got = sorted("%d\t%s\t%s\n" % (a, fmt_f64(b), fmt_f64(c)) for a, b, c in rows)
    assert got == sorted(golden["csv_query_group_by_int_min_max"]["expected"].splitlines(True))


def test_csv_query_group_by_string_min_max(ctx, golden, fmt_f64):
    # tests/sql.rs:55-67 verbatim
    ctx.register_csv("t1", os.path.join(DATA, "aggregate_test_2.csv"), [("a", A.UTF8), ("b", A.FLOAT64)], 1024)
    rows = result_rows(ctx.sql(golden["csv_query_group_by_string_min_max"]["sql"]))
    got = sorted('"%s"\t%s\t%s\n' % (a, fmt_f64(b), fmt_f64(c)) for a, b, c in rows)
    assert got == sorted(golden["csv_query_group_by_string_min_max"]["expected"].splitlines(True))


def test_min_max_lat(ctx, golden):
    register_cities(ctx)
    rel = ctx.sql("SELECT MIN(lat), MAX(lat) FROM cities")
    rows = result_rows(rel)
    assert rows == [(golden["min_lat"], golden["max_lat"])]
    assert rel.next() is None  # AggregateRelation is one-shot (aggregate.rs:616-619)


def test_small_batches_match_single_batch(ctx):
    # a CsvDataSource is a one-pass reader shared by reference (Rc<RefCell<DataSource>>, context.rs:100-102):
    # every query gets its own registration, as in the reference's tests
    fields = [("city", A.UTF8), ("lat", A.FLOAT64), ("lng", A.FLOAT64)]
    for name, bs in [("c7", 7), ("c1024", 1024), ("d7", 7), ("d1024", 1024)]:
        ctx.register_csv(name, os.path.join(DATA, "uk_cities.csv"), fields, bs)
    q = "SELECT lat * lng, lng FROM %s WHERE lng < 0 OR lat > 55.5"
    a, b = result_rows(ctx.sql(q % "c7")), result_rows(ctx.sql(q % "c1024"))
    assert a == b and len(a) > 0
    q = "SELECT SUM(lat), COUNT(lat), MIN(lng) FROM %s"
    a, b = result_rows(ctx.sql(q % "d7")), result_rows(ctx.sql(q % "d1024"))
    assert a[0][1] == b[0][1] == 36 and a[0][2] == b[0][2]
    assert abs(a[0][0] - b[0][0]) <= 1e-9 * abs(b[0][0])
    # the exhausted reader yields nothing on a second query: no GROUP BY -> one row of nulls (aggregate.rs:641-643)
    vals, mask = ctx.sql("SELECT SUM(lat) FROM d7").collect()[0][0]
    assert len(vals) == 1 and not mask[0]


def test_memory_tables_baseline_queries(ctx):
    n = 300_000
    arrays, pred, proj = workloads.c3(n)
    ctx.register_memory("t", list(zip("abcd", arrays)))
    got = ctx.sql("SELECT a+b, a*b FROM t WHERE b<a").collect()
    assert len(got) == 1
    exp = O.filter_project(arrays, pred, proj)
    for g, e in zip(got[0], exp):
        assert np.array_equal(g.view(np.uint8), e.view(np.uint8))
    ctx.register_memory("t", list(zip("abcd", arrays)))  # data sources are one-pass (see above)
    got = ctx.sql("SELECT a FROM t WHERE a > 0.5").collect()[0][0]
    assert np.array_equal(got, arrays[0][arrays[0] > 0.5])
    # batched source: output is the concatenation of per-batch outputs
    ctx.register_memory("tb", list(zip("abcd", arrays)), batch_size=65536)
    parts = ctx.sql("SELECT a FROM tb WHERE a > 0.5").collect()
    assert len(parts) == 5 and np.array_equal(np.concatenate([p[0] for p in parts]), got)

    arrays4, keys, aggs, _ = workloads.c4(n, nkeys=5000)
    ctx.register_memory("g", [("k", arrays4[0]), ("v", arrays4[1])], batch_size=100_000)
    rel = ctx.sql("SELECT k, SUM(v), COUNT(v) FROM g GROUP BY k")
    assert [f[0] for f in rel.schema()] == ["k", "SUM", "COUNT"]
    k, s, c = rel.collect()[0]
    exp = O.aggregate(arrays4, keys, aggs)
    o1, o2 = np.argsort(k), np.argsort(exp[0])
    assert np.array_equal(k[o1], exp[0][o2]) and np.array_equal(c[o1], exp[2][o2]) and c.dtype == np.uint64
    np.testing.assert_allclose(s[o1], exp[1][o2], rtol=1e-9)
    # a WHERE below an aggregate: Aggregate(Selection(TableScan))
    ctx.register_memory("g", [("k", arrays4[0]), ("v", arrays4[1])], batch_size=100_000)
    k2, mx = ctx.sql("SELECT k, MAX(v) FROM g WHERE v < 0.25 GROUP BY k").collect()[0]
    m = arrays4[1] < 0.25
    e2 = O.aggregate([arrays4[0][m], arrays4[1][m]], keys, [AggregateFunction("max", col(1))])
    o1, o2 = np.argsort(k2), np.argsort(e2[0])
    assert np.array_equal(k2[o1], e2[0][o2]) and np.array_equal(mx[o1], e2[1][o2])


def test_wide_table_where_and_pruned_aggregate(ctx):
    # 20 numeric columns: (a) WHERE alone gathers every column (filter.rs:55-57) -> several kernel passes over
    # column groups; (b) SELECT SUM(x) .. WHERE y > 0 uploads two columns and fuses the predicate into the scan
    n = 50_000
    rng = np.random.default_rng(9)
    cols = [rng.random(n) - 0.5 if i % 2 == 0 else rng.integers(-9, 9, n, dtype=np.int64) for i in range(20)]
    names = ["c%d" % i for i in range(20)]
    reg = lambda: ctx.register_memory("w", list(zip(names, cols)), batch_size=20_000)  # noqa: E731  (data sources are one-pass readers)
    reg()
    rel = ctx.sql("SELECT c0, c1, c2, c3, c4, c5, c6, c7, c8, c9, c10, c11, c12, c13, c14, c15, c16, c17, c18, c19 FROM w WHERE c2 > 0.1")
    got = rel.collect()
    m = cols[2] > 0.1
    for i in range(20):
        assert np.array_equal(np.concatenate([b[i] for b in got]), cols[i][m]), i
    reg()
    s, c = ctx.sql("SELECT SUM(c4), COUNT(c4) FROM w WHERE c19 > 0").collect()[0]
    m = cols[19] > 0
    assert c[0] == int(m.sum()) and abs(s[0] - cols[4][m].sum()) <= 1e-9 * abs(cols[4][m].sum())
    reg()
    k, mx = ctx.sql("SELECT c1, MAX(c18) FROM w WHERE c0 < 0.25 GROUP BY c1").collect()[0]
    m = cols[0] < 0.25
    e = O.aggregate([cols[1][m], cols[18][m]], [col(0)], [AggregateFunction("max", col(1))])
    o1, o2 = np.argsort(k), np.argsort(e[0])
    assert np.array_equal(k[o1], e[0][o2]) and np.array_equal(mx[o1], e[1][o2])


def test_big_aggregate_takes_chunked_upload(ctx):
    # > 8 Mi rows: GpuAggregateRelation::next streams the batch through dfgpu_aggregate_update_host
    n = 20_000_000
    arrays, keys, aggs, k_raw = workloads.c4(n, nkeys=50_000)
    ctx.register_memory("big4", [("k", arrays[0]), ("v", arrays[1])])
    k, s, c = ctx.sql("SELECT k, SUM(v), COUNT(v) FROM big4 WHERE v >= 0.5 GROUP BY k").collect()[0]
    m = arrays[1] >= 0.5
    cnt = np.bincount(k_raw[m], minlength=50_000)
    sm = np.bincount(k_raw[m], weights=arrays[1][m], minlength=50_000)
    present = np.nonzero(cnt)[0]
    mixed = workloads.mix_keys(present.astype(np.int64))
    o1, o2 = np.argsort(k), np.argsort(mixed)
    assert np.array_equal(k[o1], mixed[o2]) and np.array_equal(c[o1], cnt[present][o2].astype(np.uint64))
    np.testing.assert_allclose(s[o1], sm[present][o2], rtol=1e-9)


def test_large_batch_takes_pipelined_path(ctx):
    # >= 4 Mi rows, numeric columns: GpuFilterProjectRelation::next goes through dfgpu_filter_project_host
    n = 5_000_000
    arrays, pred, proj = workloads.c3(n, seed=77)
    ctx.register_memory("big", list(zip("abcd", arrays)))
    got = ctx.sql("SELECT a+b, a*b FROM big WHERE b<a").collect()
    a, b = arrays[0], arrays[1]
    m = b < a
    assert len(got) == 1 and np.array_equal(got[0][0], (a + b)[m]) and np.array_equal(got[0][1], (a * b)[m])


def test_error_mapping(ctx):
    register_cities(ctx)
    for sql, code, msg in [
        ("SELECT lat FROM nowhere", A.ERR_GENERAL, "no schema found for table nowhere"),
        ("SELECT lat FROM cities ORDER BY lat", A.ERR_NOT_IMPLEMENTED, "unimplemented!()"),
        ("SELECT lat FROM cities LIMIT 3", A.ERR_NOT_IMPLEMENTED, "unimplemented!()"),
        ("SELECT lat % 2 FROM cities", A.ERR_EXECUTION, "operator: Modulus"),
        ("SELECT lat FROM cities WHERE lat IS NULL", A.ERR_EXECUTION, "expression #1 IS NULL"),
        ("SELECT lat / 0 FROM cities", A.ERR_ARROW, "DivideByZero"),
        ("SELECT lat FROM cities WHERE lat + 1", A.ERR_EXECUTION, "Filter expression did not evaluate to boolean"),
        ("SELECT lat, SUM(lng) FROM cities GROUP BY lat", A.ERR_EXECUTION, "Unsupported GROUP BY data type"),
    ]:
        register_cities(ctx)  # fresh one-pass reader per query
        with pytest.raises(host.ExecutionError) as e:
            rel = ctx.sql(sql)
            rel.next()
        assert e.value.code == code, (sql, e.value.msg)
        assert msg in e.value.msg, (sql, e.value.msg)
