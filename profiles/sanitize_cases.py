#!/usr/bin/env python
"""Workload for compute-sanitizer (memcheck / racecheck): one pass through every hand-rolled synchronisation
protocol of the engine at sizes a sanitizer finishes in minutes, each result checked against numpy.
  compute-sanitizer --tool racecheck python profiles/sanitize_cases.py
  compute-sanitizer --tool memcheck  python profiles/sanitize_cases.py
Covers: the mbarrier / cp.async.bulk pipeline of k_filter_project_tma (full, ragged and single tiles, lagged
scan, dual ring), the direct filter kernel, the CAS / RED table of k_hash_agg_lean / _plain / interpreter, table
growth with overflow replay, the shared-memory front tables, wide-key slots (busy / ready publication) and the
chunked host pipelines."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datafusion_archive_b200 import engine, workloads  # noqa: E402
from datafusion_archive_b200.expr import AggregateFunction, col, lit  # noqa: E402

ctx = engine.GpuContext(0)
rng = np.random.default_rng(1)


def fp(arrays, pred, proj):
    b = ctx.upload(arrays)
    r = ctx.filter_project(b, pred, proj)
    out = r.columns()
    r.free(); b.free()
    return out


def agg(arrays, keys, aggs, nb=1, pred=None, expected=0):
    n = len(arrays[0])
    bounds = [int(x) for x in np.linspace(0, n, nb + 1)]
    bs = [ctx.upload([a[bounds[i]:bounds[i + 1]] for a in arrays]) for i in range(nb)]
    r = ctx.aggregate(bs, keys, aggs, expected, pred=pred)
    out = r.columns()
    r.free()
    for b in bs:
        b.free()
    return out


# 1. TMA filter pipeline: sizes around tile boundaries, C2 and C3 shapes
for n in [1, 4095, 4096, 4097, 700_001]:
    a = rng.random(n)
    out = fp([a], col(0) > lit(0.5), [col(0)])[0]
    assert np.array_equal(out, a[a > 0.5]), n
arrays, pred, proj = workloads.c3(300_000)
o = fp(arrays, pred, proj)
m = arrays[1] < arrays[0]
assert np.array_equal(o[0], (arrays[0] + arrays[1])[m]) and np.array_equal(o[1], (arrays[0] * arrays[1])[m])
# lean consumer loop: Int64 comparison against a column, 64-bit integer arithmetic, two projections; generic FAST loop: two terms
ki = rng.integers(-50, 50, 300_001, dtype=np.int64)
kj = rng.integers(-50, 50, 300_001, dtype=np.int64)
o = fp([ki, kj], col(0) <= col(1), [col(0) * col(1), col(1)])
assert np.array_equal(o[0], (ki * kj)[ki <= kj]) and np.array_equal(o[1], kj[ki <= kj])
a1 = rng.random(300_001)
o = fp([a1], (col(0) > lit(0.25)) & (col(0) < lit(0.75)), [col(0)])
assert np.array_equal(o[0], a1[(a1 > 0.25) & (a1 < 0.75)])
# generic interpreter (direct kernel shapes)
o = fp([arrays[0]], (col(0) * col(0)) < lit(0.3), [(col(0) + col(0)) * (col(0) - lit(1.0)) / (col(0) + lit(2.0))])
a0 = arrays[0]
assert np.array_equal(o[0], ((a0 + a0) * (a0 - 1.0) / (a0 + 2.0))[a0 * a0 < 0.3])
print("filter/project ok", flush=True)

# 2. hash aggregate: lean (SUM, COUNT / MIN, MAX, SUM), plain (with WHERE), interpreter (expression key)
n = 400_000
k = workloads.mix_keys(rng.integers(0, 3000, n, dtype=np.int64))
v = rng.random(n)
uk, inv = np.unique(k, return_inverse=True)
got = agg([k, v], [col(0)], [AggregateFunction("sum", col(1)), AggregateFunction("count", col(1))], nb=2)
o = np.argsort(got[0])
assert np.array_equal(got[0][o], uk) and np.array_equal(got[2][o], np.bincount(inv).astype(np.uint64))
got = agg([k, v], [col(0)], [AggregateFunction("min", col(1)), AggregateFunction("max", col(1)), AggregateFunction("sum", col(1))])
o = np.argsort(got[0])
mn = np.full(len(uk), np.inf); np.minimum.at(mn, inv, v)
assert np.array_equal(got[1][o], mn)
got = agg([k, v], [col(0)], [AggregateFunction("max", col(1))], pred=col(1) < lit(0.5))
assert len(got[0]) == len(np.unique(k[v < 0.5]))
got = agg([k, v], [col(0) + lit(1)], [AggregateFunction("sum", col(1) * lit(2.0))])
assert len(got[0]) == len(uk)
print("hash aggregate ok", flush=True)

# 3. table growth with overflow replay (more distinct keys than half the initial table)
n = 2_300_000
kk = workloads.mix_keys(np.arange(n, dtype=np.int64))
got = agg([kk, np.ones(n)], [col(0)], [AggregateFunction("count", col(1))])
assert len(got[0]) == n and int(got[1].sum()) == n
print("growth ok", flush=True)

# 4. front tables (few groups, > 4 Mi rows so that the sampled prefix switches them on)
n = 4_500_000
for g in [3, 500]:
    kf = workloads.mix_keys(rng.integers(0, g, n, dtype=np.int64))
    vf = rng.random(n)
    got = agg([kf, vf], [col(0)], [AggregateFunction("sum", col(1)), AggregateFunction("count", col(1)), AggregateFunction("max", col(1))])
    assert len(got[0]) == g and int(got[2].sum()) == n
print("front tables ok", flush=True)

# 5. wide keys: busy / ready publication under contention, and growth by moving slots
n = 600_000
h1 = rng.integers(0, 3, n, dtype=np.int64) * (2 ** 40)
h2 = rng.integers(0, 2, n, dtype=np.int64) - 1
got = agg([h1, h2, rng.random(n)], [col(0), col(1)], [AggregateFunction("count", col(2))])
assert len(got[0]) == 6 and int(got[2].sum()) == n
s1 = ["k%d" % i for i in rng.integers(0, 50, 100_000)]
k3 = rng.integers(0, 4, 100_000, dtype=np.int32)
got = agg([s1, k3, rng.random(100_000)], [col(0), col(1)], [AggregateFunction("sum", col(2))], nb=2)
assert len(got[0]) == len(set(zip(s1, k3.tolist())))
print("wide keys ok", flush=True)

# 6. chunked host pipelines
a = rng.random(9_000_000)
r = ctx.filter_project_host([a], col(0) > lit(0.5), [col(0)], chunk_rows=2_000_000)
assert r.nrows == int((a > 0.5).sum())
r.free()
kh = workloads.mix_keys(rng.integers(0, 1000, 9_000_000, dtype=np.int64))
r = ctx.aggregate_host([kh, a], [col(0)], [AggregateFunction("sum", col(1))], chunk_rows=2_000_000)
assert r.nrows == 1000
r.free()
print("host pipelines ok", flush=True)
ctx.close()
print("SANITIZE_CASES_OK")
