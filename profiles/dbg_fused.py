import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from datafusion_archive_b200 import engine, workloads, host, _abi as A
from datafusion_archive_b200.expr import AggregateFunction, col, lit
ctx = engine.GpuContext(0)
n = 1_000_000
arrays, keys, aggs, _ = workloads.c5(n, nkeys=20_000)
aggs = aggs + [AggregateFunction("count", col(1))]
def run(name, arrays, keys, aggs, pred, nb):
    bounds = [int(x) for x in np.linspace(0, len(arrays[0]), nb + 1)]
    batches = [ctx.upload([a[bounds[i]:bounds[i + 1]] for a in arrays]) for i in range(nb)]
    ctx.sync(); t = time.time()
    r = ctx.aggregate(batches, keys, aggs, 0, pred=pred)
    ctx.sync(); dt = time.time() - t
    print("%-40s nb=%d groups=%d  %.3f s" % (name, nb, r.nrows, dt), flush=True)
    r.free()
    for b in batches: b.free()
preds = [("v<0.25", col(1) < lit(0.25)), ("0.1<v<0.9", (col(1) > lit(0.1)) & (col(1) < lit(0.9))), ("or", (col(1) < lit(0.05)) | (col(1) >= lit(0.95))),
         ("v*2<0.5", (col(1) * lit(2.0)) < lit(0.5)), ("none", col(1) < lit(-1.0))]
for nm, p in preds:
    for nb in [1, 3]:
        run(nm, arrays, keys, aggs, p, nb)
rng = np.random.default_rng(77)
k = rng.integers(0, 5000, n, dtype=np.int64); w = rng.integers(-100, 100, n, dtype=np.int64); v = rng.random(n)
run("exprkeys", [k, w, v], [col(0) + lit(7)], [AggregateFunction("sum", col(2) * lit(3.0)), AggregateFunction("max", col(1)), AggregateFunction("count", col(2))],
    (col(1) > lit(-20)) & (col(2) < lit(0.75)), 1)
k32 = rng.integers(0, 3000, n - 1, dtype=np.int32); v32 = rng.random(n - 1).astype(np.float32)
run("i32/f32", [k32, v32], [col(0)], [AggregateFunction("min", col(1)), AggregateFunction("max", col(1)), AggregateFunction("count", col(1))], col(1) >= lit(0.5, A.FLOAT32), 1)
ctx.close()
# host-layer no-GROUP-BY with WHERE
h = host.ExecutionContext(0)
n = 50_000
rng = np.random.default_rng(9)
cols = [rng.random(n) - 0.5 if i % 2 == 0 else rng.integers(-9, 9, n, dtype=np.int64) for i in range(20)]
h.register_memory("w", list(zip(["c%d" % i for i in range(20)], cols)), batch_size=20_000)
print(h.plan("SELECT SUM(c4), COUNT(c4) FROM w WHERE c19 > 0"))
print(h.sql("SELECT SUM(c4), COUNT(c4) FROM w WHERE c19 > 0").collect(), int((cols[19] > 0).sum()), cols[4][cols[19] > 0].sum())
h.register_memory("w2", [("a", cols[4]), ("b", cols[19])], batch_size=20_000)
print(h.sql("SELECT SUM(a), COUNT(a) FROM w2 WHERE b > 0").collect())
print(h.sql("SELECT COUNT(a) FROM w2").collect())
h.close()
