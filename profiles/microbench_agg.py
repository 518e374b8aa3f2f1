#!/usr/bin/env python
"""Kernel-time microbenchmark of the aggregate operator (CUDA events around k_hash_agg / k_reduce).
usage: microbench_agg.py [rows] [groups,groups,...]   (DFGPU_TRACE=1 prints host-side phase timings)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datafusion_archive_b200 import engine, workloads  # noqa: E402
from datafusion_archive_b200.expr import AggregateFunction, col  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
ctx = engine.GpuContext(0)
v = np.random.default_rng(47).random(n)
group_counts = [int(float(x)) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [10, 1000, 100_000, 1_000_000, 10_000_000]
for nkeys in group_counts:
    k = workloads.mix_keys(np.random.default_rng(46).integers(0, nkeys, n, dtype=np.int64))
    b = ctx.upload([k, v])
    for name, aggs in [("sum,count", [AggregateFunction("sum", col(1)), AggregateFunction("count", col(1))]),
                       ("min,max,sum", [AggregateFunction("min", col(1)), AggregateFunction("max", col(1)), AggregateFunction("sum", col(1))])]:
        ctx.aggregate(b, [col(0)], aggs).free()
        ctx.profile_enable(True)
        walls = []
        for _ in range(5):
            ctx.timer_start()
            r = ctx.aggregate(b, [col(0)], aggs)
            g = r.nrows
            r.free()
            walls.append(ctx.timer_stop())
        wall = float(np.median(walls))
        ms, kn = ctx.profile_get()
        ctx.profile_enable(False)
        print("groups=%-9d %-12s scan kernels %7.3f ms/op (%d launches/op)  whole op %7.3f ms (median of 5; max %.3f)  %6.1f GB/s"
              % (g, name, ms / 5, kn // 5, wall, max(walls), 16.0 * n / wall / 1e6))
    b.free()
b = ctx.upload([v])
aggs = [AggregateFunction("min", col(0)), AggregateFunction("max", col(0)), AggregateFunction("sum", col(0)), AggregateFunction("count", col(0))]
ctx.aggregate(b, [], aggs).free()
ctx.profile_enable(True)
for _ in range(3):
    ctx.aggregate(b, [], aggs).free()
ms, kn = ctx.profile_get()
print("no GROUP BY min,max,sum,count %8.3f ms  %6.1f GB/s" % (ms / kn, 8.0 * n / (ms / kn) / 1e6))
ctx.close()
