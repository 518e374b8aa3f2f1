#!/usr/bin/env python
"""Kernel-time microbenchmark of the filter/project operator for a few query shapes (CUDA events
around the kernel, via dfgpu_profile_*).  usage: microbench_fp.py [rows]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datafusion_archive_b200 import _abi as A  # noqa: E402
from datafusion_archive_b200 import engine  # noqa: E402
from datafusion_archive_b200.expr import col, lit  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
ctx = engine.GpuContext(0)
rng = np.random.default_rng(1)


def run(batch, cases):
    for name, (pred, proj, nbytes) in cases.items():
        for _ in range(2):
            ctx.filter_project(batch, pred, proj).free()
        ctx.profile_enable(True)
        for _ in range(5):
            ctx.filter_project(batch, pred, proj).free()
        ms, k = ctx.profile_get()
        ctx.profile_enable(False)
        print("%-46s %8.3f ms  %7.1f GB/s" % (name, ms / k, nbytes / (ms / k) / 1e6))


a, b = rng.random(n), rng.random(n)
batch = ctx.upload([a, b])
run(batch, {
    "copy  SELECT a": (None, [col(0)], 16.0 * n),
    "c2    SELECT a WHERE a>0.5": (col(0) > lit(0.5), [col(0)], 12.0 * n),
    "sel1% SELECT a WHERE a>0.99": (col(0) > lit(0.99), [col(0)], 8.08 * n),
    "sel99 SELECT a WHERE a>0.01": (col(0) > lit(0.01), [col(0)], 15.92 * n),
    "c3    SELECT a+b,a*b WHERE b<a": (col(1) < col(0), [col(0) + col(1), col(0) * col(1)], 24.0 * n),
    "and2  SELECT a WHERE a>0.5 AND a<2": ((col(0) > lit(0.5)) & (col(0) < lit(2.0)), [col(0)], 12.0 * n),
})
batch.free()
del b
if os.environ.get("FP_SHORT"):
    raise SystemExit
ki = rng.integers(-1000, 1000, n, dtype=np.int64)
vf = rng.random(n).astype(np.float32)
batch = ctx.upload([ki, vf, a])
sel_deep = float(np.count_nonzero(a * a < 0.3)) / n
run(batch, {
    "i64   SELECT k WHERE k>0  (interpreter)": (col(0) > lit(0), [col(0)], 12.0 * n),
    "f32   SELECT v WHERE v<0.5  (interpreter)": (col(1) < lit(0.5, A.FLOAT32), [col(1)], 6.0 * n),
    "mixed SELECT a WHERE k>0  (interpreter)": (col(0) > lit(0), [col(2)], 20.0 * n),
    "deep  SELECT (a+a)*(a-1)/(a+2) WHERE a*a<0.3": ((col(2) * col(2)) < lit(0.3), [(col(2) + col(2)) * (col(2) - lit(1.0)) / (col(2) + lit(2.0))],
                                                     8.0 * n + 8.0 * n * sel_deep),
})
batch.free()
ctx.close()
