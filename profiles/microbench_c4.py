#!/usr/bin/env python
"""C4 / C5 kernel time only (A/B of libdfgpu builds via DFGPU_LIB)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datafusion_archive_b200 import engine, workloads
ctx = engine.GpuContext(0)
for name, wl in [("c4 1e5 keys sum,count", workloads.c4(100_000_000)), ("c5 1e6 keys min,max,sum", workloads.c5(100_000_000))]:
    arrays, keys, aggs, _ = wl
    b = ctx.upload(arrays)
    ctx.aggregate(b, keys, aggs).free()
    ctx.profile_enable(True); ctx.timer_start()
    for _ in range(3):
        ctx.aggregate(b, keys, aggs).free()
    wall = ctx.timer_stop() / 3
    ms, kn = ctx.profile_get(); ctx.profile_enable(False)
    print("%-26s scan kernels %7.3f ms/op  whole op %7.3f ms" % (name, ms / 3, wall))
    b.free()
ctx.close()
