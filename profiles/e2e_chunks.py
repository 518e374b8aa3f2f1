#!/usr/bin/env python
"""e2e step time of dfgpu_filter_project_host on C2 (1e8 pinned f64 rows in, selected rows to pinned host) per chunk size."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from datafusion_archive_b200 import engine  # noqa: E402
from datafusion_archive_b200.expr import col, lit  # noqa: E402

n = 100_000_000
ctx = engine.GpuContext(0)
pin = engine.PinnedBuffer((n,), np.float64)
np.random.default_rng(42).random(n, out=pin.array)
for chunk in [8 << 20, 4 << 20, 2 << 20, 1 << 20, 16 << 20]:
    for _ in range(2):
        ctx.filter_project_host([pin.array], col(0) > lit(0.5), [col(0)], chunk_rows=chunk).free()
    t = time.perf_counter()
    for _ in range(6):
        ctx.filter_project_host([pin.array], col(0) > lit(0.5), [col(0)], chunk_rows=chunk).free()
    print("chunk %3d Mi rows: %.2f ms per step" % (chunk >> 20, (time.perf_counter() - t) / 6 * 1e3), flush=True)
