#!/bin/bash
# (libdfgpu_pipe.so: an experiment build of filter_project_tma.cu with pipelined scan warps; the code was measured, documented in profiles/r02_history.md and removed)
# pipelined scan warps (-DDF_TM_PIPE build as libdfgpu_pipe.so) against the batched form, over the lag
run() { echo "== $*"; env "$@" FP_SHORT=1 timeout 60 python profiles/microbench_fp.py 2>&1 | grep -E "^(c2|sel1|c3|sel99)"; }
D=$PWD/datafusion_archive_b200
run X=1
for l in 3 4 5 6 8; do run DFGPU_LIB=$D/libdfgpu_pipe.so DFGPU_FP_LAG=$l; done
DFGPU_LIB=$D/libdfgpu_pipe.so timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c2 or c3 or lean or selectiv or fuzz or dtypes" 2>&1 | tail -3
