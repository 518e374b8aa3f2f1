# variants were built as libdfgpu_{b2,b3,e2,e4}.so: filter_project_tma.cu with -DDF_TM_BATCH=2|3 and an early-gather edit (not kept), linked with the other objects
run() { echo "== $*"; env "$@" FP_SHORT=1 timeout 60 python profiles/microbench_fp.py 2>&1 | grep -E "^(c2|sel1|c3|sel99)"; }
D=$PWD/datafusion_archive_b200
run X=1
for v in b2 e2; do for l in 4 6 8; do run DFGPU_LIB=$D/libdfgpu_$v.so DFGPU_FP_LAG=$l; done; done
for v in b3 e4; do for l in 6 8; do run DFGPU_LIB=$D/libdfgpu_$v.so DFGPU_FP_LAG=$l; done; done
