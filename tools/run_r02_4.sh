set -x
for v in 3 4 5; do B200SVD_FA_V=$v timeout 300 python tools/diag_fa.py; done > gpurun_out/r02_diag_fa2.txt 2>&1
grep -v Warning gpurun_out/r02_diag_fa2.txt
for v in 3 4 5; do B200SVD_FA_V=$v timeout 300 python tools/bench_fa.py; done > gpurun_out/r02_bench_fa_v5.txt 2>&1
grep "^V=" gpurun_out/r02_bench_fa_v5.txt
B200SVD_FA_V=5 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "flash_attn" > gpurun_out/r02_gputest_4.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r02_gputest_4.log
