set -x
for v in 3 4 5; do B200SVD_FA_V=$v timeout 300 python tools/diag_fa.py; done > gpurun_out/r02_diag_fa.txt 2>&1
grep -v Warning gpurun_out/r02_diag_fa.txt
for v in 3 4 5; do B200SVD_FA_V=$v timeout 300 python tools/bench_fa.py; done > gpurun_out/r02_bench_fa_v5.txt 2>&1
grep "^V=" gpurun_out/r02_bench_fa_v5.txt
B200SVD_FA_V=5 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "flash_attn" > gpurun_out/r02_gputest_fa5.log 2>&1; echo "pytest fa5 exit $?"; tail -3 gpurun_out/r02_gputest_fa5.log
timeout 300 python -m pytest tests/test_blending.py -m gpu -q > gpurun_out/r02_gputest_blend.log 2>&1; echo "pytest blend exit $?"; tail -3 gpurun_out/r02_gputest_blend.log
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -m gpu -q -x -k "not flash_attn" > gpurun_out/r02_gputest_3.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/r02_gputest_3.log
timeout 900 python tools/bench_vs_libs.py > gpurun_out/r02_bench_vs_libs.txt 2>&1; tail -16 gpurun_out/r02_bench_vs_libs.txt
B200SVD_FA_V=3 B200SVD_BENCH_SHAPES=gpurun_out/r02_shapes_3.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_3.json 2> gpurun_out/r02_bench_3.err; echo "bench exit $?"
B200SVD_FA_V=3 B200SVD_GN_FUSE=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_3_nognfuse.json 2> gpurun_out/r02_bench_3_nognfuse.err; echo "bench exit $?"
B200SVD_FA_V=3 B200SVD_LEAN_EPI=0 B200SVD_GEGLU_EPI=0 B200SVD_GN_FUSE=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_3_oldepi.json 2> gpurun_out/r02_bench_3_oldepi.err; echo "bench exit $?"
python -c "
import json
for f in ('gpurun_out/r02_bench_3.json','gpurun_out/r02_bench_3_nognfuse.json','gpurun_out/r02_bench_3_oldepi.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], {k:(v['ms']) for k,v in d['kernel_families'].items()})
"
head -40 gpurun_out/r02_shapes_3.txt
