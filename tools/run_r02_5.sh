set -x
mkdir -p gpurun_out
export B200SVD_LEAN_EPI=1 B200SVD_GEGLU_EPI=1 B200SVD_GN_FUSE=1
timeout 1200 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r02_gputest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -4 gpurun_out/r02_gputest_gemm.log
for v in 3 4 5 6 7; do B200SVD_FA_V=$v timeout 300 python tools/diag_fa.py; done > gpurun_out/r02_diag_fa.txt 2>&1
grep -v Warning gpurun_out/r02_diag_fa.txt | cut -c1-400
for v in 3 4 5 6 7; do B200SVD_FA_V=$v timeout 300 python tools/bench_fa.py; done; B200SVD_FA_V=7 B200SVD_FA_POLY=1 timeout 300 python tools/bench_fa.py; B200SVD_FA_V=7 B200SVD_FA_POLY=2 timeout 300 python tools/bench_fa.py > gpurun_out/r02_bench_fa_v5.txt 2>&1
grep "^V=" gpurun_out/r02_bench_fa_v5.txt
for v in 5 6 7; do B200SVD_FA_V=$v timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "flash_attn" > gpurun_out/r02_gputest_fa$v.log 2>&1; echo "pytest fa$v exit $?"; tail -3 gpurun_out/r02_gputest_fa$v.log; done
timeout 300 python -m pytest tests/test_blending.py tests/test_kernels_gpu.py -m gpu -q > gpurun_out/r02_gputest_misc.log 2>&1; echo "pytest misc exit $?"; tail -3 gpurun_out/r02_gputest_misc.log
B200SVD_BENCH_SHAPES=gpurun_out/r02_shapes_5.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_5.json 2> gpurun_out/r02_bench_5.err; echo "bench exit $?"
B200SVD_BN320=1 B200SVD_BENCH_SHAPES=gpurun_out/r02_shapes_5_bn320.txt timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_5_bn320.json 2> gpurun_out/r02_bench_5_bn320.err; echo "bench exit $?"
B200SVD_GN_FUSE=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_5_nognfuse.json 2> gpurun_out/r02_bench_5_nognfuse.err; echo "bench exit $?"
B200SVD_LEAN_EPI=0 B200SVD_GEGLU_EPI=0 B200SVD_GN_FUSE=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_5_oldepi.json 2> gpurun_out/r02_bench_5_oldepi.err; echo "bench exit $?"
python -c "
import json
for f in ('r02_bench_5','r02_bench_5_bn320','r02_bench_5_nognfuse','r02_bench_5_oldepi'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], {k:(v['ms']) for k,v in d['kernel_families'].items()})
    except Exception as e: print(f, 'ERR', e)
"
timeout 900 python tools/bench_vs_libs.py > gpurun_out/r02_bench_vs_libs.txt 2>&1; tail -14 gpurun_out/r02_bench_vs_libs.txt
B200SVD_BN320=1 timeout 900 python tools/bench_vs_libs.py > gpurun_out/r02_bench_vs_libs_bn320.txt 2>&1; tail -14 gpurun_out/r02_bench_vs_libs_bn320.txt
head -30 gpurun_out/r02_shapes_5.txt
