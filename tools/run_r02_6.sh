set -x
mkdir -p gpurun_out
export B200SVD_LEAN_EPI=1 B200SVD_GEGLU_EPI=1
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x > gpurun_out/r02_gputest_gemm2.log 2>&1; echo "pytest gemm exit $?"; tail -4 gpurun_out/r02_gputest_gemm2.log | cut -c1-300
for v in 6 7; do B200SVD_FA_V=$v timeout 120 python tools/diag_fa.py; done > gpurun_out/r02_diag_fa_pp.txt 2>&1
grep -v Warning gpurun_out/r02_diag_fa_pp.txt | cut -c1-300
(for v in 3 6 7; do B200SVD_FA_V=$v timeout 120 python tools/bench_fa.py; done; B200SVD_FA_V=7 B200SVD_FA_POLY=1 timeout 120 python tools/bench_fa.py; B200SVD_FA_V=7 B200SVD_FA_POLY=2 timeout 120 python tools/bench_fa.py) > gpurun_out/r02_bench_fa_pp.txt 2>&1
grep "^V=" gpurun_out/r02_bench_fa_pp.txt
for v in 6 7; do B200SVD_FA_V=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "flash_attn" > gpurun_out/r02_gputest_fa$v.log 2>&1; echo "pytest fa$v exit $?"; tail -3 gpurun_out/r02_gputest_fa$v.log | cut -c1-300; done
timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_chain_gpu.py tests/test_vae_gpu.py -m gpu -q -s > gpurun_out/r02_gputest_parity_newepi.log 2>&1; echo "pytest parity exit $?"; grep -E "rel_l2=|\[chain\]|full size|passed|failed" gpurun_out/r02_gputest_parity_newepi.log | grep -v "   " | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_6.json 2> gpurun_out/r02_bench_6.err; echo "bench exit $?"
B200SVD_BN320=1 B200SVD_BENCH_SHAPES=gpurun_out/r02_shapes_6_bn320.txt timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_6_bn320.json 2> gpurun_out/r02_bench_6_bn320.err; echo "bench exit $?"
B200SVD_BN320=1 B200SVD_FA_V=7 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_6_bn320_fa7.json 2> gpurun_out/r02_bench_6_bn320_fa7.err; echo "bench exit $?"
B200SVD_BN320=1 B200SVD_FA_V=6 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_6_bn320_fa6.json 2> gpurun_out/r02_bench_6_bn320_fa6.err; echo "bench exit $?"
python -c "
import json
for f in ('r02_bench_6','r02_bench_6_bn320','r02_bench_6_bn320_fa7','r02_bench_6_bn320_fa6'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'finite', d['finite'], {k:(v['ms']) for k,v in d['kernel_families'].items()})
    except Exception as e: print(f, 'ERR', e)
"
