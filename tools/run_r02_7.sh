set -x
mkdir -p gpurun_out
for v in 3 4; do B200SVD_FA_V=$v timeout 120 python tools/diag_fa.py; done > gpurun_out/r02_diag_fa_odone.txt 2>&1
grep -v Warning gpurun_out/r02_diag_fa_odone.txt | cut -c1-300
for v in 3 4; do B200SVD_FA_V=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "flash_attn" > gpurun_out/r02_gputest_fa_odone$v.log 2>&1; echo "pytest fa$v exit $?"; tail -3 gpurun_out/r02_gputest_fa_odone$v.log | cut -c1-300; done
(for v in 3 4; do B200SVD_FA_V=$v timeout 120 python tools/bench_fa.py; done) 2>&1 | grep "^V="
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_7.json 2> gpurun_out/r02_bench_7.err; echo "bench exit $?"
B200SVD_BN320=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_7_bn320.json 2> gpurun_out/r02_bench_7_bn320.err; echo "bench exit $?"
B200SVD_FA_V=4 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_7_fa4.json 2> gpurun_out/r02_bench_7_fa4.err; echo "bench exit $?"
B200SVD_LEAN_EPI=0 B200SVD_GEGLU_EPI=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_7_oldepi.json 2> gpurun_out/r02_bench_7_oldepi.err; echo "bench exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_7b.json 2> gpurun_out/r02_bench_7b.err; echo "bench exit $?"
python -c "
import json
for f in ('r02_bench_7','r02_bench_7_bn320','r02_bench_7_fa4','r02_bench_7_oldepi','r02_bench_7b'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'finite', d['finite'], {k:(v['ms']) for k,v in d['kernel_families'].items()})
    except Exception as e: print(f, 'ERR', e)
"
