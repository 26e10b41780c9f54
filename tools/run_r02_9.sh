set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
B200SVD_FA_V=5 timeout 120 python tools/diag_fa.py > gpurun_out/r02_diag_fa_v5.txt 2>&1; echo "diag5 exit $?"
grep -v Warning gpurun_out/r02_diag_fa_v5.txt | cut -c1-300
B200SVD_FA_V=5 timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "flash_attn" > gpurun_out/r02_gputest_fa5b.log 2>&1; echo "pytest fa5 exit $?"; tail -3 gpurun_out/r02_gputest_fa5b.log | cut -c1-300
(for v in 4 5; do B200SVD_FA_V=$v timeout 120 python tools/bench_fa.py; done) > gpurun_out/r02_bench_fa_v5b.txt 2>&1; grep "^V=" gpurun_out/r02_bench_fa_v5b.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_9.json 2> gpurun_out/r02_bench_9.err; echo "bench exit $?"
B200SVD_FA_V=5 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_9_fa5.json 2> gpurun_out/r02_bench_9_fa5.err; echo "bench exit $?"
B200SVD_BN320=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_9_bn320m2.json 2> gpurun_out/r02_bench_9_bn320m2.err; echo "bench exit $?"
python -c "
import json
for f in ('r02_bench_9','r02_bench_9_fa5','r02_bench_9_bn320m2'):
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],2), 'clk', d['clocks']['sm_mhz'], 'finite', d['finite'], {k:(v['ms']) for k,v in d['kernel_families'].items()})
    except Exception as e: print(f, 'ERR', e)
"
