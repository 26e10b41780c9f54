set -x
# FA variants first (short)
for v in 3 4; do for p in 0 1 2 3; do if [ $v = 3 ] && [ $p -gt 0 ]; then continue; fi; B200SVD_FA_V=$v B200SVD_FA_POLY=$p timeout 300 python tools/bench_fa.py; done; done > gpurun_out/r02_bench_fa_v4.txt 2>&1
cat gpurun_out/r02_bench_fa_v4.txt | grep -v Warning
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_gputest_2.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_gputest_2.log
tail -8 gpurun_out/r02_gputest_2.log
B200SVD_BENCH_SHAPES=gpurun_out/r02_shapes_2.txt timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_2.json 2> gpurun_out/r02_bench_2.err; echo "bench exit $?"
tail -c 1500 gpurun_out/r02_bench_2.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_2.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'graph', d.get('cuda_graph'), 'speedup_vs_gpu_ref', d.get('speedup_vs_gpu_reference'))
print('fam', json.dumps(d['kernel_families']))
print('chunk', json.dumps(d.get('chunk')))
print('gpu_ref', json.dumps(d.get('gpu_reference')))
print('cpu', json.dumps(d.get('cpu_baseline')))
"
