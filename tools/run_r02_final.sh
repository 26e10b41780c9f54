set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02_gputest_final.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02_gputest_final.log
tail -4 gpurun_out/r02_gputest_final.log | cut -c1-300
B200SVD_BENCH_SHAPES=gpurun_out/r02_shapes_final.txt timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'clk', d['clocks'], 'settle', d.get('settle_steps'), 'speedup_vs_gpu_ref', d.get('speedup_vs_gpu_reference'))
print('fam', json.dumps(d['kernel_families']))
print('chunk', json.dumps(d.get('chunk')))
print('gpu_ref', json.dumps(d.get('gpu_reference')))
print('cpu', json.dumps(d.get('cpu_baseline')))
print('roof', json.dumps(d.get('roofline')))
"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref arm exit $?"; tail -c 600 gpurun_out/r02_bench_reference_arm.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/run_r02_prof.sh > gpurun_out/r02_prof.log 2>&1; tail -25 gpurun_out/r02_prof.log | cut -c1-250
