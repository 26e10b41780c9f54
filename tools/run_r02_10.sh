set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_final2.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02_gputest_final2.log
tail -4 gpurun_out/r02_gputest_final2.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_final2.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'clk', d['clocks'], 'speedup_vs_gpu_ref', d.get('speedup_vs_gpu_reference'))
print({k:v['ms'] for k,v in d['kernel_families'].items()})
print('chunk', d['chunk']['ms_per_chunk'], d['chunk']['stage_frames_per_sec'])
"
