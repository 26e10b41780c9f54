set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02_bench_final3.json 2> gpurun_out/r02_bench_final3.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_final3.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'clk', d['clocks'], 'speedup_vs_gpu_ref', d.get('speedup_vs_gpu_reference'))
print('chunk', json.dumps(d['chunk'])[:600])
print('first', json.dumps(d['first_chunk'])[:900])
print('200f', json.dumps(d.get('streamingsvd_stage_200_frames')))
"
tail -3 gpurun_out/r02_bench_final3.err
