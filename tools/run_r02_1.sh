set -x
python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_gputest_1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_gputest_1.log
tail -5 gpurun_out/r02_gputest_1.log
B200SVD_BENCH_SHAPES=gpurun_out/r02_shapes_1.txt python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1.json 2> gpurun_out/r02_bench_1.err; echo "bench exit $?"
tail -c 3000 gpurun_out/r02_bench_1.json
B200SVD_NO_GRAPH=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_1_nograph.json 2> gpurun_out/r02_bench_1_nograph.err; echo "bench exit $?"
python -c "
import json
for f in ('gpurun_out/r02_bench_1.json','gpurun_out/r02_bench_1_nograph.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['e2e']['ms_per_step'], d.get('cuda_graph'), d.get('speedup_vs_gpu_reference'))
    except Exception as e: print(f, 'ERR', e)
"
