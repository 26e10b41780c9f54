set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_n4_latency.json 2> gpurun_out/r02_bench_n4_latency.err; echo "n4 latency exit $?"
python -c "
import json
txt=open('gpurun_out/r02_bench_n4_latency.json').read().strip().splitlines()
d=json.loads([l for l in txt if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','clocks','collective','replicas')}); print(d['e2e']); print(d['config']['parallelism'])
"
tail -3 gpurun_out/r02_bench_n4_latency.err
