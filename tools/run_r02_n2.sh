set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_bench_n2_latency.json 2> gpurun_out/r02_bench_n2_latency.err; echo "n2 latency exit $?"
tail -c 1200 gpurun_out/r02_bench_n2_latency.json
tail -3 gpurun_out/r02_bench_n2_latency.err
