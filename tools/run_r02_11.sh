set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py tests/test_sampler.py tests/test_stage.py tests/test_blending.py tests/test_vae_gpu.py -m gpu -q > gpurun_out/r02_gputest_final3.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02_gputest_final3.log | cut -c1-300
timeout 900 python -m pytest tests/test_denoiser_gpu.py tests/test_chain_gpu.py -m gpu -q -k "tiny or chain or graph or recycled or no_controlnet" > gpurun_out/r02_gputest_final3b.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/r02_gputest_final3b.log | cut -c1-300
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_bench_11.json 2> gpurun_out/r02_bench_11.err; echo "bench exit $?"
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_11.json').read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'clk', d['clocks'], d['finite'])
"
