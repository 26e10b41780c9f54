set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nccl_pair_gpu.py -m gpu -q -x -s > gpurun_out/r02_gputest_nccl_pair.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/r02_gputest_nccl_pair.log | cut -c1-300
