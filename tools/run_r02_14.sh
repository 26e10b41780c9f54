set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_stage.py tests/test_blending.py -m gpu -q -k "uint8 or stage or blending or glue" > gpurun_out/r02_gputest_u8.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/r02_gputest_u8.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
