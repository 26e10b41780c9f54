set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_final4.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02_gputest_final4.log
tail -4 gpurun_out/r02_gputest_final4.log | cut -c1-300
