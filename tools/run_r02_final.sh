# The commands behind the committed round-2 measurements (run on a B200 box through gpurun from the repo root):
#   profiles/r02_gputest_final*_summary.txt, r02_bench_final.json, r02_bench_reference_arm.json, r02_launches_final*,
#   r02_ncu_*_details.txt
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_gputest_final.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/r02_gputest_final.log
tail -4 gpurun_out/r02_gputest_final.log | cut -c1-300
B200SVD_BENCH_SHAPES=gpurun_out/r02_shapes_final.txt timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_reference_arm.err; echo "ref arm exit $?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
bash tools/run_r02_prof.sh > gpurun_out/r02_prof.log 2>&1; tail -25 gpurun_out/r02_prof.log | cut -c1-250
