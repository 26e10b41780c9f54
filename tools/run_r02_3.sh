bash tools/run_r02_5.sh
