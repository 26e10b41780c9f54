set -x
mkdir -p gpurun_out
KREG='regex:mtgemm|flash_attn|gn_|layernorm|pixel_attn|small_attn|nchw|nhwc|upsample|add_silu|add_rows|copy2d|timestep|sampler_|apm_mix'
# launch list of ONE step of the same bench command (eager launches: graph replay hides the kernels from -k filters)
B200SVD_NO_GRAPH=1 timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -s 1100 -c 1100 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-gpu-reference --no-chunk > gpurun_out/r02_launches_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches.csv gpurun_out/r02_launches_summary.txt | head -30
# --set full captures of the dominant kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mtgemm -s 2 -c 1 -f -o gpurun_out/r02_ncu_geglu python tools/one_gemm.py geglu 320 2560 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mtgemm -s 2 -c 1 -f -o gpurun_out/r02_ncu_conv python tools/one_gemm.py conv > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flash_attn -s 1 -c 1 -f -o gpurun_out/r02_ncu_fa python tools/one_fa.py 50 9216 5 > /dev/null 2>&1
for f in geglu conv fa; do ncu -i gpurun_out/r02_ncu_$f.ncu-rep --page details > gpurun_out/r02_ncu_${f}_details.txt 2>&1; ncu -i gpurun_out/r02_ncu_$f.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hdr=rows[0]
want=['dram__bytes_read.sum','dram__bytes_write.sum','gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed.sum','launch__registers_per_thread','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print({k:d.get(k) for k in want if k in d})
"; done
ls -la gpurun_out/*.ncu-rep
