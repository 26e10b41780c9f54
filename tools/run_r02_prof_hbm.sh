set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:gn_stats_kernel|gn_apply_kernel|layernorm|pixel_attn" -s 4 -c 4 -f -o gpurun_out/r02_ncu_hbm python tools/one_norm.py > /dev/null 2>&1
ncu -i gpurun_out/r02_ncu_hbm.ncu-rep --page details > gpurun_out/r02_ncu_hbm_details.txt 2>&1
ncu -i gpurun_out/r02_ncu_hbm.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print(d.get('Kernel Name','')[:40], 'us', d.get('gpu__time_duration.sum'), 'rd', d.get('dram__bytes_read.sum'), 'wr', d.get('dram__bytes_write.sum'), 'dram%', d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'))
"
