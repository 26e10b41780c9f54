import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "gemm frac", round(d["roofline"]["frac"], 3),
      "eff TF/s", round(d["effective_tflops_per_gpu"], 1), "clocks", d.get("clocks"))
for k, v in d["kernel_families"].items():
    print(f"  {k:11s} n={v['launches']:4d} {v['ms']:8.2f} ms  {v['tflops']:8.2f} TF -> {v['tflops']/max(v['ms'],1e-9)*1e3:7.1f} TF/s   {v['gbytes']:7.1f} GB -> {v['gbytes']/max(v['ms'],1e-9):6.2f} TB/s")
