import ast
for l in open("gpurun_out/bg.log"):
    if l.startswith("{"):
        d = ast.literal_eval(l)
        print(d["name"], round(d["ms"], 3), round(d["tflops"]), "lib:", round(d.get("cublas_tflops", d.get("cudnn_tflops", 0))))
    else:
        print(l.strip()[:200])
