"""Isolate epilogue costs of b200svd_gemm: same GEMM shape with different epilogue features."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamingt2v_b200 import _lib, ops, packing
from tools.bench_gemm import timeit


def main():
    _lib.init(0)
    dev = torch.device("cuda:0")
    M = 460800
    for K, N in [(320, 2560), (320, 320), (320, 960), (1280, 320)]:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        wt = torch.randn(N, K) * K ** -0.5
        b = torch.randn(N, device=dev)
        w = packing.pack_linear(wt, dev)
        fl = 2.0 * M * K * N
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        res = torch.randn(M, N, device=dev).to(torch.bfloat16)
        fv = torch.randn(50, N, device=dev)
        cases = {
            "plain": lambda: ops.linear(x, w, None, out=out),
            "bias": lambda: ops.linear(x, w, b, out=out),
            "bias+silu": lambda: ops.linear(x, w, b, out=out, act=ops.ACT_SILU),
            "bias+gelu": lambda: ops.linear(x, w, b, out=out, act=ops.ACT_GELU),
            "bias+res1": lambda: ops.linear(x, w, b, out=out, res1=res),
            "bias+res1+fvec": lambda: ops.linear(x, w, b, out=out, res1=res, fvec=fv, rows_per_frame=9216),
            "bias+res2": lambda: ops.linear(x, w, b, out=out, res1=res, res2=res, s2=0.5),
        }
        if N % 320 == 0 and N >= 2560:
            wp, bp, bn = packing.pack_geglu(wt, b.cpu(), dev)
            outg = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)
            cases["geglu"] = lambda: ops.linear(x, wp, bp, out=outg, act=ops.ACT_GEGLU, bn=bn)
        for name, fn in cases.items():
            ms = timeit(fn, iters=5, warmup=2)
            print(f"K{K} N{N} {name:16s} {ms:7.3f} ms {fl / ms / 1e9:7.1f} TF/s", flush=True)
        ms = timeit(lambda: torch.matmul(x, w[0].t()), iters=5, warmup=2)
        print(f"K{K} N{N} {'cublas':16s} {ms:7.3f} ms {fl / ms / 1e9:7.1f} TF/s", flush=True)


if __name__ == "__main__":
    main()
