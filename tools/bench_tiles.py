"""Tile-shape study for b200svd_gemm: one large-K problem, every wide N tile, single-CTA vs CTA-pair tiles."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamingt2v_b200 import _lib, ops, packing
from tools.bench_gemm import timeit


def main():
    _lib.init(0)
    dev = torch.device("cuda:0")
    for M, K, N in [(148 * 128 * 4, 4096, 1280), (460800, 1280, 320), (460800, 320, 1280)]:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = packing.pack_linear(torch.randn(N, K) * K ** -0.5, dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for bn in (64, 128, 160, 256):
            for mode in (0, 1):
                ops.gemm_pair_mode(mode)
                ms = timeit(lambda: ops.linear(x, w, None, out=out, bn=bn))
                print(f"M{M} K{K} N{N} bn{bn} pair{mode}: {ms:.3f} ms {2.0 * M * K * N / ms / 1e9:.0f} TF/s", flush=True)
        ms = timeit(lambda: torch.matmul(x, w[0].t()))
        print(f"M{M} K{K} N{N} cublas: {ms:.3f} ms {2.0 * M * K * N / ms / 1e9:.0f} TF/s", flush=True)


if __name__ == "__main__":
    main()
