"""Pair tiles (mode 2) vs clusters of two pairs with multicast A (mode 3) on the denoiser's dominant shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamingt2v_b200 import _lib, ops, packing
from tools.bench_gemm import timeit


def main():
    _lib.init(0)
    dev = torch.device("cuda:0")
    for M, K, N in [(460800, 1280, 320), (460800, 320, 320), (115200, 640, 1920), (28800, 1280, 10240),
                    (460800, 320, 2560), (115200, 2560, 640)]:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = packing.pack_linear(torch.randn(N, K) * K ** -0.5, dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for mode in (2, 3):
            ops.gemm_pair_mode(mode)
            ms = timeit(lambda: ops.linear(x, w, None, out=out))
            print(f"lin M{M} K{K} N{N} mode{mode}: {ms:.3f} ms {2.0 * M * K * N / ms / 1e9:.0f} TF/s", flush=True)
    for Nf, H, W, Cin, Cout in [(50, 72, 128, 320, 320), (50, 72, 128, 640, 320), (50, 36, 64, 640, 640),
                                (50, 18, 32, 1280, 1280)]:
        x = torch.randn(Nf, H, W, Cin, device=dev).to(torch.bfloat16)
        w = packing.pack_conv3x3(torch.randn(Cout, Cin, 3, 3) * (9 * Cin) ** -0.5, dev)
        out = torch.empty(Nf * H * W, Cout, device=dev, dtype=torch.bfloat16)
        for mode in (2, 3):
            ops.gemm_pair_mode(mode)
            ms = timeit(lambda: ops.conv3x3(x, w, None, out=out))
            print(f"conv {Nf}x{H}x{W} {Cin}->{Cout} mode{mode}: {ms:.3f} ms "
                  f"{2.0 * Nf * H * W * Cin * Cout * 9 / ms / 1e9:.0f} TF/s", flush=True)


if __name__ == "__main__":
    main()
