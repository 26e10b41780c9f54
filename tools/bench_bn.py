"""N-tile choice study: the denoiser's N=640/960/1920 layers with the exact 160-wide tile vs the ragged 256-wide tile."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamingt2v_b200 import _lib, ops, packing
from tools.bench_gemm import timeit


def main():
    _lib.init(0)
    dev = torch.device("cuda:0")
    for M, K, N in [(115200, 640, 640), (115200, 2560, 640), (115200, 640, 1920), (460800, 320, 960),
                    (460800, 320, 320), (460800, 1280, 320), (28800, 1280, 1280)]:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = packing.pack_linear(torch.randn(N, K) * K ** -0.5, dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for bn in (128, 160, 256):
            if N % 160 and bn == 160:
                continue
            ms = timeit(lambda: ops.linear(x, w, None, out=out, bn=bn))
            print(f"lin M{M} K{K} N{N} bn{bn}: {ms:.3f} ms {2.0 * M * K * N / ms / 1e9:.0f} TF/s", flush=True)
    for Nf, H, W, Cin, Cout in [(50, 36, 64, 640, 640), (50, 36, 64, 1280, 640), (50, 72, 128, 320, 320)]:
        x = torch.randn(Nf, H, W, Cin, device=dev).to(torch.bfloat16)
        w = packing.pack_conv3x3(torch.randn(Cout, Cin, 3, 3) * (9 * Cin) ** -0.5, dev)
        out = torch.empty(Nf * H * W, Cout, device=dev, dtype=torch.bfloat16)
        for bn in (160, 256):
            ms = timeit(lambda: ops.conv3x3(x, w, None, out=out, bn=bn))
            print(f"conv {Nf}x{H}x{W} {Cin}->{Cout} bn{bn}: {ms:.3f} ms {2.0 * Nf * H * W * Cin * Cout * 9 / ms / 1e9:.0f} TF/s",
                  flush=True)


if __name__ == "__main__":
    main()
