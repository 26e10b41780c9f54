"""Micro-benchmark of b200svd_gemm on the denoiser's dominant shapes (CUDA events, L2 flushed between iterations)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamingt2v_b200 import _lib, ops, packing


def timeit(fn, iters=10, warmup=3):
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    _lib.init(0)
    dev = torch.device("cuda:0")
    res = []
    # (name, M, K, N)
    for name, M, K, N, bn in [("L0 qkv", 460800, 320, 960, 160), ("L0 ff2", 460800, 1280, 320, 160),
                              ("L1 ff2", 115200, 2560, 640, 160), ("L2 ff2", 28800, 5120, 1280, 160),
                              ("L2 sq", 28800, 1280, 1280, 160), ("L2 sq bn128", 28800, 1280, 1280, 128),
                              ("8k cube", 8192, 8192, 8192, 128), ("8k cube 160", 8192, 8192, 8000, 160)]:
        x = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = packing.pack_linear(torch.randn(N, K) * K ** -0.5, dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.linear(x, w, None, out=out, bn=bn))
        tf = 2.0 * M * K * N / ms / 1e9
        ref_ms = timeit(lambda: torch.matmul(x, w[0].t()))
        res.append(dict(name=name, M=M, K=K, N=N, bn=bn, ms=ms, tflops=tf, cublas_ms=ref_ms,
                        cublas_tflops=2.0 * M * K * N / ref_ms / 1e9))
        print(res[-1], flush=True)
    # conv 3x3 at L0 / L1
    for name, Nf, H, W, Cin, Cout in [("conv L0 320", 50, 72, 128, 320, 320), ("conv L1 640", 50, 36, 64, 640, 640),
                                      ("conv L2 1280", 50, 18, 32, 1280, 1280), ("conv L0 960->320", 50, 72, 128, 960, 320)]:
        x = torch.randn(Nf, H, W, Cin, device=dev).to(torch.bfloat16)
        wt = torch.randn(Cout, Cin, 3, 3) * (9 * Cin) ** -0.5
        w = packing.pack_conv3x3(wt, dev)
        out = torch.empty(Nf * H * W, Cout, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.conv3x3(x, w, None, out=out))
        fl = 2.0 * Nf * H * W * Cin * Cout * 9
        xn = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
        wn = wt.to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        ref_ms = timeit(lambda: torch.nn.functional.conv2d(xn, wn, padding=1))
        res.append(dict(name=name, ms=ms, tflops=fl / ms / 1e9, cudnn_ms=ref_ms, cudnn_tflops=fl / ref_ms / 1e9))
        print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    tag = os.environ.get("B200SVD_BENCH_TAG", "")
    json.dump(res, open(f"gpurun_out/bench_gemm{tag}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
