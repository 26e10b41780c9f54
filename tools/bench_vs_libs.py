"""End-of-round table: b200svd_gemm with its DEFAULT tile selection against the NVIDIA libraries on the denoiser's twelve
dominant layer shapes (CUDA events, L2 flushed between iterations, median of 9).

Two library columns per shape:
  lib_core : the bare contraction in cuBLAS (torch.matmul) / cuDNN (F.conv2d, channels_last), bf16 — what SURVEY.md
             section 2.2 calls "the kernels to beat";
  lib_op   : the same layer as the reference executes it in eager PyTorch — contraction + bias + activation /
             residual as separate library kernels (F.linear + F.gelu * ..., conv2d + add).
`ours` is one launch with everything fused.  ratio_* > 1 means ours is faster.
    python tools/bench_vs_libs.py  ->  gpurun_out/bench_vs_libs.json + a text table on stdout"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from streamingt2v_b200 import _lib, ops, packing
from tools.bench_gemm import timeit


def main():
    _lib.init(0)
    dev = torch.device("cuda:0")
    rows = []
    g = torch.Generator(device=dev).manual_seed(0)

    def rnd(*shape, scale=1.0):
        return (torch.randn(*shape, device=dev, generator=g) * scale).to(torch.bfloat16)

    # ---- linear layers: (name, M, K, N, kind) ----
    lin = [("L0 qkv", 460800, 320, 960, "plain"), ("L0 ff1 geglu", 460800, 320, 2560, "geglu"),
           ("L0 ff2 +res", 460800, 1280, 320, "res"), ("L0 proj +res", 460800, 320, 320, "res"),
           ("L1 ff1 geglu", 115200, 640, 5120, "geglu"), ("L1 ff2 +res", 115200, 2560, 640, "res"),
           ("L2 ff1 geglu", 28800, 1280, 10240, "geglu"), ("L2 ff2 +res", 28800, 5120, 1280, "res")]
    for name, M, K, N, kind in lin:
        x = rnd(M, K)
        wt = torch.randn(N, K, device=dev, generator=g) * K ** -0.5
        b = torch.randn(N, device=dev, generator=g) * 0.1
        wb = wt.to(torch.bfloat16)
        bb = b.to(torch.bfloat16)
        flops = 2.0 * M * K * N
        if kind == "geglu":
            wp, bp, bn = packing.pack_geglu(wt.cpu(), b.cpu(), dev)
            out = torch.empty(M, N // 2, device=dev, dtype=torch.bfloat16)
            ours = timeit(lambda: ops.linear(x, wp, bp, act=ops.ACT_GEGLU, bn=bn, out=out), iters=9)

            def lib_op():
                a, gt = F.linear(x, wb, bb).chunk(2, dim=-1)
                return a * F.gelu(gt)
        elif kind == "res":
            wp = packing.pack_linear(wt.cpu(), dev)
            bp = packing.f32(b.cpu(), dev)
            res = rnd(M, N)
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            ours = timeit(lambda: ops.linear(x, wp, bp, res1=res, s1=1.0, out=out), iters=9)

            def lib_op():
                return F.linear(x, wb, bb) + res
        else:
            wp = packing.pack_linear(wt.cpu(), dev)
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            ours = timeit(lambda: ops.linear(x, wp, None, out=out), iters=9)

            def lib_op():
                return F.linear(x, wb)
        core = timeit(lambda: torch.matmul(x, wb.t()), iters=9)
        op = timeit(lib_op, iters=9)
        rows.append(dict(name=name, shape=f"M{M} K{K} N{N}", ours_ms=ours, ours_tflops=flops / ours / 1e9,
                         lib_core_ms=core, lib_core_tflops=flops / core / 1e9, lib_op_ms=op,
                         ratio_core=core / ours, ratio_op=op / ours))
        print(rows[-1], flush=True)
        del x, wt, wb, out

    # ---- 3x3 convolutions: (name, frames, H, W, Cin, Cout, kind) ----
    convs = [("conv L0 320->320 +res", 50, 72, 128, 320, 320, "res"), ("conv L0 960->320 +emb", 50, 72, 128, 960, 320, "emb"),
             ("conv L1 640->640 +res", 50, 36, 64, 640, 640, "res"), ("conv L2 1280->1280 +res", 50, 18, 32, 1280, 1280, "res")]
    for name, Nf, H, W, Cin, Cout, kind in convs:
        x = rnd(Nf, H, W, Cin)
        wt = torch.randn(Cout, Cin, 3, 3, device=dev, generator=g) * (9 * Cin) ** -0.5
        b = torch.randn(Cout, device=dev, generator=g) * 0.1
        wp = packing.pack_conv3x3(wt.cpu(), dev)
        bp = packing.f32(b.cpu(), dev)
        out = torch.empty(Nf * H * W, Cout, device=dev, dtype=torch.bfloat16)
        flops = 2.0 * Nf * H * W * Cin * Cout * 9
        xn = x.permute(0, 3, 1, 2)                                     # NCHW view of channels_last memory
        wn = wt.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
        bn_ = b.to(torch.bfloat16)
        if kind == "res":
            res = rnd(Nf * H * W, Cout)
            resn = res.view(Nf, H, W, Cout).permute(0, 3, 1, 2)
            ours = timeit(lambda: ops.conv3x3(x, wp, bp, res1=res, s1=1.0, out=out), iters=9)

            def lib_op():
                return F.conv2d(xn, wn, bn_, padding=1) + resn
        else:
            emb = torch.randn(Nf, Cout, device=dev, generator=g)
            embn = emb.to(torch.bfloat16)[:, :, None, None]
            ours = timeit(lambda: ops.conv3x3(x, wp, bp, fvec=emb, rows_per_frame=H * W, out=out), iters=9)

            def lib_op():
                return F.conv2d(xn, wn, bn_, padding=1) + embn
        core = timeit(lambda: F.conv2d(xn, wn, padding=1), iters=9)
        op = timeit(lib_op, iters=9)
        rows.append(dict(name=name, shape=f"{Nf}x{H}x{W} {Cin}->{Cout}", ours_ms=ours, ours_tflops=flops / ours / 1e9,
                         lib_core_ms=core, lib_core_tflops=flops / core / 1e9, lib_op_ms=op,
                         ratio_core=core / ours, ratio_op=op / ours))
        print(rows[-1], flush=True)
        del x, wt, out

    print(f"\n{'layer':26s} {'shape':24s} {'ours ms':>8s} {'TF/s':>7s} | {'lib core':>8s} {'TF/s':>7s} {'x':>5s} | "
          f"{'lib op':>8s} {'x':>5s}")
    for r in rows:
        print(f"{r['name']:26s} {r['shape']:24s} {r['ours_ms']:8.3f} {r['ours_tflops']:7.0f} | {r['lib_core_ms']:8.3f} "
              f"{r['lib_core_tflops']:7.0f} {r['ratio_core']:5.2f} | {r['lib_op_ms']:8.3f} {r['ratio_op']:5.2f}")
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/bench_vs_libs.json", "w"), indent=1)


if __name__ == "__main__":
    main()
