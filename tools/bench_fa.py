"""FlashAttention timing at the denoiser's three spatial-attention shapes (CUDA events, median of 7)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamingt2v_b200 import _lib, ops
from tools.bench_gemm import timeit

_lib.init(0)
dev = torch.device("cuda:0")
for n, s, h in [(50, 9216, 5), (50, 2304, 10), (50, 576, 20)]:
    qkv = torch.randn(n * s, 3 * h * 64, device=dev).to(torch.bfloat16)
    out = torch.empty(n * s, h * 64, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.flash_attn(qkv, n, s, h, out=out), iters=7)
    fl = 4.0 * n * h * s * s * 64
    q, k, v = (t.float().reshape(n, s, h, 64).permute(0, 2, 1, 3)[:1, :2, :256] for t in qkv.chunk(3, dim=1))
    kf, vf = (t.float().reshape(n, s, h, 64).permute(0, 2, 1, 3)[:1, :2] for t in qkv.chunk(3, dim=1)[1:])
    ref = torch.softmax(q @ kf.transpose(-1, -2) * 0.125, -1) @ vf
    got = out.float().reshape(n, s, h, 64).permute(0, 2, 1, 3)[:1, :2, :256]
    err = (got - ref).abs().max().item()
    print(f"V={os.environ.get('B200SVD_FA_V', '4')} POLY={os.environ.get('B200SVD_FA_POLY', 'default')} n{n} s{s} h{h}: {ms:.3f} ms {fl / ms / 1e9:.0f} TF/s  max_abs_err {err:.2e}", flush=True)
