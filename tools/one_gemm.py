"""Launch a single GEMM shape a few times (for ncu captures)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamingt2v_b200 import _lib, ops, packing

_lib.init(0)
dev = torch.device("cuda:0")
kind = sys.argv[1]
if kind == "lin":
    M, K, N = 460800, int(sys.argv[2]), int(sys.argv[3])
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = packing.pack_linear(torch.randn(N, K) * K ** -0.5, dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    res = torch.randn(M, N, device=dev).to(torch.bfloat16) if len(sys.argv) > 4 else None
    for _ in range(4):
        ops.linear(x, w, None, out=out, res1=res)
elif kind == "geglu":
    M, K, N2 = 460800, int(sys.argv[2]), int(sys.argv[3])
    x = torch.randn(M, K, device=dev).to(torch.bfloat16)
    wp, bp, bn = packing.pack_geglu(torch.randn(N2, K) * K ** -0.5, torch.randn(N2), dev)
    out = torch.empty(M, N2 // 2, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.linear(x, wp, bp, act=ops.ACT_GEGLU, out=out, bn=bn)
else:
    x = torch.randn(50, 72, 128, 320, device=dev).to(torch.bfloat16)
    w = packing.pack_conv3x3(torch.randn(320, 320, 3, 3) * 0.02, dev)
    out = torch.empty(50 * 72 * 128, 320, device=dev, dtype=torch.bfloat16)
    for _ in range(4):
        ops.conv3x3(x, w, None, out=out)
torch.cuda.synchronize()
