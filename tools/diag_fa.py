"""Diagnostic: which (frame, head, 256-row query slab) of a flash_attn launch disagree with SDPA."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from streamingt2v_b200 import _lib, ops

_lib.init(0)
dev = torch.device("cuda:0")
for n, s, heads in [(2, 2304, 10), (1, 9216, 5), (2, 2304, 3)]:
    Cc = heads * 64
    g = torch.Generator().manual_seed(n * 1000 + s)
    qkv = (torch.randn((n * s, 3 * Cc), generator=g) * 1.5).to(dev).to(torch.bfloat16)
    outs = []
    for rep in range(3):
        outs.append(ops.flash_attn(qkv, n, s, heads).float())
    torch.cuda.synchronize()
    q, k, v = (t.float().reshape(n, s, heads, 64).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=1))
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(n * s, Cc)
    for rep, out in enumerate(outs):
        err = (out - ref).abs().reshape(n, s, heads, 64).amax(-1)          # [n, s, heads]
        bad = err > 0.05
        slabs = (s + 255) // 256
        msg = []
        for f in range(n):
            for h in range(heads):
                for sl in range(slabs):
                    b = bad[f, sl * 256:(sl + 1) * 256, h]
                    if b.any():
                        rows = b.nonzero().flatten()
                        msg.append(f"(f{f} h{h} slab{sl} cta#{(f * heads + h) * slabs + sl}: {int(b.sum())} rows, first {int(rows[0])} last {int(rows[-1])})")
        print(f"V={os.environ.get('B200SVD_FA_V', '4')} n{n} s{s} h{heads} rep{rep}: max_err {float(err.max()):.3e} bad slabs {len(msg)}: {' '.join(msg[:12])}",
              flush=True)
    print("   reps identical:", torch.equal(outs[0], outs[1]), torch.equal(outs[1], outs[2]))
