import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from streamingt2v_b200 import _lib, ops
_lib.init(0)
dev = torch.device("cuda:0")
n, s, h = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
qkv = torch.randn(n * s, 3 * h * 64, device=dev).to(torch.bfloat16)
for _ in range(3):
    ops.flash_attn(qkv, n, s, h)
torch.cuda.synchronize()
