"""ORACLE — test infrastructure only.  CPU (PyTorch fp32, functional) restatement of the temporal VAE decoder:
    AutoencodingEngine.decode -> VideoDecoder.forward(z, timesteps=n)
    /root/reference/code/models/svd/sgm/models/autoencoder.py:210-212
    /root/reference/code/models/svd/sgm/modules/autoencoding/temporal_ae.py:16-105,291-347
    /root/reference/code/models/svd/sgm/modules/diffusionmodules/model.py:52-201,604-748
Pinned against outputs of the unmodified reference VideoDecoder by oracle/make_golden_vae.py (tests/golden/vae_*.npz).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from streamingt2v_b200.arch import VaeConfig, vae_decoder_plan

SD = Dict[str, torch.Tensor]


def _gn(sd, p, x, eps):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def _swish(x):
    return x * torch.sigmoid(x)


def video_resblock(sd: SD, p: str, x: torch.Tensor, T: int) -> torch.Tensor:
    """VideoResBlock.forward (temporal_ae.py:62-81) = ResnetBlock.forward (model.py:139-160, temb None, GN eps 1e-6)
    + time_stack ResBlock(dims=3, skip_t_emb) (openaimodel.py:328-354, GroupNorm32 eps 1e-5) + learned blend
    x = sigmoid(mix)*x_temporal + (1-sigmoid(mix))*x_spatial."""
    h = F.conv2d(_swish(_gn(sd, p + ".norm1", x, 1e-6)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(_swish(_gn(sd, p + ".norm2", h, 1e-6)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    x = x + h
    bt, c, hh, ww = x.shape
    x5 = x.reshape(bt // T, T, c, hh, ww).permute(0, 2, 1, 3, 4)
    t = p + ".time_stack"
    g = F.conv3d(F.silu(_gn(sd, t + ".in_layers.0", x5, 1e-5)), sd[t + ".in_layers.2.weight"],
                 sd[t + ".in_layers.2.bias"], padding=(1, 0, 0))
    g = F.conv3d(F.silu(_gn(sd, t + ".out_layers.0", g, 1e-5)), sd[t + ".out_layers.3.weight"],
                 sd[t + ".out_layers.3.bias"], padding=(1, 0, 0))
    xt = x5 + g
    alpha = torch.sigmoid(sd[p + ".mix_factor"])
    out = alpha * xt + (1.0 - alpha) * x5
    return out.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)


def attn_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """AttnBlock.forward (model.py:161-201): single head, head dim = channels."""
    h = _gn(sd, p + ".norm", x, 1e-6)
    q = F.conv2d(h, sd[p + ".q.weight"], sd[p + ".q.bias"])
    k = F.conv2d(h, sd[p + ".k.weight"], sd[p + ".k.bias"])
    v = F.conv2d(h, sd[p + ".v.weight"], sd[p + ".v.bias"])
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, c, hh * ww).permute(0, 2, 1)[:, None] for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v)[:, 0].permute(0, 2, 1).reshape(b, c, hh, ww)
    return x + F.conv2d(o, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


def decode(sd: SD, cfg: VaeConfig, z: torch.Tensor, timesteps: int, taps: Optional[dict] = None) -> torch.Tensor:
    """VideoDecoder.forward(z, timesteps) (model.py:715-748 with the video blocks of temporal_ae.py)."""
    h = z
    for kind, p, cin, cout in vae_decoder_plan(cfg):
        if kind == "conv_in":
            h = F.conv2d(h, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
        elif kind == "res":
            h = video_resblock(sd, p, h, timesteps)
        elif kind == "attn":
            h = attn_block(sd, p, h)
        elif kind == "up":
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = F.conv2d(h, sd[p + ".conv.weight"], sd[p + ".conv.bias"], padding=1)
        elif kind == "out":
            h = _swish(_gn(sd, "norm_out", h, 1e-6))
            h = F.conv2d(h, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)          # AE3DConv (temporal_ae.py:84-105)
            bt, c, hh, ww = h.shape
            h5 = h.reshape(bt // timesteps, timesteps, c, hh, ww).permute(0, 2, 1, 3, 4)
            h5 = F.conv3d(h5, sd["conv_out.time_mix_conv.weight"], sd["conv_out.time_mix_conv.bias"], padding=(1, 0, 0))
            h = h5.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)
        if taps is not None and kind != "out":
            taps[p] = h
    return h
