"""ORACLE — test infrastructure only.  CPU (PyTorch fp32, functional) restatement of the SD-VAE encoder of the conditioner:
    AutoencoderKLModeOnly.encode(x) = mode(quant_conv(Encoder(x)))
    /root/reference/code/models/svd/sgm/modules/diffusionmodules/model.py:52-55 (Normalize), :73-92 (Downsample: pad
    (0,1,0,1) + stride-2 conv), :95-159 (ResnetBlock, temb None), :161-201 (AttnBlock), :487-601 (Encoder)
    /root/reference/code/models/svd/sgm/models/autoencoder.py:454-473 (quant_conv, encode), :602-615 (regulariser with
    sample=False: the posterior MODE = the mean = first z_channels channels)
Pinned against the unmodified reference `Encoder` by oracle/make_golden_vae_enc.py (tests/golden/vae_enc_*.npz)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

from streamingt2v_b200.arch import VaeConfig, vae_encoder_plan

SD = Dict[str, torch.Tensor]


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)


def _swish(x):
    return x * torch.sigmoid(x)


def resnet_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """ResnetBlock.forward with temb = None (model.py:139-159)."""
    h = F.conv2d(_swish(_gn(sd, p + ".norm1", x)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(_swish(_gn(sd, p + ".norm2", h)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".nin_shortcut.weight"], sd[p + ".nin_shortcut.bias"])
    return x + h


def attn_block(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """AttnBlock.forward (model.py:180-201): one head of width C over the h*w positions."""
    h = _gn(sd, p + ".norm", x)
    q, k, v = (F.conv2d(h, sd[f"{p}.{n}.weight"], sd[f"{p}.{n}.bias"]) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q, k, v = (t.reshape(b, 1, c, hh * ww).permute(0, 1, 3, 2) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v).permute(0, 1, 3, 2).reshape(b, c, hh, ww)
    return x + F.conv2d(o, sd[p + ".proj_out.weight"], sd[p + ".proj_out.bias"])


def encode(sd: SD, cfg: VaeConfig, x: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
    """x [n, 3, H, W] in [-1, 1] -> posterior mode [n, z_channels, H/8, W/8] (no scale factor applied)."""
    h = x
    for kind, p, cin, cout in vae_encoder_plan(cfg):
        if kind == "conv_in":
            h = F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1)
        elif kind == "res":
            h = resnet_block(sd, p, h)
        elif kind == "down":
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[p + ".conv.weight"], sd[p + ".conv.bias"], stride=2)
        elif kind == "attn":
            h = attn_block(sd, p, h)
        elif kind == "out":
            h = F.conv2d(_swish(_gn(sd, "norm_out", h)), sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
        if taps is not None and kind != "out":
            taps[p] = h
    moments = F.conv2d(h, sd["quant_conv.weight"], sd["quant_conv.bias"])
    return moments[:, :cfg.z_channels]
