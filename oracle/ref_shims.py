"""ORACLE tooling — build-container only (needs /root/reference).  Import shims that let the UNMODIFIED reference
modules (VideoUNet, ControlNet, StreamingWrapper) import and run on CPU here.

The reference imports pytorch_lightning / omegaconf / kornia / open_clip / diffusers at module import time
(code/models/svd/sgm/__init__.py:1, cam/conditioning.py:4), none of which is installed offline.  Only one of the
shimmed symbols does arithmetic on this path: `diffusers.models.attention_processor.Attention`, used by CAM
(cam/conditioning.py:31-32).  The stand-in below restates diffusers==0.30.2 `Attention` + `AttnProcessor2_0`
from its published definition (to_q/to_k/to_v Linear without bias, per-head softmax(QK^T/sqrt(d))V through SDPA,
to_out = [Linear(inner, query_dim, bias=True), Dropout]) with the same state-dict key names as the StreamingSVD
checkpoint (`attention.to_q.weight`, `attention.to_out.0.{weight,bias}`) — "parity unpinned" at that boundary.
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REFERENCE_CODE = "/root/reference/code"


class _Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, dim_head=64, bias=False,
                 upcast_attention=False, **_):
        super().__init__()
        inner = heads * dim_head
        cross = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner, bias=bias)
        self.to_k = nn.Linear(cross, inner, bias=bias)
        self.to_v = nn.Linear(cross, inner, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **_):
        ctx = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        q, k, v = self.to_q(hidden_states), self.to_k(ctx), self.to_v(ctx)
        b, n, c = q.shape
        d = c // self.heads
        q, k, v = (t.view(t.shape[0], t.shape[1], self.heads, d).transpose(1, 2) for t in (q, k, v))
        o = F.scaled_dot_product_attention(q, k, v, attn_mask=attention_mask)
        o = o.transpose(1, 2).reshape(b, n, c)
        return self.to_out[1](self.to_out[0](o))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install():
    """Register the shim modules and put the reference on sys.path.  Idempotent."""
    if REFERENCE_CODE not in sys.path:
        sys.path.insert(0, REFERENCE_CODE)
    if "pytorch_lightning" not in sys.modules:
        _mod("pytorch_lightning", LightningModule=nn.Module, LightningDataModule=object)
    if "omegaconf" not in sys.modules:
        _mod("omegaconf", ListConfig=list, OmegaConf=dict, DictConfig=dict)
    for name in ("kornia", "open_clip"):
        if name not in sys.modules:
            _mod(name)
    if "diffusers" not in sys.modules:
        d = _mod("diffusers")
        dm = _mod("diffusers.models")
        dap = _mod("diffusers.models.attention_processor", Attention=_Attention)
        d.models = dm
        dm.attention_processor = dap


def build_reference(cfg):
    """Instantiate the reference's own StreamingWrapper(VideoUNet, ControlNet) for an arch.UNetConfig, with the
    literal init_args of code/config.yaml:69-115 / :47-59 except attn type 'softmax' (xformers is CUDA-only)."""
    install()
    from models.control.controlnet import ControlNet
    from models.diffusion.video_model import VideoUNet
    from models.diffusion.wrappers import StreamingWrapper
    from models.svd.sgm.modules.diffusionmodules.wrappers import OpenAIWrapper
    unet = VideoUNet(
        in_channels=cfg.in_channels, model_channels=cfg.model_channels, out_channels=cfg.out_channels,
        num_res_blocks=cfg.num_res_blocks, num_conditional_frames=None,
        attention_resolutions=list(cfg.attention_resolutions), dropout=0.0, channel_mult=list(cfg.channel_mult),
        conv_resample=True, dims=2, num_classes="sequential", use_checkpoint=False, num_heads=-1,
        num_head_channels=cfg.num_head_channels, num_heads_upsample=-1, use_scale_shift_norm=False,
        resblock_updown=False, transformer_depth=1, transformer_depth_middle=None, context_dim=cfg.context_dim,
        time_downup=False, time_context_dim=None, extra_ff_mix_layer=True, use_spatial_context=True,
        merge_strategy="learned_with_images", merge_factor=0.5, spatial_transformer_attn_type="softmax",
        video_kernel_size=[3, 1, 1], use_linear_in_transformer=True, adm_in_channels=cfg.adm_in_channels,
        disable_temporal_crossattention=False, max_ddpm_temb_period=10000,
        merging_mode="attention_cross_attention", controlnet_mode=True, use_apm=cfg.use_apm)
    ctrl = ControlNet.from_unet(
        OpenAIWrapper(unet), merging_mode="addition", zero_conv_mode="Identity", frame_expansion="none",
        downsample_controlnet_cond=True, use_image_encoder_normalization=True, use_controlnet_mask=False,
        condition_encoder="", conditioning_embedding_out_channels=list(cfg.cond_embed_channels))
    wrapper = StreamingWrapper(unet, ctrl, num_frame_conditioning=cfg.num_frame_conditioning)
    return wrapper.eval()
