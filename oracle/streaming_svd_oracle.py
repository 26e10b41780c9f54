"""ORACLE — test infrastructure only.  Never imported by the product path (streamingt2v_b200/).

CPU (PyTorch fp32, functional) restatement of the reference's StreamingSVD denoiser forward:
    StreamingWrapper.forward        /root/reference/code/models/diffusion/wrappers.py:23-78
      ControlNet.forward            code/models/control/controlnet.py:496-554
      VideoUNet.forward             code/models/diffusion/video_model.py:540-618
Each function cites the reference lines it follows.  It consumes a plain state dict with the reference's own SGM
key names and the block plan of streamingt2v_b200/arch.py — no nn.Module, no reference import —
so it runs on the GPU box where /root/reference does not exist.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path, so the
oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, produced in the build container by
oracle/make_golden.py (imports the unmodified reference modules through oracle/ref_shims.py, loads the same
synthetic state dict, and stores input/output vectors under tests/golden/).  tests/test_oracle.py checks this file
against those vectors.  Third-party arithmetic the reference calls but does not vendor (diffusers==0.30.2
`Attention` inside CAM, cam/conditioning.py:4,31-32; xformers) is restated from its published definition
(to_q/to_k/to_v without bias, softmax(QK^T/sqrt(d))V per head, to_out.0 with bias) and remains
"parity unpinned" at that boundary: the golden generator uses the same restatement as a stand-in class.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

from streamingt2v_b200.arch import Attn, Down, Res, Up, UNetConfig, build_plan

SD = Dict[str, torch.Tensor]


# ---- primitives ------------------------------------------------------------------------------------------------
def timestep_embedding(t: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """diffusionmodules/util.py:207-231 (repeat_only=False)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def linear(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def group_norm(sd: SD, p: str, x: torch.Tensor, eps: float) -> torch.Tensor:
    """GroupNorm32 (util.py:274-276, eps 1e-5) / Normalize (attention.py:132-135, eps 1e-6); 32 groups."""
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps)


def layer_norm(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def emb_mlp(sd: SD, p0: str, p2: str, x: torch.Tensor) -> torch.Tensor:
    """Linear -> SiLU -> Linear (time_embed / label_emb / time_pos_embed; video_model.py:195-222,
    video_attention.py:248-252)."""
    return linear(sd, p2, F.silu(linear(sd, p0, x)))


# ---- ResBlock / VideoResBlock ------------------------------------------------------------------------------------
def resblock2d(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """ResBlock._forward, dims=2, no up/down, no scale-shift (openaimodel.py:328-354)."""
    h = F.conv2d(F.silu(group_norm(sd, p + ".in_layers.0", x, 1e-5)), sd[p + ".in_layers.2.weight"],
                 sd[p + ".in_layers.2.bias"], padding=1)
    emb_out = linear(sd, p + ".emb_layers.1", F.silu(emb))
    h = h + emb_out[:, :, None, None]
    h = F.conv2d(F.silu(group_norm(sd, p + ".out_layers.0", h, 1e-5)), sd[p + ".out_layers.3.weight"],
                 sd[p + ".out_layers.3.bias"], padding=1)
    if (p + ".skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + ".skip_connection.weight"], sd[p + ".skip_connection.bias"])
    return x + h


def resblock3d(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor) -> torch.Tensor:
    """time_stack ResBlock: dims=3, kernel (3,1,1), exchange_temb_dims (openaimodel.py:328-354, :350-351).
    x: [b, c, t, h, w]; emb: [b, t, temb]."""
    h = F.conv3d(F.silu(group_norm(sd, p + ".in_layers.0", x, 1e-5)), sd[p + ".in_layers.2.weight"],
                 sd[p + ".in_layers.2.bias"], padding=(1, 0, 0))
    emb_out = linear(sd, p + ".emb_layers.1", F.silu(emb))          # [b, t, c]
    h = h + emb_out.permute(0, 2, 1)[:, :, :, None, None]            # "b t c ... -> b c t ..."
    h = F.conv3d(F.silu(group_norm(sd, p + ".out_layers.0", h, 1e-5)), sd[p + ".out_layers.3.weight"],
                 sd[p + ".out_layers.3.bias"], padding=(1, 0, 0))
    return x + h


def video_resblock(sd: SD, p: str, x: torch.Tensor, emb: torch.Tensor, T: int) -> torch.Tensor:
    """VideoResBlock.forward (video_model.py:66-85); AlphaBlender learned_with_images with indicator == 0
    (util.py:346-369): alpha = sigmoid(mix_factor)."""
    x = resblock2d(sd, p, x, emb)
    bt, c, hh, ww = x.shape
    b = bt // T
    x5 = x.reshape(b, T, c, hh, ww).permute(0, 2, 1, 3, 4)
    xt = resblock3d(sd, p + ".time_stack", x5, emb.reshape(b, T, -1))
    alpha = torch.sigmoid(sd[p + ".time_mixer.mix_factor"])
    out = alpha * x5 + (1.0 - alpha) * xt
    return out.permute(0, 2, 1, 3, 4).reshape(bt, c, hh, ww)


# ---- attention ---------------------------------------------------------------------------------------------------
def mha(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) v per head (attention.py:320-351; SDPA default scale)."""
    b, n, c = q.shape
    d = c // heads
    q, k, v = (t.reshape(t.shape[0], t.shape[1], heads, d).permute(0, 2, 1, 3) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v)
    return o.permute(0, 2, 1, 3).reshape(b, n, c)


def cross_attention(sd: SD, p: str, x: torch.Tensor, context: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """CrossAttention.forward (attention.py:288-351); context=None -> self attention."""
    ctx = x if context is None else context
    o = mha(linear(sd, p + ".to_q", x), linear(sd, p + ".to_k", ctx), linear(sd, p + ".to_v", ctx), heads)
    return linear(sd, p + ".to_out.0", o)


def feed_forward(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward with GEGLU (attention.py:94-120)."""
    a, g = linear(sd, p + ".net.0.proj", x).chunk(2, dim=-1)
    return linear(sd, p + ".net.2", a * F.gelu(g))


def apm_context(sd: SD, p: str, context: torch.Tensor) -> torch.Tensor:
    """BasicTransformerBlockWithAPM.forward (attention.py:612-620): mix the 17 CLIP tokens into one."""
    mixed = F.conv1d(context, sd[p + ".apm_conv.weight"], sd[p + ".apm_conv.bias"], padding=1)
    mixed = layer_norm(sd, p + ".apm_ln", mixed)
    return context[:, :1] + mixed * F.silu(sd[p + ".apm_alpha"])


def basic_transformer_block(sd: SD, p: str, x: torch.Tensor, context: torch.Tensor, heads: int,
                            use_apm: bool) -> torch.Tensor:
    """BasicTransformerBlock._forward (attention.py:567-593)."""
    if use_apm and context.shape[1] > 1:
        context = apm_context(sd, p, context)
    x = cross_attention(sd, p + ".attn1", layer_norm(sd, p + ".norm1", x), None, heads) + x
    x = cross_attention(sd, p + ".attn2", layer_norm(sd, p + ".norm2", x), context, heads) + x
    x = feed_forward(sd, p + ".ff", layer_norm(sd, p + ".norm3", x)) + x
    return x


def video_transformer_block(sd: SD, p: str, x: torch.Tensor, context: torch.Tensor, heads: int,
                            T: int) -> torch.Tensor:
    """VideoTransformerBlock._forward (video_attention.py:125-168), ff_in=True, is_res=True."""
    B, S, Cc = x.shape
    b = B // T
    x = x.reshape(b, T, S, Cc).permute(0, 2, 1, 3).reshape(b * S, T, Cc)     # "(b t) s c -> (b s) t c"
    x = feed_forward(sd, p + ".ff_in", layer_norm(sd, p + ".norm_in", x)) + x
    x = cross_attention(sd, p + ".attn1", layer_norm(sd, p + ".norm1", x), None, heads) + x
    x = cross_attention(sd, p + ".attn2", layer_norm(sd, p + ".norm2", x), context, heads) + x
    x = feed_forward(sd, p + ".ff", layer_norm(sd, p + ".norm3", x)) + x
    return x.reshape(b, S, T, Cc).permute(0, 2, 1, 3).reshape(B, S, Cc)


def spatial_video_transformer(sd: SD, p: str, x: torch.Tensor, context: torch.Tensor, heads: int, T: int,
                              use_apm: bool) -> torch.Tensor:
    """SpatialVideoTransformer.forward (video_attention.py:260-333): use_linear, use_spatial_context,
    depth 1, merge 'learned_with_images' with indicator == 0."""
    bt, c, h, w = x.shape
    x_in = x
    time_context = context[::T]                                             # :281
    time_context = time_context.repeat_interleave(h * w, dim=0)             # "b ... -> (b n) ...", :282-285
    if use_apm and time_context.shape[1] > 1:
        # the temporal block is a plain VideoTransformerBlock: it cross-attends to ALL tokens it is given
        pass
    x = group_norm(sd, p + ".norm", x, 1e-6)
    x = x.permute(0, 2, 3, 1).reshape(bt, h * w, c)
    x = linear(sd, p + ".proj_in", x)
    frames = torch.arange(T, device=x.device).repeat(bt // T)
    emb = emb_mlp(sd, p + ".time_pos_embed.0", p + ".time_pos_embed.2", timestep_embedding(frames, c))[:, None, :]
    x = basic_transformer_block(sd, p + ".transformer_blocks.0", x, context, heads, use_apm)
    x_mix = video_transformer_block(sd, p + ".time_stack.0", x + emb, time_context, heads, T)
    alpha = torch.sigmoid(sd[p + ".time_mixer.mix_factor"])
    x = alpha * x + (1.0 - alpha) * x_mix
    x = linear(sd, p + ".proj_out", x)
    x = x.reshape(bt, h, w, c).permute(0, 3, 1, 2)
    return x + x_in


# ---- CAM ---------------------------------------------------------------------------------------------------------
def cam_merger(sd: SD, p: str, sample: torch.Tensor, cond: torch.Tensor, Fq: int, Fc: int) -> torch.Tensor:
    """ConditionalModel.forward -> CrossAttention.forward (cam/conditioning.py:117-146, :39-81), eval mode
    (dropout inactive).  sample: [(B Fq), C, H, W]; cond (ControlNet features): [(B Fc), C, H, W]."""
    t = p + ".temporal_transformer"
    bf, c, h, w = sample.shape
    B = bf // Fq
    heads = c // 64
    kv = cond.reshape(B, Fc, c, h, w).permute(0, 3, 4, 1, 2).reshape(B * h * w, Fc, c)       # "(B H W) F C"
    x5 = sample.reshape(B, Fq, c, h, w).permute(0, 2, 1, 3, 4)                               # B C F H W
    xn = F.group_norm(x5, 32, sd[t + ".norm.weight"], sd[t + ".norm.bias"], 1e-6)            # stats over (C/32,F,H,W)
    xn = xn.permute(0, 3, 4, 2, 1).reshape(B * h * w, Fq, c)
    xn = linear(sd, t + ".proj_in", xn)
    a = t + ".attention"
    o = mha(linear(sd, a + ".to_q", xn), linear(sd, a + ".to_k", kv), linear(sd, a + ".to_v", kv), heads)
    o = linear(sd, a + ".to_out.0", o)
    res = linear(sd, t + ".proj_out", o)                                                      # (B H W) F C
    res = res.reshape(B, h, w, Fq, c).permute(0, 3, 4, 1, 2).reshape(bf, c, h, w)
    return sample + res


# ---- ControlNet ----------------------------------------------------------------------------------------------------
def cond_embedding(sd: SD, p: str, cond: torch.Tensor, n_pairs: int) -> torch.Tensor:
    """ControlNetConditioningEmbedding.forward with use_normalization (controlnet.py:104-121)."""
    e = F.silu(F.conv2d(cond, sd[p + ".conv_in.weight"], sd[p + ".conv_in.bias"], padding=1))
    for i in range(2 * n_pairs):
        stride = 2 if i % 2 == 1 else 1
        e = F.conv2d(e, sd[f"{p}.blocks.{i}.weight"], sd[f"{p}.blocks.{i}.bias"], padding=1, stride=stride)
        e = layer_norm(sd, f"{p}.norms.{i}", e.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
        e = F.silu(e)
    return F.conv2d(e, sd[p + ".conv_out.weight"], sd[p + ".conv_out.bias"], padding=1)


def _run_block(sd: SD, blk, h, emb, context, T, use_apm, cfg: UNetConfig):
    for layer in blk.layers:
        if isinstance(layer, tuple):
            h = F.conv2d(h, sd[layer[1] + ".weight"], sd[layer[1] + ".bias"], padding=1)
        elif isinstance(layer, Res):
            h = video_resblock(sd, layer.prefix, h, emb, T)
        elif isinstance(layer, Attn):
            h = spatial_video_transformer(sd, layer.prefix, h, context, layer.heads, T, use_apm)
        elif isinstance(layer, Down):
            h = F.conv2d(h, sd[layer.prefix + ".op.weight"], sd[layer.prefix + ".op.bias"], padding=1, stride=2)
        elif isinstance(layer, Up):
            h = F.interpolate(h, scale_factor=2, mode="nearest")
            h = F.conv2d(h, sd[layer.prefix + ".conv.weight"], sd[layer.prefix + ".conv.bias"], padding=1)
    return h


def _embed(sd: SD, root: str, cfg: UNetConfig, timesteps, y):
    t_emb = timestep_embedding(timesteps, cfg.model_channels)
    emb = emb_mlp(sd, root + "time_embed.0", root + "time_embed.2", t_emb)
    return emb + emb_mlp(sd, root + "label_emb.0.0", root + "label_emb.0.2", y)


def controlnet_forward(sd: SD, cfg: UNetConfig, x, timesteps, context, y, controlnet_cond, T: int,
                       root: str = "", taps: Optional[dict] = None):
    """ControlNet.forward (controlnet.py:496-554); Merger 'addition', frame_expansion 'none' (:23-48)."""
    plan = build_plan(cfg, root, decoder=False)
    emb = _embed(sd, root, cfg, timesteps, y)
    ce = cond_embedding(sd, root + "controlnet_cond_embedding", controlnet_cond, len(cfg.cond_embed_channels) - 1)
    if taps is not None:
        taps["ctrl.cond_embedding"] = ce
    hs = []
    h = x
    for i, blk in enumerate(plan.input_blocks):
        h = _run_block(sd, blk, h, emb, context, T, False, cfg)
        if i == 0:
            h = h + ce
        hs.append(h)
        if taps is not None:
            taps[f"ctrl.input_blocks.{i}"] = h
    h = _run_block(sd, plan.middle, h, emb, context, T, False, cfg)
    if taps is not None:
        taps["ctrl.middle"] = h
    return hs, h


def unet_forward(sd: SD, cfg: UNetConfig, x, timesteps, context, y, T: int, Fc: int, hs_ctrl=None, mid_ctrl=None,
                 root: str = "", taps: Optional[dict] = None):
    """VideoUNet.forward (video_model.py:540-618)."""
    plan = build_plan(cfg, root, decoder=True)
    emb = _embed(sd, root, cfg, timesteps, y)
    hs: List[torch.Tensor] = []
    h = x
    for i, blk in enumerate(plan.input_blocks):
        h = _run_block(sd, blk, h, emb, context, T, cfg.use_apm, cfg)
        hs.append(h)
        if taps is not None:
            taps[f"unet.input_blocks.{i}"] = h
    if hs_ctrl is not None:
        hs = [cam_merger(sd, f"{root}cross_attention_merger_input_blocks.{i}", a, c, T, Fc)
              for i, (a, c) in enumerate(zip(hs, hs_ctrl))]
        if taps is not None:
            for i, t_ in enumerate(hs):
                taps[f"unet.merged.{i}"] = t_
    h = _run_block(sd, plan.middle, h, emb, context, T, cfg.use_apm, cfg)
    if mid_ctrl is not None:
        h = cam_merger(sd, f"{root}cross_attention_merger_mid_block", h, mid_ctrl, T, Fc)
    if taps is not None:
        taps["unet.middle"] = h
    for i, blk in enumerate(plan.output_blocks):
        h = torch.cat([h, hs.pop()], dim=1)
        h = _run_block(sd, blk, h, emb, context, T, cfg.use_apm, cfg)
        if taps is not None:
            taps[f"unet.output_blocks.{i}"] = h
    h = F.silu(group_norm(sd, root + "out.0", h, 1e-5))
    return F.conv2d(h, sd[root + "out.2.weight"], sd[root + "out.2.bias"], padding=1)


def streaming_wrapper_forward(sd_unet: SD, sd_ctrl: SD, cfg: UNetConfig, x, t, c: dict, *, batch_size: int,
                              num_video_frames: int, ctrl_frames, image_only_indicator=None,
                              num_conditional_frames=None, taps: Optional[dict] = None):
    """StreamingWrapper.forward (wrappers.py:23-78).  image_only_indicator must be all zeros (streaming_svd.py:208);
    num_conditional_frames is accepted and ignored like the reference does."""
    Fc = cfg.num_frame_conditioning
    B, T = batch_size, num_video_frames

    def reduce(inp):                                                         # wrappers.py:28-31
        return inp.reshape(B, T, *inp.shape[1:])[:, :Fc].reshape(B * Fc, *inp.shape[1:])

    x = torch.cat((x, c["concat"]), dim=1)                                   # :33
    context = c["crossattn"]
    y = c["vector"]
    cc = ctrl_frames.repeat(2, *([1] * (ctrl_frames.dim() - 1)))           # "B ... -> (2 B) ..." :45-46
    controlnet_cond = cc.reshape(-1, *cc.shape[2:])                          # "B F ... -> (B F) ..." :47-48
    hs_c, mid_c = controlnet_forward(sd_ctrl, cfg, reduce(x), reduce(t), reduce(context[:, :1]), reduce(y),
                                     controlnet_cond, Fc, taps=taps)
    return unet_forward(sd_unet, cfg, x, t, context, y, T, Fc, hs_c, mid_c, taps=taps)
