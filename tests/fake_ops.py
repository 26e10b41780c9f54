"""TEST INFRASTRUCTURE — a CPU stand-in for streamingt2v_b200.ops with the same call signatures, used only by
`-m "not gpu"` tests to check the HOST logic of the executor (weight packing, epilogue composition, residual and
concat plumbing, hoisting) against the oracle without a GPU.  Every function computes in fp32 with torch CPU ops
and rounds to bf16 where the CUDA kernel would.  The product never imports this module."""
from __future__ import annotations

import math
import types

import torch
import torch.nn.functional as F

ACT_NONE, ACT_SILU, ACT_GELU, ACT_GEGLU = 0, 1, 2, 3
_lib = types.SimpleNamespace(init=lambda *_a, **_k: None)
_launch_count = 0


def launches():
    return _launch_count


def _count():
    global _launch_count
    _launch_count += 1


def _epilogue(v, rows, *, bias, act, fvec, rows_per_frame, s_acc, res1, s1, res2, s2, out, out_fp32, bn):
    _count()
    if bias is not None:
        v = v + bias.float()
    if act == ACT_GEGLU:
        n2 = v.shape[1]
        h = bn // 2
        t = v.reshape(rows, n2 // bn, 2, h)
        if fvec is not None:
            raise AssertionError("fvec with GEGLU unsupported")
        v = (t[:, :, 0] * F.gelu(t[:, :, 1])).reshape(rows, n2 // 2)
    else:
        if fvec is not None:
            idx = torch.arange(rows) // rows_per_frame
            v = v + fvec.float()[idx]
        if act == ACT_SILU:
            v = F.silu(v)
        elif act == ACT_GELU:
            v = F.gelu(v)
    v = s_acc * v
    if res1 is not None:
        v = v + s1 * res1.float()
    if res2 is not None:
        v = v + s2 * res2.float()
    if out is None:
        return v if out_fp32 else v.to(torch.bfloat16)
    out.copy_(v if out.dtype == torch.float32 else v.to(torch.bfloat16))
    return out


def linear(x, w, bias=None, *, act=ACT_NONE, out=None, out_fp32=False, fvec=None, rows_per_frame=1, s_acc=1.0,
           res1=None, s1=1.0, res2=None, s2=1.0, bn=0, gn_rows=None):
    assert x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and w.dim() == 3 and w.shape[0] == 1
    v = x.float() @ w[0].float().t()
    return _epilogue(v, x.shape[0], bias=bias, act=act, fvec=fvec, rows_per_frame=rows_per_frame, s_acc=s_acc,
                     res1=res1, s1=s1, res2=res2, s2=s2, out=out, out_fp32=out_fp32, bn=bn)


def _conv(x, w, stride, bias, out, epi):
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous()
    N, H, W, Cc = x.shape
    cout = w.shape[1]
    wt = w.float().reshape(3, 3, cout, Cc).permute(2, 3, 0, 1)
    xin = x.float().permute(0, 3, 1, 2)
    if epi.pop("pad_after_only", False):
        v = F.conv2d(F.pad(xin, (0, 1, 0, 1)), wt, None, stride=stride).permute(0, 2, 3, 1)
    else:
        v = F.conv2d(xin, wt, None, stride=stride, padding=1).permute(0, 2, 3, 1)
    v = v.reshape(-1, cout)
    return _epilogue(v, v.shape[0], bias=bias, act=epi.get("act", ACT_NONE), fvec=epi.get("fvec"),
                     rows_per_frame=epi.get("rows_per_frame", 1), s_acc=epi.get("s_acc", 1.0), res1=epi.get("res1"),
                     s1=epi.get("s1", 1.0), res2=epi.get("res2"), s2=epi.get("s2", 1.0), out=out,
                     out_fp32=epi.get("out_fp32", False), bn=epi.get("bn", 0))


def conv3x3(x, w, bias=None, *, out=None, **epi):
    return _conv(x, w, 1, bias, out, epi)


def conv3x3_s2(x, w, bias=None, *, out=None, **epi):
    return _conv(x, w, 2, bias, out, epi)


def tconv3(x, w, bias=None, *, out=None, **epi):
    assert x.dtype == torch.bfloat16 and x.dim() == 4 and x.is_contiguous()
    B, T, P, Cc = x.shape
    cout = w.shape[1]
    wt = w.float().permute(1, 2, 0)[..., None, None]                      # [cout, cin, 3, 1, 1]
    v = F.conv3d(x.float().permute(0, 3, 1, 2)[..., None], wt, None, padding=(1, 0, 0))[..., 0]
    v = v.permute(0, 2, 3, 1).reshape(-1, cout)
    return _epilogue(v, v.shape[0], bias=bias, act=epi.get("act", ACT_NONE), fvec=epi.get("fvec"),
                     rows_per_frame=epi.get("rows_per_frame", 1), s_acc=epi.get("s_acc", 1.0), res1=epi.get("res1"),
                     s1=epi.get("s1", 1.0), res2=epi.get("res2"), s2=epi.get("s2", 1.0), out=out,
                     out_fp32=epi.get("out_fp32", False), bn=epi.get("bn", 0))


def flash_attn(qkv, n, s, heads, out=None):
    _count()
    Cc = heads * 64
    q, k, v = (t.float().reshape(n, s, heads, 64).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=1))
    o = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(n * s, Cc).to(torch.bfloat16)
    if out is not None:
        out.copy_(o)
        return out
    return o


def small_attn(q, k, v, *, b, s, heads, lq, lk, kv_per_pixel=True, out=None):
    _count()
    Cc = heads * 64
    qf = q.float().reshape(b, lq, s, heads, 64).permute(0, 2, 3, 1, 4)
    if kv_per_pixel:
        kf = k.float().reshape(b, lk, s, heads, 64).permute(0, 2, 3, 1, 4)
        vf = v.float().reshape(b, lk, s, heads, 64).permute(0, 2, 3, 1, 4)
    else:
        kf = k.float().reshape(b, lk, 1, heads, 64).permute(0, 2, 3, 1, 4).expand(b, s, heads, lk, 64)
        vf = v.float().reshape(b, lk, 1, heads, 64).permute(0, 2, 3, 1, 4).expand(b, s, heads, lk, 64)
    o = F.scaled_dot_product_attention(qf, kf, vf).permute(0, 3, 1, 2, 4).reshape(b * lq * s, Cc).to(torch.bfloat16)
    if out is not None:
        out.copy_(o)
        return out
    return o


def group_norm(x, n, p, gamma, beta, eps, *, silu=False, out=None, sums=None):
    _count()
    _count()
    c = x.shape[1]
    y = F.group_norm(x.float().reshape(n, p, c).permute(0, 2, 1), 32, gamma, beta, eps)
    if silu:
        y = F.silu(y)
    y = y.permute(0, 2, 1).reshape(n * p, c).to(torch.bfloat16).contiguous()
    if out is not None:
        out.copy_(y)
        return out
    return y


def layer_norm(x, gamma, beta, eps=1e-5, *, fvec=None, rows_per_frame=1, xsum=None, silu=False, out=None):
    _count()
    v = x.float()
    if fvec is not None:
        v = v + fvec.float()[torch.arange(x.shape[0]) // rows_per_frame]
        if xsum is not None:
            xsum.copy_(v.to(torch.bfloat16))
            v = xsum.float()
    y = F.layer_norm(v, (x.shape[1],), gamma, beta, eps)
    if silu:
        y = F.silu(y)
    y = y.to(torch.bfloat16)
    if out is not None:
        out.copy_(y)
        return out
    return y


def nchw_to_nhwc(src, dst, c_off=0):
    _count()
    N, Cs, H, W = src.shape
    dst[:, c_off:c_off + Cs] = src.permute(0, 2, 3, 1).reshape(N * H * W, Cs).to(torch.bfloat16)
    return dst


def nhwc_to_nchw(src, n, c, hw, out):
    _count()
    out.copy_(src[:, :c].float().reshape(n, hw, c).permute(0, 2, 1).reshape(out.shape))
    return out


def upsample2x(x, n, h, w):
    _count()
    Cc = x.shape[-1]
    y = x.reshape(n, h, w, Cc).repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
    return y.reshape(n * 4 * h * w, Cc).contiguous()


def timestep_embed(t, dim, max_period=10000.0):
    _count()
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], -1).to(torch.bfloat16)


def add_silu(a, b=None, silu=True):
    _count()
    v = a + (b if b is not None else 0.0)
    return (F.silu(v) if silu else v).to(torch.bfloat16)


def copy2d(src, dst):
    _count()
    dst.copy_(src)
    return dst


def add_rows(dst, src):
    _count()
    reps = dst.shape[0] // src.shape[0]
    dst.copy_((dst.float() + src.float().repeat(reps, 1)).to(torch.bfloat16))
    return dst


def apm_mix(ctx, w, wb, ln_g, ln_b, alpha):
    _count()
    L, D = ctx.shape[1], ctx.shape[2]
    mixed = F.conv1d(ctx, w.reshape(1, L, 3), wb, padding=1)
    mixed = F.layer_norm(mixed, (D,), ln_g, ln_b, 1e-5)
    return (ctx[:, :1] + mixed * F.silu(alpha))[:, 0].to(torch.bfloat16)


def softmax_rows(scores, out=None):
    _count()
    p = torch.softmax(scores.float(), dim=-1).to(torch.bfloat16)
    if out is not None:
        out.copy_(p)
        return out
    return p


def transpose(x):
    _count()
    return x.t().contiguous()


def sampler_prepare(x, c_in, out=None):
    _count()
    v = torch.cat([x, x], 0) * c_in
    if out is not None:
        out.copy_(v)
        return out
    return v


def sampler_step(net, x, scale, *, num_frames, c_skip, c_out, sigma, next_sigma, out=None):
    _count()
    ex = (slice(None),) + (None,) * (x.dim() - 1)
    xx = torch.cat([x, x], 0)
    den = net * c_out + xx * c_skip
    d_u, d_c = den.chunk(2)
    s = scale.repeat(x.shape[0] // num_frames)[ex]
    den = d_u + s * (d_c - d_u)
    r = x + (next_sigma - sigma) * ((x - den) / sigma)
    if out is not None:
        out.copy_(r)
        return out
    return r


def attention_single_head(q, k, v, n, s):
    Cc = q.shape[1]
    out = torch.empty((n * s, Cc), dtype=torch.bfloat16)
    scale = float(Cc) ** -0.5
    for f in range(n):
        sl = slice(f * s, (f + 1) * s)
        scores = linear(q[sl], k[sl][None], None, out_fp32=True, s_acc=scale)
        probs = softmax_rows(scores)
        vt = transpose(v[sl])
        linear(probs, vt[None], None, out=out[sl])
    return out


def ddim_blend_step(noise, latents, out, *, lat_start, out_start, offset, guidance, alpha_t, alpha_prev,
                    v_prediction=True):
    _count()
    cs = noise.shape[2]
    e = noise[:1] if guidance is None else noise[:1] + guidance * (noise[1:] - noise[:1])
    x = latents[:, :, lat_start:lat_start + cs]
    sa, sb = math.sqrt(alpha_t), math.sqrt(1.0 - alpha_t)
    if v_prediction:
        x0, eps = sa * x - sb * e, sa * e + sb * x
    else:
        x0, eps = (x - sb * e) / sa, e
    res = math.sqrt(alpha_prev) * x0 + math.sqrt(1.0 - alpha_prev) * eps
    out[:, :, out_start + offset:out_start + cs] = res[:, :, offset:]
    return out


def frames_to_uint8(x, vmin=0.0, vmax=255.0):
    _count()
    return (255 * (x.clip(vmin, vmax) - vmin) / (vmax - vmin)).permute(0, 2, 3, 1).to(torch.uint8).contiguous()
