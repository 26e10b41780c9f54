"""ORACLE tooling — deterministic CPU stand-ins for the heavy components of the StreamingSVD stage (sampler, denoiser
network, temporal VAE decoder, conditioner).  They let the UNMODIFIED reference stage driver
(code/diffusion_trainer/streaming_svd.py:124-356, run as unbound methods on a mock `self` by
oracle/make_golden_stage.py) and `streamingt2v_b200/stage.py` execute the SAME arithmetic, so that the frame
bookkeeping of row a22 (which frames condition which chunk, decode grouping, clamping, range conversion, keyword
arguments reaching the network) is compared bit for bit.  Test infrastructure only."""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch

NUM_FRAMES, NCOND, ANCHOR = 8, 3, 6
CTX_DIM, VEC_DIM = 5, 3


# ---- the "network" behind the seam: records what reaches it --------------------------------------------------------
class StubNetwork:
    """Stands for StreamingWrapper.forward(x, t, c, **kwargs) (wrappers.py:23-78)."""

    def __init__(self):
        self.calls = []

    def __call__(self, x, t, c, **kw):
        ctrl = kw["ctrl_frames"]
        self.calls.append(dict(n=x.shape[0], bs=kw["batch_size"], nvf=kw["num_video_frames"],
                               ncf=kw["num_conditional_frames"], ioi=tuple(kw["image_only_indicator"].shape),
                               ctrl_shape=tuple(ctrl.shape), ctrl_sum=float(ctrl.double().sum()),
                               vec_sum=float(c["vector"].double().sum())))
        ex = (slice(None),) + (None,) * (x.dim() - 1)
        g = c["concat"] * 0.1 + c["crossattn"][:, 0, :1, None, None] * 0.05
        return torch.tanh(x * 0.7 + g) * (1.0 + 0.01 * t[ex]) + 0.02 * ctrl.mean() + 0.001 * c["vector"][ex][..., 0]


def stub_denoiser(network, x, sigma, c, **kw):
    """Stands for Denoiser.forward (denoiser.py:23-39): fixed scalings instead of VScalingWithEDMcNoise."""
    ex = (slice(None),) + (None,) * (x.dim() - 1)
    return network(x * 0.5, 0.25 * sigma.log(), c, **kw) * 0.3 + x * (1.0 / (1.0 + sigma[ex]))


def sample_math(denoise, x, cond, uc, num_frames):
    """Three guided 'steps' with the doubled batch of LinearPredictionGuider (guiders.py:88-97).
    denoise(x2, sigma2, c2) -> [2n, ...]."""
    c2 = {k: (torch.cat((uc[k], cond[k]), 0) if k in ("vector", "crossattn", "concat") else cond[k]) for k in cond}
    scale = torch.linspace(1.5, 3.0, num_frames)[:, None, None, None]
    for sigma in (4.0, 1.5, 0.3):
        s2 = torch.full((2 * x.shape[0],), sigma)
        d = denoise(torch.cat([x, x]), s2, c2)
        x_u, x_c = d.chunk(2)
        x = 0.6 * x + 0.4 * (x_u + scale * (x_c - x_u))
    return x


class RefStyleSampler:
    """Call convention of the reference: sampler(denoiser_closure, randn, cond=c, uc=uc) (streaming_svd.py:214-216)."""

    def __init__(self, num_frames=NUM_FRAMES):
        self.guider = SimpleNamespace(num_frames=num_frames)

    def __call__(self, denoiser, x, cond=None, uc=None):
        return sample_math(denoiser, x, cond, uc, self.guider.num_frames)


class B200StyleSampler:
    """Call convention of B200EulerEDMSampler: sampler(network, x, cond, uc, **additional_model_inputs)."""

    def __init__(self, num_frames=NUM_FRAMES):
        self.num_frames = num_frames

    def __call__(self, network, x, cond, uc, **kw):
        return sample_math(lambda a, s, c: stub_denoiser(network, a, s, c, **kw), x, cond, uc, self.num_frames)


# ---- temporal VAE decoder -------------------------------------------------------------------------------------------
def decode_math(z, timesteps):
    up = torch.nn.functional.interpolate(z[:, :3], scale_factor=8, mode="nearest")
    ramp = torch.arange(z.shape[0], dtype=torch.float32)[:, None, None, None]
    return torch.tanh(up * 0.05) * 1.5 + 0.001 * float(timesteps) + 0.002 * ramp     # exceeds [-1,1]: exercises clamp


class StubDecoder:
    def __init__(self):
        self.sizes = []

    def decode(self, z, timesteps=None):
        self.sizes.append((z.shape[0], timesteps))
        return decode_math(z, timesteps)


# ---- conditioner ------------------------------------------------------------------------------------------------------
def cond_math(cond_frames_without_noise, cond_frames, fps_id, motion_bucket_id, cond_aug):
    """Stand-in for GeneralConditioner: crossattn from the clean frame, concat from the noised frame (pooled to the
    latent grid), vector from the three scalar ids (one row per video frame)."""
    crossattn = cond_frames_without_noise.mean(dim=(1, 2, 3))[:, None, None] * torch.linspace(1, 2, CTX_DIM)[None, None]
    concat = torch.nn.functional.avg_pool2d(cond_frames, 8)
    concat = torch.cat([concat, concat[:, :1]], 1)                                   # 4 latent channels
    vector = torch.stack([fps_id.float(), motion_bucket_id.float() / 100.0, cond_aug.float() * 10.0], -1)
    return {"crossattn": crossattn, "concat": concat, "vector": vector}


class RefStyleConditioner:
    """The interface streaming_svd.py:176-188 uses: .embedders[i].input_key and get_unconditional_conditioning."""
    embedders = [SimpleNamespace(input_key=k) for k in ("cond_frames_without_noise", "fps_id", "motion_bucket_id",
                                                        "cond_frames", "cond_aug")]

    def get_unconditional_conditioning(self, batch_c, batch_uc=None, force_uc_zero_embeddings=None):
        c = cond_math(batch_c["cond_frames_without_noise"], batch_c["cond_frames"], batch_c["fps_id"],
                      batch_c["motion_bucket_id"], batch_c["cond_aug"])
        uc = cond_math(batch_uc["cond_frames_without_noise"], batch_uc["cond_frames"], batch_uc["fps_id"],
                       batch_uc["motion_bucket_id"], batch_uc["cond_aug"])
        force = set(force_uc_zero_embeddings or ())
        if "cond_frames_without_noise" in force:
            uc["crossattn"] = torch.zeros_like(uc["crossattn"])
        if "cond_frames" in force:
            uc["concat"] = torch.zeros_like(uc["concat"])
        return c, uc


def b200_style_conditioner(frame, num_frames):
    """`conditioner(svd_input_frame, num_frames) -> (c, uc)` as injected into B200StreamingSVDStage: the value dict of
    streaming_svd.py:166-176 (motion bucket 127, fps id 6, cond_aug 0.02 with UNIFORM noise) fed to the same math."""
    image = frame[None, :]
    cond_aug = 0.02
    noised = image + cond_aug * torch.rand_like(image)
    n = num_frames
    args = (image, noised, torch.tensor([6]).repeat(n), torch.tensor([127]).repeat(n), torch.tensor([cond_aug]).repeat(n))
    c = cond_math(*args)
    uc = cond_math(*args)
    uc["crossattn"] = torch.zeros_like(uc["crossattn"])
    uc["concat"] = torch.zeros_like(uc["concat"])
    return c, uc


def first_chunk(seed=3, frames=NUM_FRAMES, H=16, W=24):
    g = torch.Generator().manual_seed(seed)
    return torch.rand((frames, 3, H, W), generator=g) * 2.0 - 1.0
