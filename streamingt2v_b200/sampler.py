"""Host mirror of the reference sampler wrapped around the denoiser seam:
`EulerEDMSampler` with the `AlignYourSteps` discretisation, the `LinearPredictionGuider` and `Denoiser` with
`VScalingWithEDMcNoise`, as configured in code/config.yaml:139-157 (30 steps, guidance 1.5 -> 3.0 over 25 frames,
s_churn = 0).  Same call shape as the reference objects:

    sampler = B200EulerEDMSampler(num_steps=30, num_frames=25)
    z = sampler(network, x, cond, uc, **additional_model_inputs)       # sampling.py:105-127

where `network(x_in, c_noise, cond, **kw)` is the `StreamingWrapper` seam (wrappers.py:23-78) — normally a
`B200StreamingWrapper`.  Every per-step tensor operation is one of two CUDA kernels (`b200svd_sampler_prepare`,
`b200svd_sampler_step`); the 31-entry noise schedule is host arithmetic, as in the reference (numpy,
models/diffusion/discretizer.py:15-33).  There is no CPU fallback: the ops raise without the CUDA library."""
from __future__ import annotations

import math

import numpy as np
import torch

from . import dist_utils, ops

# models/diffusion/discretizer.py:29-30
AYS_SCHEDULE = (700.00, 54.5, 15.886, 7.977, 4.248, 1.789, 0.981, 0.403, 0.173, 0.034, 0.002)
_GUIDED_KEYS = ("vector", "crossattn", "concat")  # guiders.py:91


class B200EulerEDMSampler:
    def __init__(self, num_steps: int = 30, num_frames: int = 25, min_scale: float = 1.5, max_scale: float = 3.0,
                 additional_cond_keys=(), cfg_parallel: bool = False, schedule: str = "ays",
                 sigma_min: float = 0.002, sigma_max: float = 700.0, rho: float = 7.0):
        if schedule not in ("ays", "karras"):
            raise ValueError(schedule)
        # "karras": the first chunk of a request is sampled by diffusers' StableVideoDiffusionPipeline
        # (streaming_svd.py:390): EulerDiscreteScheduler with Karras sigmas 700 -> 0.002 (rho 7), 25 steps, v-prediction,
        # continuous timesteps 0.25 ln(sigma), guidance 1.0 -> 3.0 over the frames — the same Euler / EDM arithmetic as
        # the AlignYourSteps sampler of the later chunks with another noise schedule (restated from the published
        # scheduler; diffusers is absent offline: parity unpinned)
        self.schedule, self.sigma_min, self.sigma_max, self.rho = schedule, float(sigma_min), float(sigma_max), float(rho)
        self.num_steps = int(num_steps)
        self.num_frames = int(num_frames)
        self.min_scale, self.max_scale = float(min_scale), float(max_scale)
        self.additional_cond_keys = tuple(additional_cond_keys)
        # opt-in 2-rank latency mode: rank 0 evaluates the unconditional half of the doubled
        # batch, rank 1 the conditional half; one all-gather of the network output [(b t),4,h,w] per step
        self.cfg_parallel = bool(cfg_parallel)
        self._scale = None

    # -- AlignYourSteps.get_sigmas + Discretization.__call__(do_append_zero=True) ---------------------------------
    def get_sigmas(self, n=None) -> np.ndarray:
        n = self.num_steps if n is None else int(n)
        if self.schedule == "karras":
            ramp = np.linspace(0.0, 1.0, n)
            lo, hi = self.sigma_min ** (1.0 / self.rho), self.sigma_max ** (1.0 / self.rho)
            return np.concatenate([(hi + ramp * (lo - hi)) ** self.rho, [0.0]])
        t = np.asarray(AYS_SCHEDULE, dtype=np.float64)
        ys = np.interp(np.linspace(0, 1, n), np.linspace(0, 1, len(t)), np.log(t[::-1]))
        return np.concatenate([np.exp(ys)[::-1], [0.0]])

    # -- VScalingWithEDMcNoise (denoiser_scaling.py:51-59) on a python float -------------------------------------
    @staticmethod
    def scalings(sigma: float):
        c_skip = 1.0 / (sigma * sigma + 1.0)
        c_out = -sigma / math.sqrt(sigma * sigma + 1.0)
        c_in = 1.0 / math.sqrt(sigma * sigma + 1.0)
        c_noise = 0.25 * math.log(sigma)
        return c_skip, c_out, c_in, c_noise

    # -- LinearPredictionGuider.prepare_inputs for the conditioning (guiders.py:88-97) ---------------------------
    def prepare_cond(self, cond: dict, uc: dict) -> dict:
        out = {}
        for k in cond:
            if k in _GUIDED_KEYS + self.additional_cond_keys:
                out[k] = torch.cat((uc[k], cond[k]), 0)
            else:
                out[k] = cond[k]
        return out

    def _my_half(self, cond2: dict, kw: dict):
        """This rank's half of the doubled conditioning and of the per-batch keyword arguments."""
        rank, world = dist_utils.rank_world()
        if world != 2:
            raise RuntimeError(f"cfg_parallel needs exactly 2 ranks (one per guidance half), got {world}")

        def half(v):
            n = v.shape[0] // 2
            return v[rank * n:(rank + 1) * n]

        c = {k: (half(v) if k in _GUIDED_KEYS + self.additional_cond_keys else v) for k, v in cond2.items()}
        kw = dict(kw)
        if "batch_size" in kw:
            kw["batch_size"] = kw["batch_size"] // 2
        if kw.get("image_only_indicator") is not None:
            kw["image_only_indicator"] = half(kw["image_only_indicator"])
        return c, kw

    def _guider_scale(self, device):
        if self._scale is None or self._scale.device != torch.device(device):
            self._scale = torch.linspace(self.min_scale, self.max_scale, self.num_frames).to(device)  # guiders.py:71
        return self._scale

    # -- EDMSampler.sampler_step with gamma = 0 (sampling.py:82-103) ----------------------------------------------
    def sampler_step(self, network, x, sigma: float, next_sigma: float, cond2: dict, **kw):
        c_skip, c_out, c_in, c_noise = self.scalings(float(sigma))
        xin = ops.sampler_prepare(x, c_in)
        if self.cfg_parallel:
            rank, _ = dist_utils.rank_world()
            bt = x.shape[0]
            xin = xin[rank * bt:(rank + 1) * bt]          # both halves of the doubled input are identical
        t = torch.full((xin.shape[0],), c_noise, dtype=torch.float32, device=x.device)
        net = network(xin, t, cond2, **kw)
        if self.cfg_parallel:
            net = dist_utils.all_gather_cat(net.contiguous())   # rank order == (unconditional, conditional)
        return ops.sampler_step(net.contiguous(), x, self._guider_scale(x.device), num_frames=self.num_frames,
                                      c_skip=c_skip, c_out=c_out, sigma=float(sigma), next_sigma=float(next_sigma))

    # -- EDMSampler.__call__ (sampling.py:105-127; prepare_sampling_loop :42-58) ----------------------------------
    def __call__(self, network, x, cond, uc=None, num_steps=None, **kw):
        sigmas = self.get_sigmas(num_steps)
        cond2 = self.prepare_cond(cond, cond if uc is None else uc)
        if self.cfg_parallel:
            cond2, kw = self._my_half(cond2, kw)
            # both guidance halves must be evaluated on the same latent: rank 0's initial noise wins
            x = dist_utils.broadcast_from_rank0(x.to(torch.float32).contiguous())
        # x *= sqrt(1 + sigma_0^2) (sampling.py:47): the first half of the doubling kernel's output
        x = ops.sampler_prepare(x.to(torch.float32).contiguous(), math.sqrt(1.0 + float(sigmas[0]) ** 2))[:x.shape[0]]
        for i in range(len(sigmas) - 1):
            x = self.sampler_step(network, x, float(sigmas[i]), float(sigmas[i + 1]), cond2, **kw)
        return x
