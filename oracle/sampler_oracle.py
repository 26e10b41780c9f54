"""ORACLE — test infrastructure only (imported by tests/, smoke() and bench.py's CPU arm; never by the product).

CPU restatement (numpy / torch fp32) of the sampler arithmetic wrapped around the denoiser seam, rows a19-a21 of
SURVEY.md section 8:
  * VScalingWithEDMcNoise           code/models/svd/sgm/modules/diffusionmodules/denoiser_scaling.py:51-59
  * Denoiser.forward                .../diffusionmodules/denoiser.py:23-39
  * LinearPredictionGuider          .../diffusionmodules/guiders.py:60-99
  * EDMSampler.sampler_step / Euler .../diffusionmodules/sampling.py:82-103,211-215 (to_d: sampling_utils)
  * AlignYourSteps.get_sigmas       code/models/diffusion/discretizer.py:8-33 (+ the appended 0 of
                                    Discretization.__call__, .../diffusionmodules/discretizer.py)
Pinned against the unmodified reference classes by oracle/make_golden_sampler.py -> tests/golden/sampler_*.npz."""
from __future__ import annotations

import numpy as np
import torch

AYS_SCHEDULE = [700.00, 54.5, 15.886, 7.977, 4.248, 1.789, 0.981, 0.403, 0.173, 0.034, 0.002]


def v_scaling_edm_cnoise(sigma: torch.Tensor):
    """denoiser_scaling.py:51-59 -> (c_skip, c_out, c_in, c_noise)."""
    c_skip = 1.0 / (sigma ** 2 + 1.0)
    c_out = -sigma / (sigma ** 2 + 1.0) ** 0.5
    c_in = 1.0 / (sigma ** 2 + 1.0) ** 0.5
    c_noise = 0.25 * sigma.log()
    return c_skip, c_out, c_in, c_noise


def align_your_steps_sigmas(n: int) -> np.ndarray:
    """discretizer.py:15-33 log-linear interpolation of the 11-point schedule to n points, then the trailing 0
    (Discretization.__call__ with do_append_zero=True); float64 like the reference's numpy path."""
    t = np.asarray(AYS_SCHEDULE, dtype=np.float64)
    xs = np.linspace(0, 1, len(t))
    ys = np.log(t[::-1])
    new_ys = np.interp(np.linspace(0, 1, n), xs, ys)
    return np.concatenate([np.exp(new_ys)[::-1], [0.0]])


def guider_scale(num_frames: int, min_scale: float, max_scale: float) -> torch.Tensor:
    """guiders.py:71 torch.linspace(min_scale, max_scale, num_frames)."""
    return torch.linspace(min_scale, max_scale, num_frames)


def sampler_step(network, x, sigma, next_sigma, num_frames, min_scale, max_scale):
    """One EulerEDMSampler step (gamma = 0) with the doubled-batch LinearPredictionGuider.
    x: [(b t), ...]; sigma / next_sigma: python floats; network(x_in[2bt], c_noise[2bt]) -> [2bt, ...]
    (rows [0, bt) unconditional, [bt, 2bt) conditional, guiders.py:88-97)."""
    bt = x.shape[0]
    s = torch.full((2 * bt,), float(sigma), dtype=x.dtype)
    xin = torch.cat([x, x], 0)
    c_skip, c_out, c_in, c_noise = v_scaling_edm_cnoise(s)
    ex = (slice(None),) + (None,) * (x.dim() - 1)
    den = network(xin * c_in[ex], c_noise) * c_out[ex] + xin * c_skip[ex]          # denoiser.py:33-39
    x_u, x_c = den.chunk(2)
    scale = guider_scale(num_frames, min_scale, max_scale).repeat(bt // num_frames)[ex]
    den = x_u + scale * (x_c - x_u)                                              # guiders.py:78-86
    d = (x - den) / float(sigma)                                                 # to_d
    return x + (float(next_sigma) - float(sigma)) * d                            # sampling.py:100-103,213-215
