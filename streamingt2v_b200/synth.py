"""Deterministic synthetic inputs for the denoiser seam (no datasets / checkpoints are available offline).

Shapes follow the call made by the reference stage driver (code/diffusion_trainer/streaming_svd.py:186-216):
x [(B T),4,h,w], t = c_noise [(B T)], c = {concat [(B T),4,h,w], crossattn [(B T),L,1024], vector [(B T),768]},
ctrl_frames [1,7,3,8h,8w] in [-1,1].  Values come from numpy Generators seeded by (seed, crc32(name)) so that the
build container (golden generation) and the GPU box (tests, bench) see identical tensors.
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch


def _rng(seed: int, name: str):
    return np.random.default_rng([seed, zlib.crc32(name.encode())])


def make_inputs(cfg, *, T: int, h: int, w: int, B: int = 2, seed: int = 1, sigma: float = 4.0, ctx_tokens: int = 1):
    n = B * T

    def normal(name, shape, scale=1.0):
        return torch.from_numpy((_rng(seed, name).normal(size=shape) * scale).astype(np.float32))

    c_in = 1.0 / math.sqrt(sigma * sigma + 1.0)                      # denoiser_scaling.py:51-59
    x = normal("x", (n, 4, h, w), sigma * c_in)
    t = torch.full((n,), 0.25 * math.log(sigma), dtype=torch.float32)  # c_noise
    concat = normal("concat", (n, 4, h, w), 1.0)
    cross = normal("crossattn", (n, ctx_tokens, cfg.context_dim), 1.0)
    # vector = concat of three sinusoidal scalar embeddings in the reference; any bounded vector serves
    vec = torch.from_numpy(np.cos(_rng(seed, "vector").uniform(0, 2 * math.pi, size=(n, cfg.adm_in_channels))
                                  ).astype(np.float32))
    ctrl = torch.from_numpy(_rng(seed, "ctrl_frames").uniform(-1, 1, size=(1, cfg.num_frame_conditioning, 3, 8 * h,
                                                                           8 * w)).astype(np.float32))
    c = {"concat": concat, "crossattn": cross, "vector": vec}
    kwargs = dict(batch_size=B, num_video_frames=T, image_only_indicator=torch.zeros(B, T), ctrl_frames=ctrl,
                  num_conditional_frames=cfg.num_frame_conditioning)
    return x, t, c, kwargs
