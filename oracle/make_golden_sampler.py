"""ORACLE tooling — build-container only.  Golden vectors for the sampler arithmetic (SURVEY.md section 8 rows
a19-a21) from the UNMODIFIED reference classes, and pinning of oracle/sampler_oracle.py against them.
    python oracle/make_golden_sampler.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from oracle import sampler_oracle as sorc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def main():
    ref_shims.install()
    from models.diffusion.discretizer import AlignYourSteps
    from models.svd.sgm.modules.diffusionmodules.denoiser import Denoiser
    from models.svd.sgm.modules.diffusionmodules.guiders import LinearPredictionGuider
    from models.svd.sgm.modules.diffusionmodules.sampling import EulerEDMSampler

    # schedule (config.yaml: num_steps 30, AlignYourSteps)
    ref_sig = AlignYourSteps()(30, do_append_zero=True, device="cpu").numpy()
    mine = sorc.align_your_steps_sigmas(30)
    assert np.allclose(ref_sig, mine.astype(ref_sig.dtype), rtol=1e-6, atol=0), np.abs(ref_sig - mine).max()

    T, C, H, W = 5, 4, 6, 8
    rng = np.random.default_rng([2024, 5])
    x = torch.from_numpy(rng.normal(size=(T, C, H, W)).astype(np.float32)) * 3.0
    net_out = torch.from_numpy(rng.normal(size=(2 * T, C, H, W)).astype(np.float32))
    seen = {}

    def network(xin, c_noise, cond, **kw):
        seen["xin"], seen["c_noise"] = xin.clone(), c_noise.clone()
        return net_out

    den = Denoiser({"target": "models.svd.sgm.modules.diffusionmodules.denoiser_scaling.VScalingWithEDMcNoise"})
    sampler = EulerEDMSampler(
        discretization_config={"target": "models.diffusion.discretizer.AlignYourSteps"}, num_steps=30,
        guider_config={"target": "models.svd.sgm.modules.diffusionmodules.guiders.LinearPredictionGuider",
                       "params": {"max_scale": 3.0, "min_scale": 1.5, "num_frames": T}}, device="cpu")
    assert isinstance(sampler.guider, LinearPredictionGuider)
    cases = {}
    for idx in (0, 7, 28):
        sigma, nxt = float(ref_sig[idx]), float(ref_sig[idx + 1])
        s_in = x.new_ones([T])
        cond = {"vector": torch.zeros(T, 1)}
        out = sampler.sampler_step(s_in * sigma, s_in * nxt, lambda a, b, c: den(network, a, b, c), x, cond, cond)
        got = sorc.sampler_step(lambda a, b: network(a, b, None), x, sigma, nxt, T, 1.5, 3.0)
        err = (out - got).abs().max().item()
        assert err <= 1e-5 * out.abs().max().item(), err
        cases[f"step{idx}_sigma"] = np.float64(sigma)
        cases[f"step{idx}_next"] = np.float64(nxt)
        cases[f"step{idx}_out"] = out.numpy()
        cases[f"step{idx}_xin"] = seen["xin"].numpy()
        cases[f"step{idx}_cnoise"] = seen["c_noise"].numpy()
        print(f"step {idx}: sigma {sigma:.4f} -> {nxt:.4f} |oracle - reference| = {err:.2e}")
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN, "sampler_t5_6x8.npz"), x=x.numpy(), net_out=net_out.numpy(),
                        sigmas30=ref_sig, min_scale=1.5, max_scale=3.0, **cases)
    print("wrote tests/golden/sampler_t5_6x8.npz")


if __name__ == "__main__":
    main()
