"""Sampler arithmetic around the denoiser seam.
CPU: the oracle against the golden vectors produced from the unmodified reference classes
(oracle/make_golden_sampler.py), and the host mirror `B200EulerEDMSampler` (schedule, conditioning doubling, loop
structure) against the oracle with the CPU stand-in ops.  GPU: the two CUDA kernels through the C ABI against the
same golden vectors — fp32 elementwise work evaluated in the reference's operation order, tolerance 2 ulp-ish:
|err| <= 2e-6 * max|ref| (FMA contraction and the reciprocal-free division are the only freedom)."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_t5_6x8.npz")
STEPS = (0, 7, 28)


def _golden():
    return np.load(GOLDEN)


def test_oracle_schedule_and_steps_match_reference_golden():
    from oracle import sampler_oracle as sorc
    g = _golden()
    assert np.allclose(sorc.align_your_steps_sigmas(30), g["sigmas30"].astype(np.float64), rtol=1e-6, atol=0)
    x, net = torch.from_numpy(g["x"]), torch.from_numpy(g["net_out"])
    for i in STEPS:
        seen = {}

        def network(xin, c_noise):
            seen["xin"], seen["c"] = xin, c_noise
            return net

        out = sorc.sampler_step(network, x, float(g[f"step{i}_sigma"]), float(g[f"step{i}_next"]), 5,
                                float(g["min_scale"]), float(g["max_scale"]))
        ref = torch.from_numpy(g[f"step{i}_out"])
        assert (out - ref).abs().max() <= 1e-5 * ref.abs().max()
        assert torch.allclose(seen["xin"], torch.from_numpy(g[f"step{i}_xin"]), rtol=1e-6, atol=1e-7)
        assert torch.allclose(seen["c"], torch.from_numpy(g[f"step{i}_cnoise"]), rtol=1e-6, atol=1e-7)


def _fake_network(seed):
    """Deterministic stand-in for the denoiser network: a fixed elementwise function of its inputs."""
    def network(xin, c_noise, cond, **kw):
        ex = (slice(None),) + (None,) * (xin.dim() - 1)
        return torch.tanh(xin * 0.7 + 0.1 * seed) * (1.0 + 0.05 * c_noise[ex]) + 0.01 * cond["vector"][ex][..., 0]
    return network


def _oracle_loop(x, cond, uc, num_steps, T):
    from oracle import sampler_oracle as sorc
    sig = sorc.align_your_steps_sigmas(num_steps)
    cc = {"vector": torch.cat([uc["vector"], cond["vector"]], 0)}
    net = _fake_network(3)
    x = x * float(np.sqrt(1.0 + sig[0] ** 2))
    for i in range(num_steps):
        x = sorc.sampler_step(lambda a, b: net(a, b, cc), x, float(sig[i]), float(sig[i + 1]), T, 1.5, 3.0)
    return x


def test_host_sampler_loop_matches_oracle_cpu(monkeypatch):
    import fake_ops
    from streamingt2v_b200 import sampler as sampler_mod
    from streamingt2v_b200.sampler import B200EulerEDMSampler
    monkeypatch.setattr(sampler_mod, "ops", fake_ops)   # host logic only: CPU stand-in for the two CUDA kernels
    T = 5
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2 * T, 4, 6, 8, generator=g)      # b = 2 videos of T frames
    cond = {"vector": torch.randn(2 * T, 3, generator=g), "flag": 7}
    uc = {"vector": torch.zeros(2 * T, 3), "flag": 7}
    smp = B200EulerEDMSampler(num_steps=6, num_frames=T)
    out = smp(_fake_network(3), x.clone(), cond, uc)
    ref = _oracle_loop(x.clone(), cond, uc, 6, T)
    assert out.shape == x.shape
    assert (out - ref).abs().max() <= 1e-5 * ref.abs().max()
    assert np.allclose(smp.get_sigmas(30), _golden()["sigmas30"], rtol=1e-6)
    c2 = smp.prepare_cond(cond, uc)
    assert c2["vector"].shape[0] == 4 * T and c2["flag"] == 7


@pytest.mark.gpu
def test_sampler_kernels_match_reference_golden(cuda_dev):
    from streamingt2v_b200 import ops
    from streamingt2v_b200.sampler import B200EulerEDMSampler
    g = _golden()
    x = torch.from_numpy(g["x"]).to(cuda_dev)
    net = torch.from_numpy(g["net_out"]).to(cuda_dev)
    scale = torch.linspace(float(g["min_scale"]), float(g["max_scale"]), 5).to(cuda_dev)
    for i in STEPS:
        sigma, nxt = float(g[f"step{i}_sigma"]), float(g[f"step{i}_next"])
        c_skip, c_out, c_in, c_noise = B200EulerEDMSampler.scalings(sigma)
        xin = ops.sampler_prepare(x, c_in)
        ref_in = torch.from_numpy(g[f"step{i}_xin"]).to(cuda_dev)
        assert (xin - ref_in).abs().max() <= 2e-6 * ref_in.abs().max()
        assert abs(c_noise - float(g[f"step{i}_cnoise"][0])) <= 1e-6 * max(1.0, abs(c_noise))
        out = ops.sampler_step(net, x, scale, num_frames=5, c_skip=c_skip, c_out=c_out, sigma=sigma, next_sigma=nxt)
        torch.cuda.synchronize()
        ref = torch.from_numpy(g[f"step{i}_out"]).to(cuda_dev)
        err = (out - ref).abs().max().item()
        print(f"sampler step {i}: max_abs_err {err:.3e} of {ref.abs().max().item():.3e}")
        assert err <= 2e-6 * ref.abs().max().item()


@pytest.mark.gpu
def test_sampler_loop_gpu_matches_oracle(cuda_dev):
    from streamingt2v_b200.sampler import B200EulerEDMSampler
    T = 5
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2 * T, 4, 6, 8, generator=g)
    cond = {"vector": torch.randn(2 * T, 3, generator=g)}
    uc = {"vector": torch.zeros(2 * T, 3)}
    smp = B200EulerEDMSampler(num_steps=6, num_frames=T)
    out = smp(_fake_network(3), x.to(cuda_dev), {k: v.to(cuda_dev) for k, v in cond.items()},
              {k: v.to(cuda_dev) for k, v in uc.items()})
    ref = _oracle_loop(x.clone(), cond, uc, 6, T)
    # tanh on the GPU vs the CPU differs by a few ulp and is amplified by 1/sigma in the last steps
    assert (out.cpu() - ref).abs().max() <= 1e-4 * ref.abs().max()


def test_karras_schedule_of_the_first_chunk():
    """EulerDiscreteScheduler(use_karras_sigmas) as StableVideoDiffusionPipeline uses it for chunk 1: 25 sigmas from 700
    to 0.002 with rho = 7, then 0 (published formula; diffusers absent: parity unpinned)."""
    import numpy as np
    from streamingt2v_b200.sampler import B200EulerEDMSampler
    smp = B200EulerEDMSampler(num_steps=25, num_frames=25, min_scale=1.0, max_scale=3.0, schedule="karras")
    sig = smp.get_sigmas()
    assert sig.shape == (26,) and sig[-1] == 0.0
    assert abs(sig[0] - 700.0) < 1e-9 and abs(sig[24] - 0.002) < 1e-12 and np.all(np.diff(sig) < 0)
    i = 7
    expect = (700.0 ** (1 / 7) + i / 24 * (0.002 ** (1 / 7) - 700.0 ** (1 / 7))) ** 7
    assert abs(sig[i] - expect) < 1e-9 * expect
