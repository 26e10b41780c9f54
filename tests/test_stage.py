"""Stage driver: host logic of `B200StreamingSVDStage` against the independent
restatement in oracle/stage_oracle.py, with deterministic CPU stand-ins for the heavy components (the real ones are
covered by their own parity tests), plus one GPU test that runs the real wrapper + sampler + VAE decoder through it."""
import numpy as np
import pytest
import torch

T, NCOND = 6, 3


class _StubSampler:
    num_frames = T

    def __init__(self):
        self.calls = []

    def __call__(self, network, x, cond, uc, **kw):
        self.calls.append(dict(ctrl=kw["ctrl_frames"].clone(), bs=kw["batch_size"], nvf=kw["num_video_frames"],
                               ioi=tuple(kw["image_only_indicator"].shape), ncf=kw["num_conditional_frames"],
                               cshape=tuple(cond["crossattn"].shape), ushape=tuple(uc["concat"].shape)))
        return _sample(x, cond, uc, kw["ctrl_frames"])


def _sample(x, cond, uc, ctrl):
    return x * 0.5 + cond["concat"][: x.shape[0]] * 0.1 - uc["crossattn"][: x.shape[0], 0, :1, None, None] * 0.01 \
        + ctrl.mean() * 0.2


class _StubDecoder:
    def __init__(self):
        self.sizes = []

    def decode(self, z, timesteps=None):
        self.sizes.append((z.shape[0], timesteps))
        return _decode(z, timesteps)


def _decode(z, timesteps):
    up = torch.nn.functional.interpolate(z[:, :3], scale_factor=8, mode="nearest")
    return torch.tanh(up * 0.05) * 1.5 + 0.001 * timesteps          # exceeds [-1, 1] in places: exercises the clamp


def _conditioner(frame, num_frames):
    g = frame.mean()
    c = {"crossattn": torch.full((1, 2, 5), 1.0) * g, "concat": torch.ones(1, 4, frame.shape[-2] // 8,
                                                                          frame.shape[-1] // 8) * g,
         "vector": torch.arange(num_frames * 3, dtype=torch.float32).reshape(num_frames, 3)}
    uc = {k: torch.zeros_like(v) for k, v in c.items()}
    return c, uc


def test_stage_driver_matches_oracle_cpu():
    from oracle import stage_oracle
    from streamingt2v_b200.stage import B200StreamingSVDStage
    rng = np.random.default_rng(3)
    first = torch.from_numpy(rng.uniform(-1, 1, size=(T, 3, 16, 24)).astype(np.float32))
    smp, dec = _StubSampler(), _StubDecoder()
    stage = B200StreamingSVDStage(inference_model=None, sampler=smp, vae_decoder=dec, conditioner=_conditioner,
                                  num_conditional_frames=NCOND, anchor_frame=0, device="cpu", max_decode_chunk=4)
    gen = torch.Generator().manual_seed(11)
    video = stage.autoregressive_generation(first.permute(0, 2, 3, 1), 3, generator=gen)   # [F,H,W,C] input accepted

    gen2 = torch.Generator().manual_seed(11)
    noises = [torch.randn((T, 4, 2, 3), generator=gen2) for _ in range(3)]
    ref, ctrl_seen = stage_oracle.autoregressive_generation(
        first, 3, conditioner=_conditioner, sample=lambda x, c, uc, ctrl: _sample(x, c, uc, ctrl),
        decode=lambda z, n: _decode(z, n), num_frames=T, n_cond=NCOND, anchor=0, noise=lambda i: noises[i])
    assert video.shape == (T + 3 * (T - NCOND), 3, 16, 24)
    assert float(video.min()) >= 0.0 and float(video.max()) <= 255.0
    # frame bookkeeping: chunk i is conditioned on the last NCOND frames of chunk i-1
    for call, ctrl in zip(smp.calls, ctrl_seen):
        assert call["ctrl"].shape == (1, NCOND, 3, 16, 24)
        assert call["bs"] == 2 and call["nvf"] == T and call["ioi"] == (2, T) and call["ncf"] == NCOND
        assert call["cshape"] == (T, 2, 5) and call["ushape"] == (T, 4, 2, 3)
    assert torch.equal(smp.calls[0]["ctrl"], ctrl_seen[0])
    assert dec.sizes == [(4, 4), (2, 2)] * 3                      # ceil(6/4) rounds of <= 4 frames
    # with the reference's group size the whole video is identical to the oracle's
    smp8, dec8 = _StubSampler(), _StubDecoder()
    stage8 = B200StreamingSVDStage(None, smp8, dec8, _conditioner, num_conditional_frames=NCOND, anchor_frame=0,
                                   device="cpu")
    video8 = stage8.autoregressive_generation(first, 3, generator=torch.Generator().manual_seed(11))
    assert dec8.sizes == [(T, T)] * 3
    assert torch.allclose(video8, ref, rtol=0, atol=1e-4)
    for call, ctrl in zip(smp8.calls, ctrl_seen):
        assert torch.allclose(call["ctrl"], ctrl, rtol=0, atol=1e-5)


def test_stage_driver_matches_unmodified_reference_methods():
    """Row a22 pinned against the reference itself: tests/golden/stage_reference.npz was produced by the UNMODIFIED
    `StreamingSVD._autoregressive_generation` (+ _generate_conditional_output, extract_ctrl_frames,
    decode_first_stage, get_batch_sgm) running on a mock `self` with the stand-in components of
    oracle/stage_stubs.py (oracle/make_golden_stage.py).  The same scenario through B200StreamingSVDStage must give
    the same video, the same keyword arguments at the network and the same decode grouping."""
    import os

    from oracle import stage_oracle
    from oracle import stage_stubs as st
    from streamingt2v_b200.stage import B200StreamingSVDStage
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "stage_reference.npz"))
    T_, ncond, anchor, n_gen, seed = (int(v) for v in g["meta"])
    network, decoder = st.StubNetwork(), st.StubDecoder()
    stage = B200StreamingSVDStage(network, st.B200StyleSampler(T_), decoder, st.b200_style_conditioner,
                                  num_conditional_frames=ncond, device="cpu")
    assert stage.anchor_frame == anchor == 6          # default = inference_params.anchor_frames '6' (config.yaml:316)
    torch.manual_seed(seed)                            # conditioner noise + initial noise from the global RNG, like
    video = stage.autoregressive_generation(st.first_chunk(), n_gen)     # the reference (streaming_svd.py:174,196)
    # the reference hands the [0,255] float video to its IImage container, which stores uint8 [F,H,W,C]
    # (lib/farancia/libimage/iimage.py torch2np: 255 * (x.clip(vmin, vmax) - vmin) / (vmax - vmin) -> uint8)
    u8 = (255 * (video.clip(0, 255) - 0) / (255 - 0)).permute(0, 2, 3, 1).to(torch.uint8).numpy()
    assert u8.shape == g["video_u8"].shape
    assert np.array_equal(u8, g["video_u8"]), int(np.abs(u8.astype(int) - g["video_u8"].astype(int)).max())
    assert decoder.sizes == [tuple(r) for r in g["decode_sizes"].tolist()]
    calls = network.calls
    got = np.array([[c["n"], c["bs"], c["nvf"], c["ncf"], c["ioi"][0], c["ioi"][1]] for c in calls], np.int64)
    assert np.array_equal(got, g["net_calls"])
    assert tuple(calls[0]["ctrl_shape"]) == tuple(g["ctrl_shape"].tolist())
    assert np.allclose([c["ctrl_sum"] for c in calls], g["ctrl_sums"], rtol=1e-6, atol=1e-6)
    assert np.allclose([c["vec_sum"] for c in calls], g["vec_sums"], rtol=1e-6, atol=1e-6)
    # the on-device uint8 conversion (host stand-in here) gives the container array directly
    import fake_ops
    from streamingt2v_b200 import ops as real_ops
    saved = real_ops.frames_to_uint8
    real_ops.frames_to_uint8 = fake_ops.frames_to_uint8
    try:
        assert np.array_equal(stage.to_uint8_frames(video).numpy(), g["video_u8"])
    finally:
        real_ops.frames_to_uint8 = saved
    # and the independent restatement (oracle/stage_oracle.py) agrees with the reference as well
    torch.manual_seed(seed)
    net2 = st.StubNetwork()
    smp = st.B200StyleSampler(T_)

    def conditioner(frame, n):
        return st.b200_style_conditioner(frame, n)

    def sample(noise, c, uc, ctrl):
        return smp(net2, noise, c, uc, image_only_indicator=torch.zeros(2, T_), num_video_frames=T_, batch_size=2,
                   num_conditional_frames=ncond, ctrl_frames=ctrl)

    H, W = st.first_chunk().shape[-2:]
    ref_video, _ = stage_oracle.autoregressive_generation(
        st.first_chunk(), n_gen, conditioner=conditioner, sample=sample, decode=lambda z, n: st.decode_math(z, n),
        num_frames=T_, n_cond=ncond, anchor=anchor, noise=lambda i: torch.randn((T_, 4, H // 8, W // 8)))
    u8o = (255 * (ref_video.clip(0, 255) - 0) / 255).permute(0, 2, 3, 1).to(torch.uint8).numpy()
    assert np.array_equal(u8o, g["video_u8"])


def test_decode_first_stage_chunking():
    from oracle import stage_oracle
    from streamingt2v_b200.stage import B200StreamingSVDStage
    dec = _StubDecoder()
    stage = B200StreamingSVDStage(None, _StubSampler(), dec, _conditioner, device="cpu")
    z = torch.randn(25, 4, 2, 3)
    out = stage.decode_first_stage(z)
    assert dec.sizes == [(8, 8), (8, 8), (8, 8), (1, 1)]          # 8, 8, 8, 1 as in SURVEY.md row a23
    assert torch.allclose(out, stage_oracle.decode_first_stage(lambda p, n: _decode(p, n), z), atol=1e-6)


@pytest.mark.gpu
def test_stage_runs_on_real_components(cuda_dev):
    """One autoregressive generation through the real denoiser wrapper, sampler and VAE decoder (reduced UNet config,
    2 sampler steps, 8 frames at 128x128): shapes, ranges, finiteness, determinism and the conditioning hand-over."""
    import dataclasses
    from streamingt2v_b200 import arch
    from streamingt2v_b200.sampler import B200EulerEDMSampler
    from streamingt2v_b200.stage import B200StreamingSVDStage
    from streamingt2v_b200.vae import B200VaeDecoder
    from streamingt2v_b200.wrapper import B200StreamingWrapper
    Tg, ncond = 8, 3
    cfg = dataclasses.replace(arch.TINY, num_frame_conditioning=ncond)
    wrapper = B200StreamingWrapper(cfg, arch.synth_state_dict(arch.unet_param_shapes(cfg), 1),
                                   arch.synth_state_dict(arch.controlnet_param_shapes(cfg), 2), cuda_dev)
    vcfg = arch.VaeConfig()
    dec = B200VaeDecoder(vcfg, arch.synth_state_dict(arch.vae_decoder_param_shapes(vcfg), 3), cuda_dev)
    sampler = B200EulerEDMSampler(num_steps=2, num_frames=Tg)

    def conditioner(frame, num_frames):
        g = torch.Generator().manual_seed(5)
        c = {"crossattn": torch.randn(1, 1, cfg.context_dim, generator=g).to(cuda_dev),
             "concat": torch.randn(1, 4, frame.shape[-2] // 8, frame.shape[-1] // 8, generator=g).to(cuda_dev),
             "vector": torch.randn(num_frames, cfg.adm_in_channels, generator=g).to(cuda_dev)}
        return c, {k: torch.zeros_like(v) for k, v in c.items()}

    stage = B200StreamingSVDStage(wrapper, sampler, dec, conditioner, num_conditional_frames=ncond, device=cuda_dev)
    first = torch.rand(Tg, 3, 128, 128, generator=torch.Generator().manual_seed(1)) * 2 - 1
    outs = []
    for _ in range(2):
        video = stage.autoregressive_generation(first.to(cuda_dev), 1, generator=torch.Generator().manual_seed(9))
        torch.cuda.synchronize()
        outs.append(video.cpu())
    video = outs[0]
    assert video.shape == (Tg + (Tg - ncond), 3, 128, 128)
    assert torch.isfinite(video).all() and float(video.min()) >= 0.0 and float(video.max()) <= 255.0
    assert torch.equal(outs[0], outs[1])                                  # deterministic end to end
    assert torch.allclose(video[:Tg], (first + 1) * 127.5, atol=1e-3)     # first chunk passes through unchanged
    assert float(video[Tg:].std()) > 0.0
