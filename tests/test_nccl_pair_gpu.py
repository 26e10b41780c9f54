"""Two real GPUs, NCCL: the library's multi-GPU modes on the actual kernels (skipped on a single-GPU box; the CPU/gloo
versions of the same checks are tests/test_dist_modes.py and tests/test_blending.py).
  * B200EulerEDMSampler(cfg_parallel=True): guidance halves on two ranks, one all-gather per step — must reproduce the
    single-GPU sampler bit for bit (same kernels, same arithmetic, the halves are batch-independent);
  * B200StreamingSVDStage(shard_decode=True): VAE decode groups round-robin over the ranks;
  * B200RandomizedBlending(shard=True): chunks of a timestep over the ranks, one all-gather per timestep."""
import os
import random
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
T, H, W, STEPS = 8, 16, 8, 3


def _setup(dev):
    from streamingt2v_b200 import arch, synth
    from streamingt2v_b200.sampler import B200EulerEDMSampler
    from streamingt2v_b200.vae import B200VaeDecoder
    from streamingt2v_b200.wrapper import B200StreamingWrapper
    cfg, vcfg = arch.TINY, arch.VaeConfig()
    model = B200StreamingWrapper(cfg, arch.synth_state_dict_fast(arch.unet_param_shapes(cfg), 3),
                                 arch.synth_state_dict_fast(arch.controlnet_param_shapes(cfg), 4), dev)
    dec = B200VaeDecoder(vcfg, arch.synth_state_dict_fast(arch.vae_decoder_param_shapes(vcfg), 5), dev)
    _, _, c, kw = synth.make_inputs(cfg, T=T, h=H, w=W, B=1, seed=9)
    cond = {k: v.to(dev) for k, v in c.items()}
    uc = {"crossattn": torch.zeros_like(cond["crossattn"]), "concat": torch.zeros_like(cond["concat"]),
          "vector": cond["vector"].clone()}
    extra = dict(image_only_indicator=torch.zeros(2, T, device=dev), num_video_frames=T, batch_size=2,
                 num_conditional_frames=cfg.num_frame_conditioning, ctrl_frames=kw["ctrl_frames"].to(dev))
    noise = torch.from_numpy(np.random.default_rng(5).normal(size=(T, 4, H, W)).astype(np.float32)).to(dev)
    return model, dec, cond, uc, extra, noise, B200EulerEDMSampler


def _blend_unet(x, t, image_latents=None):
    return torch.tanh(x * 0.6 + image_latents.mean() * 0.1) * (1.0 + 0.001 * t)


def _blend_inputs(dev):
    g = torch.Generator().manual_seed(7)
    lat = torch.randn(1, 4, 14, 4, 6, generator=g).to(dev)
    per = [dict(image_latents=torch.randn(2, 3, generator=g).to(dev)) for _ in range(3)]
    alphas = [float(np.cos((i / 1000 + 0.008) / 1.008 * np.pi / 2) ** 2 * 0.999 + 1e-4) for i in range(1000)]
    return lat, per, [961, 921, 881], alphas


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from streamingt2v_b200.blending import B200RandomizedBlending
    from streamingt2v_b200.stage import B200StreamingSVDStage
    model, dec, cond, uc, extra, noise, Sampler = _setup(dev)
    smp = Sampler(num_steps=STEPS, num_frames=T, cfg_parallel=True)
    # rank 1 deliberately starts from another latent: rank 0's must win
    z = smp(model, noise + float(rank), cond, uc, **extra)
    stage = B200StreamingSVDStage(model, smp, dec, None, device=dev, shard_decode=True, max_decode_chunk=3)
    frames = stage.decode_first_stage(z)
    lat, per, ts, alphas = _blend_inputs(dev)
    bl = B200RandomizedBlending(_blend_unet, alphas, chunk_size=6, overlap_size=2, guidance_scale=9.0,
                                rng=random.Random(33 if rank == 0 else 1), shard=True)
    blended = bl(lat, ts, per, num_inference_steps=25)
    torch.cuda.synchronize()
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, z.cpu().numpy(), frames.cpu().numpy(), blended.cpu().numpy()))


def test_multi_gpu_modes_on_nccl_match_single_gpu(cuda_dev):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from streamingt2v_b200.blending import B200RandomizedBlending
    from streamingt2v_b200.stage import B200StreamingSVDStage
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + os.getpid() % 100
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=600) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    model, dec, cond, uc, extra, noise, Sampler = _setup(cuda_dev)
    z_ref = Sampler(num_steps=STEPS, num_frames=T)(model, noise, cond, uc, **extra)
    frames_ref = B200StreamingSVDStage(model, None, dec, None, device=cuda_dev, max_decode_chunk=3).decode_first_stage(z_ref)
    lat, per, ts, alphas = _blend_inputs(cuda_dev)
    blended_ref = B200RandomizedBlending(_blend_unet, alphas, chunk_size=6, overlap_size=2, guidance_scale=9.0,
                                         rng=random.Random(33))(lat, ts, per, num_inference_steps=25)
    zr = z_ref.cpu().numpy()
    assert np.array_equal(res[0][1], res[1][1]), "the two ranks hold different latents"
    for rank, z, frames, blended in res:
        # a batch-1 forward per guidance half against one batch-2 forward: same arithmetic per video; the GroupNorm
        # chunking may depend on the number of samples in a launch, so allow the last bits (fp32 sampler state)
        rel = np.abs(z - zr).max() / np.abs(zr).max()
        print(f"rank {rank}: cfg-parallel vs single GPU max rel diff {rel:.3e}")
        assert rel < 2e-3, f"rank {rank}: cfg-parallel sampler differs from one GPU ({rel})"
        # the sharded decode of THIS latent must equal its single-GPU decode bit for bit
        stage1 = B200StreamingSVDStage(model, None, dec, None, device=cuda_dev, max_decode_chunk=3)
        assert np.array_equal(frames, stage1.decode_first_stage(torch.from_numpy(z).to(cuda_dev)).cpu().numpy()), \
            f"rank {rank}: sharded decode differs"
        assert np.array_equal(blended, blended_ref.cpu().numpy()), f"rank {rank}: sharded blending differs"
    del frames_ref
