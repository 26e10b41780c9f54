"""Randomized-blending loop of the enhance stage (row a24, loop part): host logic of
streamingt2v_b200/blending.py against the literal restatement of the reference loop (oracle/blending_oracle.py,
parity unpinned: diffusers is absent), single process and sharded over two gloo ranks; the fused CUDA kernel is
checked on the GPU against the same oracle."""
import math
import os
import random
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
CS, OV, F, C, H, W = 6, 2, 14, 4, 3, 5            # 3 chunks: starts 0, 4, 8; 14 = 3*4 + 2


def _alphas(n=1000):
    return [math.cos((i / n + 0.008) / 1.008 * math.pi / 2) ** 2 * 0.999 + 1e-4 for i in range(n)]


def _unet(x, t, image_latents=None, image_embeddings=None, fps=None):
    g = image_latents.mean() * 0.1 + image_embeddings.mean() * 0.05
    h = torch.tanh(x * 0.6 + g) * (1.0 + 0.001 * t) + 0.01 * fps
    h[x.shape[0] // 2:] += 0.05 * x[x.shape[0] // 2:]          # conditional half differs from the unconditional one
    return h


def _inputs(device="cpu"):
    g = torch.Generator().manual_seed(7)
    lat = torch.randn(1, C, F, H, W, generator=g).to(device)
    per = [dict(image_latents=torch.randn(2, 3, generator=g).to(device), image_embeddings=torch.randn(2, 3, generator=g).to(device))
           for _ in range(3)]
    ts = [961, 921, 881, 841]
    return lat, per, ts


def _reference(lat, per, ts, seed):
    from oracle import blending_oracle as bo
    return bo.blending_loop(_unet, lat.clone(), ts, per, chunk_size=CS, overlap_size=OV, guidance_scale=9.0,
                            alphas_cumprod=_alphas(), num_train_timesteps=1000, num_inference_steps=25,
                            rng=random.Random(seed), fps=16.0)


def test_blending_host_logic_matches_reference_loop(monkeypatch):
    sys.path.insert(0, HERE)
    import fake_ops
    from streamingt2v_b200 import blending
    monkeypatch.setattr(blending, "ops", fake_ops)
    lat, per, ts = _inputs()
    ref = _reference(lat, per, ts, 33)
    bl = blending.B200RandomizedBlending(_unet, _alphas(), chunk_size=CS, overlap_size=OV, guidance_scale=9.0,
                                         rng=random.Random(33))
    out = bl(lat.clone(), ts, per, num_inference_steps=25, fps=16.0)
    assert out.shape == ref.shape and torch.allclose(out, ref, rtol=1e-5, atol=1e-5)
    # offsets are drawn in the reference's order: one per non-first chunk per timestep
    r1, r2 = random.Random(5), random.Random(5)
    offs = blending.draw_offsets(3, 3, OV, r1)
    assert offs == [[0, r2.randint(0, OV - 1), r2.randint(0, OV - 1)] for _ in range(3)]
    with pytest.raises(NotImplementedError, match="not dividable"):
        blending.chunk_starts(13, CS, OV, 3)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    import fake_ops
    from streamingt2v_b200 import blending
    blending.ops = fake_ops
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lat, per, ts = _inputs()
    # only rank 0's generator matters: the offsets are drawn there and broadcast
    bl = blending.B200RandomizedBlending(_unet, _alphas(), chunk_size=CS, overlap_size=OV, guidance_scale=9.0,
                                         rng=random.Random(33 if rank == 0 else 999), shard=True)
    out = bl(lat.clone(), ts, per, num_inference_steps=25, fps=16.0)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out.numpy()))


def test_blending_sharded_world2_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    lat, per, ts = _inputs()
    ref = _reference(lat, per, ts, 33).numpy()
    for rank, out in res:
        assert np.allclose(out, ref, rtol=1e-5, atol=1e-5), rank
    assert np.array_equal(res[0][1], res[1][1])


@pytest.mark.gpu
def test_blending_cuda_kernel_vs_reference_loop(cuda_dev):
    from streamingt2v_b200 import blending
    lat, per, ts = _inputs()
    ref = _reference(lat, per, ts, 33)
    per_d = [{k: v.to(cuda_dev) for k, v in d.items()} for d in per]
    for pred in ("v_prediction", "epsilon"):
        bl = blending.B200RandomizedBlending(_unet, _alphas(), chunk_size=CS, overlap_size=OV, guidance_scale=9.0,
                                             prediction_type=pred, rng=random.Random(33))
        out = bl(lat.to(cuda_dev), ts, per_d, num_inference_steps=25, fps=16.0).cpu()
        if pred == "v_prediction":
            assert torch.allclose(out, ref, rtol=1e-5, atol=1e-5), float((out - ref).abs().max())
        else:
            from oracle import blending_oracle as bo
            ref_e = bo.blending_loop(_unet, lat.clone(), ts, per, chunk_size=CS, overlap_size=OV, guidance_scale=9.0,
                                     alphas_cumprod=_alphas(), num_train_timesteps=1000, num_inference_steps=25,
                                     prediction_type="epsilon", rng=random.Random(33), fps=16.0)
            assert torch.allclose(out, ref_e, rtol=1e-4, atol=1e-4), float((out - ref_e).abs().max())
