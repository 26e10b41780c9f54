"""CPU (gloo, world_size 2): the two opt-in multi-GPU latency modes that follow the partitions SURVEY.md section 8(e)
names — the classifier-free-guidance halves of a denoise step (`B200EulerEDMSampler(cfg_parallel=True)`: one
all-gather of the network output per step) and the groups of <= 8 frames of the VAE decode
(`B200StreamingSVDStage(shard_decode=True)`).  Both must reproduce the single-rank result; the CUDA kernels are
replaced by the CPU stand-ins of tests/fake_ops.py (host logic and collectives are what is under test)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
T = 5


def _fake_network(xin, c_noise, cond, **kw):
    assert xin.shape[0] == cond["vector"].shape[0] == kw["image_only_indicator"].shape[0] * T
    assert kw["batch_size"] * T == xin.shape[0]
    ex = (slice(None),) + (None,) * (xin.dim() - 1)
    return torch.tanh(xin * 0.7 + 0.3) * (1.0 + 0.05 * c_noise[ex]) + 0.01 * cond["vector"][ex][..., 0]


def _inputs():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(T, 4, 6, 8, generator=g)
    cond = {"vector": torch.randn(T, 3, generator=g), "flag": 7}
    uc = {"vector": torch.zeros(T, 3), "flag": 7}
    return x, cond, uc


class _StubDecoder:
    def decode(self, z, timesteps=None):
        up = torch.nn.functional.interpolate(z[:, :3], scale_factor=8, mode="nearest")
        return torch.tanh(up * 0.05) + 0.001 * timesteps


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, HERE)
    import fake_ops
    from streamingt2v_b200 import sampler as sampler_mod
    from streamingt2v_b200.stage import B200StreamingSVDStage
    sampler_mod.ops = fake_ops                       # host logic only: CPU stand-in for the two CUDA kernels
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    x, cond, uc = _inputs()
    smp = sampler_mod.B200EulerEDMSampler(num_steps=4, num_frames=T, cfg_parallel=True)
    # rank 1 deliberately starts from a DIFFERENT latent: the sampler must broadcast rank 0's (advisor finding r1)
    out = smp(_fake_network, x.clone() + float(rank), cond, uc, batch_size=2, num_video_frames=T,
              image_only_indicator=torch.zeros(2, T))
    stage = B200StreamingSVDStage(None, smp, _StubDecoder(), None, device="cpu", shard_decode=True, max_decode_chunk=2)
    z = torch.randn(7, 4, 2, 3, generator=torch.Generator().manual_seed(4))       # 4 groups: 2, 2, 2, 1 frames
    dec = stage.decode_first_stage(z)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, out.numpy(), dec.numpy()))


def test_cfg_parallel_and_sharded_decode_world2():
    import torch.multiprocessing as mp
    sys.path.insert(0, HERE)
    import fake_ops
    from streamingt2v_b200 import sampler as sampler_mod
    from streamingt2v_b200.stage import B200StreamingSVDStage
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-rank result of the same sampler (doubled batch on one rank) and of the same decode
    saved = sampler_mod.ops
    sampler_mod.ops = fake_ops
    try:
        x, cond, uc = _inputs()
        ref = sampler_mod.B200EulerEDMSampler(num_steps=4, num_frames=T)(
            _fake_network, x.clone(), cond, uc, batch_size=2, num_video_frames=T,
            image_only_indicator=torch.zeros(2, T))
    finally:
        sampler_mod.ops = saved
    stage = B200StreamingSVDStage(None, None, _StubDecoder(), None, device="cpu", max_decode_chunk=2)
    z = torch.randn(7, 4, 2, 3, generator=torch.Generator().manual_seed(4))
    dec_ref = stage.decode_first_stage(z)
    for rank, out, dec in res:
        assert np.array_equal(out, res[0][1])                       # both ranks hold the same latents
        assert np.allclose(out, ref.numpy(), rtol=0, atol=1e-6), rank
        assert dec.shape == tuple(dec_ref.shape) and np.array_equal(dec, dec_ref.numpy()), rank
