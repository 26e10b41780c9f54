"""CPU: architecture plan, tile picker, weight packing and the distributed helpers (gloo, world_size 2)."""
import os

import torch


def test_plan_topology():
    from streamingt2v_b200 import arch
    p = arch.build_plan(arch.UNetConfig(), "", decoder=True)
    assert len(p.input_blocks) == 12 and len(p.output_blocks) == 12
    assert p.skip_chans == [320, 320, 320, 320, 640, 640, 640, 1280, 1280, 1280, 1280, 1280]
    n_res = sum(isinstance(l, arch.Res) for b in p.input_blocks + [p.middle] + p.output_blocks for l in b.layers)
    n_att = sum(isinstance(l, arch.Attn) for b in p.input_blocks + [p.middle] + p.output_blocks for l in b.layers)
    assert (n_res, n_att) == (22, 16)           # 22 VideoResBlocks, 16 SpatialVideoTransformers
    cin = [b.layers[0].cin for b in p.output_blocks]
    assert cin == [2560, 2560, 2560, 2560, 2560, 1920, 1920, 1280, 960, 960, 640, 640]
    pc = arch.build_plan(arch.UNetConfig(), "", decoder=False)
    n_res_c = sum(isinstance(l, arch.Res) for b in pc.input_blocks + [pc.middle] for l in b.layers)
    assert n_res_c == 10 and not pc.output_blocks


def test_pick_box():
    from streamingt2v_b200.ops import pick_box
    for ext in [(128, 72, 50), (64, 36, 50), (32, 18, 50), (16, 9, 50), (2, 2, 16), (9216, 25, 2), (144, 25, 2), (1, 1, 1)]:
        b = pick_box(*ext)
        assert b[0] * b[1] * b[2] == 128 and all(v & (v - 1) == 0 for v in b)
    assert pick_box(128, 72, 50) == (128, 1, 1)
    assert pick_box(16, 9, 50)[0] == 16            # rows never split below the image width when it fits


def test_geglu_packing_roundtrip():
    from streamingt2v_b200 import packing
    torch.manual_seed(0)
    for k, f2 in [(320, 2560), (64, 512)]:
        w, b = torch.randn(f2, k), torch.randn(f2)
        wp, bp, bn = packing.pack_geglu(w, b, "cpu")
        x = torch.randn(10, k).to(torch.bfloat16)
        y = x.float() @ wp[0].float().t() + bp
        t = y.reshape(10, f2 // bn, 2, bn // 2)
        mine = (t[:, :, 0] * torch.nn.functional.gelu(t[:, :, 1])).reshape(10, f2 // 2)
        a, g = (x.float() @ w.to(torch.bfloat16).float().t() + b).chunk(2, -1)
        assert torch.allclose(mine, a * torch.nn.functional.gelu(g), atol=1e-4)


def test_conv_packing_layouts():
    from streamingt2v_b200 import packing
    w = torch.randn(6, 3, 3, 3)
    p = packing.pack_conv3x3(w, "cpu")
    assert p.shape == (9, 6, 8) and p[:, :, 3:].abs().sum() == 0
    assert torch.equal(p[1 * 3 + 2, :, :3], w[:, :, 1, 2].to(torch.bfloat16))
    wt = torch.randn(8, 8, 3, 1, 1)
    assert torch.equal(packing.pack_tconv3(wt, "cpu")[2], wt[:, :, 2, 0, 0].to(torch.bfloat16))


def test_synth_is_deterministic():
    from streamingt2v_b200 import arch
    shapes = {"a.weight": (4, 3), "a.bias": (4,), "n.weight": (5,), "m.time_mixer.mix_factor": (1,)}
    s1, s2 = arch.synth_state_dict(shapes, 3), arch.synth_state_dict(dict(reversed(list(shapes.items()))), 3)
    assert all(torch.equal(s1[k], s2[k]) for k in shapes) and all(float(v.abs().sum()) > 0 for v in s1.values())


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from streamingt2v_b200 import dist_utils
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mx = dist_utils.max_over_ranks([10.0 + rank, 5.0 - rank])
    items = list(dist_utils.shard_items(7, rank, world))
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, mx, items))


def test_replica_plumbing_world2():
    """N>1 path of bench.py on CPU: gloo, world_size 2 — max-over-ranks timing and work partition."""
    import torch.multiprocessing as mp
    from streamingt2v_b200 import dist_utils
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1] == [11.0, 5.0]
    assert res[0][2] + res[1][2] == list(range(7))
    assert dist_utils.aggregate_throughput(10, 2, 500.0) == 40.0


def test_diffusers_svd_key_map_is_a_bijection_onto_the_sgm_grammar():
    """First-chunk weights come in the diffusers layout (streaming_svd.py:390).  The translation table must cover
    every tensor of the plain SVD UNet exactly once, keep shapes, and follow diffusers' block structure."""
    import torch
    from streamingt2v_b200 import arch
    for cfg in (arch.UNetConfig(), arch.TINY):
        m = arch.sgm_to_diffusers_svd_keys(cfg)
        shapes = arch.plain_unet_param_shapes(cfg)
        assert set(m) == set(shapes) and len(set(m.values())) == len(m)
        assert not any("cross_attention_merger" in k for k in m)
        tops = {v.split(".")[0] for v in m.values()}
        assert tops == {"time_embedding", "add_embedding", "conv_in", "down_blocks", "mid_block", "up_blocks",
                        "conv_norm_out", "conv_out"}
        # structure: 4 down blocks with 2 resnets, attentions on the first 3; up block 0 has no attentions
        assert m["input_blocks.11.0.in_layers.2.weight"] == "down_blocks.3.resnets.1.spatial_res_block.conv1.weight"
        assert not any(v.startswith("down_blocks.3.attentions") or v.startswith("up_blocks.0.attentions") for v in m.values())
        assert m["output_blocks.2.1.conv.weight"] == "up_blocks.0.upsamplers.0.conv.weight"
        assert m["input_blocks.1.1.time_stack.0.attn1.to_q.weight"] == \
            "down_blocks.0.attentions.0.temporal_transformer_blocks.0.attn1.to_q.weight"
        assert m["input_blocks.1.1.time_pos_embed.2.bias"] == "down_blocks.0.attentions.0.time_pos_embed.linear_2.bias"
    # round trip on the reduced network: rename a synthetic SGM state dict to diffusers names and back
    cfg = arch.TINY
    sd = {k: torch.full(s_, float(i % 7)) for i, (k, s_) in enumerate(sorted(arch.plain_unet_param_shapes(cfg).items()))}
    m = arch.sgm_to_diffusers_svd_keys(cfg)
    sd_diff = {m[k]: v for k, v in sd.items()}
    back = arch.from_diffusers_svd_state_dict(sd_diff, cfg)
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    del sd_diff["conv_in.weight"]
    import pytest
    with pytest.raises(KeyError, match="missing"):
        arch.from_diffusers_svd_state_dict(sd_diff, cfg)
