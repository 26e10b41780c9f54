"""Full-size (config.yaml) denoiser forward timing on one B200: T=25, 72x128 latent, CFG batch 2, CAM on."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamingt2v_b200 import arch, ops, synth
from streamingt2v_b200.model import B200Denoiser


def main():
    dev = torch.device("cuda:0")
    cfg = arch.UNetConfig()
    T, h, w = int(os.environ.get("T", 25)), int(os.environ.get("H", 72)), int(os.environ.get("W", 128))
    t0 = time.time()
    sd_u = arch.synth_state_dict_device(arch.unet_param_shapes(cfg), dev, 1)
    sd_c = arch.synth_state_dict_device(arch.controlnet_param_shapes(cfg), dev, 2)
    eng = B200Denoiser(cfg, sd_u, sd_c, dev)
    del sd_u, sd_c
    torch.cuda.empty_cache()
    print(f"weights packed in {time.time() - t0:.1f}s; mem {torch.cuda.memory_allocated() / 2**30:.1f} GiB", flush=True)
    x, t, c, kw = synth.make_inputs(cfg, T=T, h=h, w=w, seed=1)
    x, t = x.to(dev), t.to(dev)
    c = {k: v.to(dev) for k, v in c.items()}
    ctrl = kw["ctrl_frames"].to(dev)
    times = []
    for it in range(6):
        l0 = ops.launches()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.time()
        e0.record()
        out = eng.forward(x, t + 0.01 * it, c, batch_size=2, num_video_frames=T, ctrl_frames=ctrl)
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        print(f"iter {it}: {times[-1]:.1f} ms (host {1e3 * (time.time() - w0):.1f} ms) launches {ops.launches() - l0} "
              f"out std {out.std().item():.3f} finite {bool(torch.isfinite(out).all())} "
              f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"ms": times, "T": T, "h": h, "w": w}, open("gpurun_out/run_full.json", "w"))


if __name__ == "__main__":
    main()
