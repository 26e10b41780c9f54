"""Full-size temporal VAE decode timing on one B200: 8 frames per call at 72x128 latent -> 576x1024 (reference chunking,
streaming_svd.py:127-146); 6.94 TFLOP per output frame."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from streamingt2v_b200 import arch, ops
from streamingt2v_b200.vae import B200VaeDecoder


def main():
    dev = torch.device("cuda:0")
    cfg = arch.VaeConfig()
    sd = arch.synth_state_dict_device(arch.vae_decoder_param_shapes(cfg), dev, 3)
    dec = B200VaeDecoder(cfg, sd, dev)
    n = int(os.environ.get("FRAMES", 8))
    z = torch.randn(n, 4, 72, 128, device=dev) * 5
    ms = []
    for it in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = ops.launches()
        e0.record()
        out = dec.decode(z, timesteps=n)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        print(f"iter {it}: {ms[-1]:.1f} ms for {n} frames = {ms[-1] / n:.2f} ms/frame, {6.94 * n / ms[-1] * 1e3:.0f} TFLOP/s, "
              f"launches {ops.launches() - l0}, finite {bool(torch.isfinite(out).all())}, "
              f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    with ops.profile() as prof:
        dec.decode(z, timesteps=n)
    fam = {k: dict(launches=v["launches"], ms=round(v["ms"], 2), tflops=round(v["flops"] / 1e12, 2)) for k, v in prof.families.items()}
    print(fam)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"ms": ms, "frames": n, "families": fam}, open("gpurun_out/run_vae.json", "w"))


if __name__ == "__main__":
    main()
