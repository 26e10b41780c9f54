#!/usr/bin/env python
"""bench.py — denoise-step throughput of the StreamingSVD hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" = ONE pass of the hot path over one batch: StreamingWrapper.forward (ControlNet on 2x7 frames + VideoUNet
with 13 CAM mergers on 2x25 frames, 576x1024 -> latent 72x128, classifier-free-guidance batch 2) followed by the
denoiser scalings, guider combine and Euler update of the sampler (denoiser.py:33-39, guiders.py:78-86,
sampling.py:100-103; two fused elementwise kernels, streamingt2v_b200/sampler.py) — the work of one of
the 150 autoregressive denoise steps of a 200-frame request.  181.96 TFLOP algorithmic.
Weights are random-init of the shipped architecture (no checkpoints offline), inputs synthetic; bf16 compute.

Timed region (per rank): [invalidate conditioning cache] K x forward(+sampler math), bracketed by barrier +
torch.cuda.synchronize(), CUDA events on the launching stream, max over ranks.  The step-invariant conditioning
work (ControlNet conditioning embedding, cross-attention vectors) is executed once INSIDE the timed region, as it
is once per 30-step chunk in the real pipeline.
  value  : steps/s with inputs resident in HBM (whole job: N replicas x per-GPU rate; the path does not shard —
           "replicas only").
  e2e    : same through the public module API (B200StreamingWrapper.forward) with x, t read from pinned HOST
           memory and the result copied back to the host EVERY step; conditioning uploaded from the host once
           inside the timed region (it is constant over a chunk).
  roofline / cpu_baseline: see DESIGN.md §measurement.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

STEP_TFLOP = 181.96          # UNet + 13 CAM mergers + ControlNet, T=25, 72x128, CFG 2
FULL_FRAME_PIXELS = 2 * 25 * 72 * 128
METRIC = "denoise_steps_per_sec"
UNIT = "steps/s"


def _workload(args):
    return dict(workload="StreamingSVD denoise step (ControlNet+VideoUNet+CAM), 25 frames, 576x1024 (latent 72x128), "
                         "CFG batch 2 — BASELINE configs[1]/[2] per-step unit",
                frames=25, latent=[72, 128], cfg_batch=2, cam=True, step_tflop=STEP_TFLOP,
                l2="inputs larger than L2 (activations ~19 GB/step >> 126 MB)", parallelism=f"replicas x{args.gpus}")


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi) during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def _usable_cores(cap=32):
    """Host threads the CPU arm may really use: scheduler affinity and cgroup CPU quota (os.cpu_count() reports the
    machine, not the container: 128 threads on a quota of a few cores made one sample forward take 140 s instead of
    ~10 s), capped at 32 -- the bounded sample's convolutions / matmuls do not scale past that."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as fh:
            quota, period = fh.read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        pass
    return max(1, min(n, cap))


# ----------------------------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (oracle port, fp32, torch CPU kernels, all host threads)
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_sample(runs=1, warmup=0, budget_s=200.0):
    """Time the oracle (CPU restatement of StreamingWrapper.forward, pinned against the reference) on a bounded
    sample of the SAME full-size network: B=2, T=8 frames, 32x32 latent (=1/28.1 of the full frame-pixels), scaled
    to full-step units by the frame-pixel ratio (attention is super-linear in pixels, so this flatters the CPU)."""
    import torch
    from oracle import streaming_svd_oracle as orc
    from streamingt2v_b200 import arch, synth
    cores = _usable_cores()
    torch.set_num_threads(cores)
    cfg = arch.UNetConfig()
    g = torch.Generator().manual_seed(0)

    def rnd(shapes):
        sd = {}
        for k, s in shapes.items():
            fan = 1
            for d in s[1:]:
                fan *= d
            if k.endswith("weight") and len(s) == 1:
                sd[k] = torch.ones(s)
            elif len(s) <= 1:
                sd[k] = torch.zeros(s) if k.endswith("bias") else torch.full(s, 0.5)
            else:
                sd[k] = torch.randn(s, generator=g) * fan ** -0.5
        return sd

    sd_u, sd_c = rnd(arch.unet_param_shapes(cfg)), rnd(arch.controlnet_param_shapes(cfg))
    T, h, w = 8, 32, 32
    x, t, c, kw = synth.make_inputs(cfg, T=T, h=h, w=w, seed=1)
    times = []
    t_begin = time.time()
    for i in range(warmup + runs):
        t0 = time.time()
        with torch.no_grad():
            orc.streaming_wrapper_forward(sd_u, sd_c, cfg, x, t, c, **kw)
        dt = time.time() - t0
        if i >= warmup or (time.time() - t_begin) > budget_s:
            times.append(dt)
        if (time.time() - t_begin) > budget_s:
            break
    dt = sum(times) / len(times)
    ratio = FULL_FRAME_PIXELS / (2 * T * h * w)
    return dict(value=1.0 / (dt * ratio), unit=UNIT, cores=cores, kind="port",
                sample=f"oracle port (fp32 torch-CPU) of StreamingWrapper.forward, full-size weights, B=2 T={T} "
                       f"latent {h}x{w}: {dt:.2f}s/forward over {len(times)} run(s); scaled x{ratio:.1f} by "
                       f"frame-pixels to the 25-frame 72x128 step",
                sample_seconds=dt)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_sample(runs=args.steps, warmup=args.warmup)
    v = cb["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _workload(args), "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from streamingt2v_b200 import arch, dist_utils, ops, synth
    from streamingt2v_b200.sampler import B200EulerEDMSampler
    from streamingt2v_b200.wrapper import B200StreamingWrapper

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for the product arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner (and NCCL_DEBUG output) on stdout when the communicator is created: keep
        # stdout for the one JSON line
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    cfg = arch.UNetConfig()
    T, h, w, B = 25, 72, 128, 2
    sd_u = arch.synth_state_dict_device(arch.unet_param_shapes(cfg), dev, 1)
    sd_c = arch.synth_state_dict_device(arch.controlnet_param_shapes(cfg), dev, 2)
    model = B200StreamingWrapper(cfg, sd_u, sd_c, dev)
    del sd_u, sd_c
    torch.cuda.empty_cache()
    x, t, c, kw = synth.make_inputs(cfg, T=T, h=h, w=w, seed=1 + rank)
    # host (pinned) copies for the e2e leg, device copies for the resident leg
    host = dict(x=x.pin_memory(), t=t.pin_memory(), ctrl=kw["ctrl_frames"].pin_memory(),
                **{k: v.pin_memory() for k, v in c.items()})
    xd, td = x.to(dev), t.to(dev)
    cd = {k: v.to(dev) for k, v in c.items()}
    ctrl_d = kw["ctrl_frames"].to(dev)
    scale = torch.linspace(1.5, 3.0, T).to(dev)                          # LinearPredictionGuider (guiders.py:60-86)
    sigmas = torch.exp(torch.linspace(math.log(700.0), math.log(0.002), args.steps + args.warmup + 2)).tolist()

    def step(cur, tin, cc, ctrl, i):
        """one sampler step around the seam (EulerEDMSampler.sampler_step, gamma = 0): input scaling + batch doubling
        (kernel), denoiser forward (the hot path), output scaling + CFG combine + Euler update (kernel).
        cur: the latent state [T,4,h,w]; returns the next state."""
        sig, sig_next = sigmas[i], sigmas[i + 1]
        c_skip, c_out, c_in, c_noise = B200EulerEDMSampler.scalings(sig)  # denoiser_scaling.py:51-59
        tin.fill_(c_noise)
        xin2 = ops.sampler_prepare(cur, c_in)
        net = model(xin2, tin, cc, batch_size=B, num_video_frames=T, image_only_indicator=None, ctrl_frames=ctrl,
                    num_conditional_frames=7)
        return ops.sampler_step(net, cur, scale, num_frames=T, c_skip=c_skip, c_out=c_out, sigma=sig,
                                next_sigma=sig_next)            # denoiser.py:33-39, guiders.py:78-86, sampling.py:100-103

    # ---------------- resident leg ----------------
    cur = xd[T:].clone()
    for i in range(args.warmup):
        cur = step(cur, td, cd, ctrl_d, i)
    model.engine._cond_key = None     # the conditioning hoist is re-done inside the timed region (once per chunk)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        cur = step(cur, td, cd, ctrl_d, args.warmup + i)
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ops.launches() - l0
    ms = e0.elapsed_time(e1)
    finite = bool(torch.isfinite(cur).all())

    # ---------------- e2e leg: host buffers, H2D + D2H every step ----------------
    out_host = torch.empty((T, 4, h, w), dtype=torch.float32).pin_memory()
    x_host = host["x"]
    barrier()
    model.engine._cond_key = None
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    cc2 = {k: host[k].to(dev, non_blocking=True) for k in ("concat", "crossattn", "vector")}
    ctrl2 = host["ctrl"].to(dev, non_blocking=True)
    cond_bytes = sum(v.numel() * 4 for v in cc2.values()) + ctrl2.numel() * 4
    for i in range(args.steps):
        xin = x_host[T:].to(dev, non_blocking=True)
        tin = host["t"].to(dev, non_blocking=True)
        nxt = step(xin, tin, cc2, ctrl2, args.warmup + i)
        out_host.copy_(nxt, non_blocking=True)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    h2d = x_host[T:].numel() * 4 + host["t"].numel() * 4 + cond_bytes / args.steps
    d2h = out_host.numel() * 4

    # ---------------- max over ranks ----------------
    ms, ms_e2e = dist_utils.max_over_ranks([ms, ms_e2e], device=dev)

    # ---------------- per-family live profile (one extra step, outside the timed regions) ----------------
    roof = None
    fam = None
    if rank == 0:
        with ops.profile() as prof:
            step(cur, td, cd, ctrl_d, 1)
        fam = {k: dict(launches=v["launches"], ms=round(v["ms"], 3), tflops=round(v["flops"] / 1e12, 3),
                       gbytes=round(v["bytes"] / 1e9, 3)) for k, v in prof.families.items()}
        if os.environ.get("B200SVD_BENCH_SHAPES"):
            with open(os.environ["B200SVD_BENCH_SHAPES"], "w") as fh:
                for famname in ("mtgemm", "flash_attn", "small_attn", "groupnorm", "layernorm"):
                    fh.write(f"== {famname}\n")
                    for d_, n_, ms_, tf_ in ops.summarize_records(prof.launch_records, famname, 60):
                        fh.write(f"{ms_:9.3f} ms n={n_:3d} avg={ms_ / n_:7.3f} {tf_:7.1f} TF/s  {d_}\n")
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        which = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if peaks else "fallback 1.4 PFLOP/s sustained"
        g = prof.families["mtgemm"]
        ach = g["flops"] / (g["ms"] * 1e-3) / 1e12
        roof = {"kernel": "mtgemm_kernel (tcgen05 multi-tap GEMM: all Linear / Conv2d / Conv3d)", "bound": "tensor",
                "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "peak_source": which,
                "launches": g["launches"], "avg_launch_ms": g["ms"] / g["launches"],
                "share_of_step": g["ms"] / sum(v["ms"] for v in prof.families.values()),
                "traffic": None, "traffic_note": "see profiles/ for the ncu --set full capture of this kernel"}
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")))["mtgemm_conv3x3_L0_320to320"]
            roof["traffic"] = tr["traffic_bytes"]
            roof["traffic_note"] = (f"ncu --set full, one launch of the dominant conv shape ({tr['launch']}): "
                                    f"{tr['traffic_bytes'] / 1e6:.0f} MB DRAM vs {tr['algorithmic_bytes'] / 1e6:.0f} MB "
                                    f"algorithmic; roofline.achieved aggregates all {g['launches']} launches of the step")
        except Exception:
            pass

    if rank == 0:
        sps = dist_utils.aggregate_throughput(args.steps, world, ms)
        sps_e2e = dist_utils.aggregate_throughput(args.steps, world, ms_e2e)
        line = {
            "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic (random-init weights of the shipped architecture, seeded inputs)",
            "config": _workload(args), "clocks": clocks,
            "e2e": {"value": sps_e2e, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": launches, "roofline": roof, "kernel_families": fam,
            "effective_tflops_per_gpu": STEP_TFLOP / (ms / args.steps * 1e-3) / 1e0 / 1e0,
            "finite": finite,
        }
        line["effective_tflops_per_gpu"] = STEP_TFLOP / (ms / args.steps * 1e-3)
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_reference_sample()
            except Exception as exc:  # the GPU number must not be lost to a host-side problem
                line["cpu_baseline"] = {"error": repr(exc)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun on this node
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29511"),
               os.path.abspath(__file__), "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup",
               str(args.warmup)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else [])
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
