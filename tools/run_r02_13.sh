set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vae_gpu.py -m gpu -q -s > gpurun_out/r02_gputest_vae_enc.log 2>&1; echo "pytest exit $?"; grep -E "REFERENCE golden|passed|failed" gpurun_out/r02_gputest_vae_enc.log | cut -c1-200
python - <<'PY'
import torch, json
from streamingt2v_b200 import arch, ops
from streamingt2v_b200.vae import B200VaeEncoder
dev = torch.device("cuda:0")
cfg = arch.VaeConfig()
enc = B200VaeEncoder(cfg, arch.synth_state_dict_device(arch.vae_encoder_param_shapes(cfg), dev, 3), dev)
x = torch.rand(1, 3, 576, 1024, device=dev) * 2 - 1
ms = []
for it in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); z = enc.encode(x); e1.record(); torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
print("VAE encode 1 frame 576x1024:", ms, "ms; finite", bool(torch.isfinite(z).all()), tuple(z.shape), f"{2.61 / (min(ms) * 1e-3):.0f} TFLOP/s")
json.dump({"ms": ms, "frames": 1, "tflop_per_frame": 2.61}, open("gpurun_out/r02_run_vae_enc.json", "w"))
PY
