"""ORACLE tooling — build-container only.  Golden vectors for the SD-VAE encoder from the UNMODIFIED reference `Encoder`
(decoder_config / encoder_config of code/config.yaml:242-257) followed by the engine's quant_conv and the posterior mode
(autoencoder.py:454-473, regulariser sample=False), and pinning of oracle/vae_encoder_oracle.py + the parameter
grammar against it.    python oracle/make_golden_vae_enc.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from oracle import vae_encoder_oracle as eorc  # noqa: E402
from streamingt2v_b200 import arch  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = {"vae_enc_2_64x96": (2, 64, 96, 21), "vae_enc_1_128x64": (1, 128, 64, 22)}


def make_image(n, H, W, seed):
    rng = np.random.default_rng([seed, 55])
    return torch.from_numpy(rng.uniform(-1, 1, size=(n, 3, H, W)).astype(np.float32))


def main():
    ref_shims.install()
    from models.svd.sgm.modules.diffusionmodules.model import Encoder
    cfg = arch.VaeConfig()
    enc = Encoder(attn_type="vanilla", double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                  ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0).eval()
    quant = torch.nn.Conv2d(8, 8, 1)                     # AutoencodingEngineLegacy.quant_conv (autoencoder.py:454-458)
    ref_shapes = {k: tuple(v.shape) for k, v in enc.state_dict().items()}
    ref_shapes.update({"quant_conv." + k: tuple(v.shape) for k, v in quant.state_dict().items()})
    mine = arch.vae_encoder_param_shapes(cfg)
    assert ref_shapes == mine, (sorted(set(ref_shapes) ^ set(mine))[:8],
                                [(k, ref_shapes[k], mine[k]) for k in ref_shapes if k in mine and ref_shapes[k] != mine[k]][:5])
    print(f"grammar: {len(mine)} tensors, {sum(int(np.prod(s)) for s in mine.values()) / 1e6:.1f} M params OK")
    for name, (n, H, W, seed) in CASES.items():
        sd = arch.synth_state_dict(mine, seed=seed)
        enc.load_state_dict({k: v for k, v in sd.items() if not k.startswith("quant_conv.")}, strict=True)
        quant.load_state_dict({k[len("quant_conv."):]: v for k, v in sd.items() if k.startswith("quant_conv.")})
        x = make_image(n, H, W, seed)
        with torch.no_grad():
            ref = quant(enc(x.clone()))[:, :4]           # DiagonalGaussianDistribution.mode() = mean
            out = eorc.encode(sd, cfg, x)
        err = (out - ref).abs().max().item()
        print(f"[{name}] ref absmax {ref.abs().max():.3f} std {ref.std():.3f}; oracle vs reference {err:.3e}")
        assert err <= 2e-4 * max(1.0, ref.abs().max().item())
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), out=ref.numpy(), meta=np.array([n, H, W, seed], np.int64),
                            oracle_vs_reference_maxerr=np.array([err]))


if __name__ == "__main__":
    main()
