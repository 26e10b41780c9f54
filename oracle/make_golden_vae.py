"""ORACLE tooling — build-container only.  Golden vectors for the temporal VAE decoder from the UNMODIFIED reference
`VideoDecoder` (full-size decoder_config of code/config.yaml:242-257, small latents), and pinning of
oracle/vae_decoder_oracle.py + the parameter grammar against it.    python oracle/make_golden_vae.py"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims  # noqa: E402
from oracle import vae_decoder_oracle as vorc  # noqa: E402
from streamingt2v_b200 import arch  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = {"vae_t4_8x16": (4, 8, 16, 11), "vae_t3_16x8": (3, 16, 8, 12)}


def make_latent(T, h, w, seed):
    rng = np.random.default_rng([seed, 77])
    return torch.from_numpy((rng.normal(size=(T, 4, h, w)) * 5.0).astype(np.float32))  # z / 0.18215 scale


def main():
    ref_shims.install()
    from models.svd.sgm.modules.autoencoding.temporal_ae import VideoDecoder
    cfg = arch.VaeConfig()
    dec = VideoDecoder(attn_type="vanilla", double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3,
                       ch=128, ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0,
                       video_kernel_size=[3, 1, 1]).eval()
    ref_shapes = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    mine = arch.vae_decoder_param_shapes(cfg)
    assert ref_shapes == mine, (sorted(set(ref_shapes) ^ set(mine))[:8],
                                [(k, ref_shapes[k], mine[k]) for k in ref_shapes if k in mine and ref_shapes[k] != mine[k]][:5])
    print(f"grammar: {len(mine)} tensors, {sum(int(np.prod(s)) for s in mine.values()) / 1e6:.1f} M params OK")
    for name, (T, h, w, seed) in CASES.items():
        sd = arch.synth_state_dict(mine, seed=seed)
        dec.load_state_dict(sd, strict=True)
        z = make_latent(T, h, w, seed)
        with torch.no_grad():
            ref = dec(z.clone(), timesteps=T)
            out = vorc.decode(sd, cfg, z, T)
        err = (out - ref).abs().max().item()
        print(f"[{name}] ref absmax {ref.abs().max():.3f} std {ref.std():.3f}; oracle vs reference {err:.3e}")
        assert err <= 2e-4 * max(1.0, ref.abs().max().item())
        np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), out=ref.numpy(), meta=np.array([T, h, w, seed], np.int64),
                            oracle_vs_reference_maxerr=np.array([err]))


if __name__ == "__main__":
    main()
