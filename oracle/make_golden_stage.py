"""ORACLE tooling — build-container only (needs /root/reference).  Pins row a22 (the stage driver) against the
UNMODIFIED reference: the methods of `StreamingSVD` (code/diffusion_trainer/streaming_svd.py)

    decode_first_stage :124-151, _generate_conditional_output :155-221, extract_anchor_frames :224-261,
    extract_ctrl_frames :263-290, _autoregressive_generation :293-356, get_batch_sgm :86-120

are imported from the reference and executed, unbound, on a mock `self` on the CPU, with the deterministic
stand-ins of oracle/stage_stubs.py for the heavy components (sampler, denoiser network, VAE decoder, conditioner).
The class cannot be constructed here (Lightning, diffusers, OpenCLIP, checkpoints), but these methods touch only
`self.<component>` attributes, and with `self.device = "cpu"` none of their hard-coded "cuda" lines is reached
(`use_memopt` False).  Missing third-party modules are stubbed at import time only.  Output:
tests/golden/stage_reference.npz = the video the reference assembles (uint8, its IImage container's array) plus the
log of what reached the network and the decoder; tests/test_stage.py replays the same scenario through
streamingt2v_b200/stage.py and oracle/stage_oracle.py.

    python oracle/make_golden_stage.py
"""
from __future__ import annotations

import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shims, stage_stubs as st  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
_STUB = ("jsonargparse", "gdown", "diffusers", "omegaconf", "pytorch_lightning", "kornia", "open_clip", "matplotlib",
         "imageio", "timm", "xformers")


class _Any(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = type(name, (), {"__init__": lambda self, *a, **k: None})
        setattr(self, name, obj)
        return obj


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Empty stand-in modules for third-party packages that are absent offline (import-time only)."""

    def __init__(self):
        self.missing = set()
        for r in _STUB:
            try:
                if importlib.util.find_spec(r) is None:
                    self.missing.add(r)
            except Exception:
                self.missing.add(r)

    def find_spec(self, name, path, target=None):
        if name.split(".")[0] in self.missing:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _Any(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, m):
        pass


def import_reference_stage():
    sys.meta_path.append(_Finder())
    sys.path.insert(0, ref_shims.REFERENCE_CODE)
    import pytorch_lightning
    pytorch_lightning.LightningModule = nn.Module
    from diffusion_trainer.streaming_svd import StreamingSVD
    from models.svd.sgm.modules.autoencoding.temporal_ae import VideoDecoder
    return StreamingSVD, VideoDecoder


def main():
    StreamingSVD, VideoDecoder = import_reference_stage()

    class _Dec(VideoDecoder):                       # isinstance(..., VideoDecoder) selects the timesteps kwarg (:139-142)
        def __init__(self):
            nn.Module.__init__(self)

    network, decoder = st.StubNetwork(), st.StubDecoder()
    me = types.SimpleNamespace()
    for name in ("decode_first_stage", "_generate_conditional_output", "extract_anchor_frames", "extract_ctrl_frames",
                 "_autoregressive_generation", "get_batch_sgm", "get_unique_embedder_keys_from_conditioner"):
        setattr(me, name, types.MethodType(getattr(StreamingSVD, name), me))
    params = types.SimpleNamespace(num_conditional_frames=st.NCOND, anchor_frames=str(st.ANCHOR),
                                   n_autoregressive_generations=3)
    me.inference_params = params
    me.diff_trainer_params = types.SimpleNamespace(scale_factor=0.18215, disable_first_stage_autocast=True)
    me.use_memopt = False
    me.device = torch.device("cpu")
    me.sampler = st.RefStyleSampler()
    me.conditioner = st.RefStyleConditioner()
    me.denoiser = st.stub_denoiser
    me.inference_model = network
    me.first_stage_model = types.SimpleNamespace(decoder=_Dec(), decode=decoder.decode)

    first = st.first_chunk()
    torch.manual_seed(1234)                          # the reference draws cond_aug noise and randn from the global RNG
    with torch.no_grad():
        result = me._autoregressive_generation(first.clone(), params)
    data = np.asarray(result.data)                   # IImage: uint8 [F, H, W, C]
    print("reference video:", data.shape, data.dtype, "mean", data.mean())
    calls = network.calls
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(
        os.path.join(GOLDEN, "stage_reference.npz"), video_u8=data,
        decode_sizes=np.array(decoder.sizes, np.int64),
        net_calls=np.array([[c["n"], c["bs"], c["nvf"], c["ncf"], c["ioi"][0], c["ioi"][1]] for c in calls], np.int64),
        ctrl_shape=np.array(calls[0]["ctrl_shape"], np.int64),
        ctrl_sums=np.array([c["ctrl_sum"] for c in calls], np.float64),
        vec_sums=np.array([c["vec_sum"] for c in calls], np.float64),
        meta=np.array([st.NUM_FRAMES, st.NCOND, st.ANCHOR, 3, 1234], np.int64))
    print("wrote tests/golden/stage_reference.npz;", len(calls), "network calls, decode groups", decoder.sizes)


if __name__ == "__main__":
    main()
