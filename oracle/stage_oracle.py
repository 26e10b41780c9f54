"""ORACLE — test infrastructure only.  Plain restatement of the frame bookkeeping of the reference stage driver
(code/diffusion_trainer/streaming_svd.py:124-151 decode_first_stage, :155-221 _generate_conditional_output,
:263-290 extract_ctrl_frames, :293-356 _autoregressive_generation; utils/result_processor.py:4-14 convert_range)
with the heavy components passed in as callables, written independently of streamingt2v_b200/stage.py (explicit
python loops over frames and chunks).  Pinned against the reference: oracle/make_golden_stage.py executes the UNMODIFIED reference methods (unbound, on a
mock `self`, CPU) with the stand-in components of oracle/stage_stubs.py and stores what they produce in
tests/golden/stage_reference.npz; tests/test_stage.py replays that scenario through this file and through
streamingt2v_b200/stage.py."""
from __future__ import annotations

import math

import torch


def decode_first_stage(decode, z, scale_factor=0.18215, max_chunk=8):
    z = z / scale_factor
    n = z.shape[0]
    per = min(n, max_chunk)
    out = []
    for r in range(math.ceil(n / per)):
        part = z[r * per:min(n, (r + 1) * per)]
        out.append(decode(part, part.shape[0]))
    return torch.cat(out, 0)


def autoregressive_generation(first_chunk, n_gen, *, conditioner, sample, decode, num_frames, n_cond=7, anchor=0,
                              noise):
    """first_chunk [F,C,H,W] in [-1,1]; sample(noise, c, uc, ctrl_frames) -> latents [T,4,h,w];
    noise(i) -> the i-th chunk's initial noise.  Returns (video [0,255], list of ctrl_frames seen)."""
    chunks = [first_chunk]
    seen_ctrl = []
    for i in range(n_gen):
        last = chunks[-1]
        ctrl = torch.stack([last[last.shape[0] - n_cond + j] for j in range(n_cond)])[None]
        seen_ctrl.append(ctrl)
        frame = chunks[0][anchor]
        c, uc = conditioner(frame, num_frames)
        c = {k: (torch.cat([v] * num_frames, 0) if k in ("crossattn", "concat") else v) for k, v in c.items()}
        uc = {k: (torch.cat([v] * num_frames, 0) if k in ("crossattn", "concat") else v) for k, v in uc.items()}
        z = sample(noise(i), c, uc, ctrl)
        x = decode_first_stage(decode, z).clamp(-1.0, 1.0)
        chunks.append(x[n_cond:])
    video = torch.cat([(ch + 1.0) / 2.0 * 255.0 for ch in chunks], 0)
    return video, seen_ctrl
