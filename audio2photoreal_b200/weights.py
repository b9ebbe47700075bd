"""Checkpoint contract of the denoiser: parameter names / shapes, and a deterministic synthetic init.

The names and shapes are the reference FiLMTransformer's `state_dict()` (model/diffusion.py:83-199,
model/modules/transformer_modules.py:127-176; dump in SURVEY.md appendix A.3) so that
`utils.model_util.load_model` (utils/model_util.py:30-38) accepts/produces the same checkpoints.
Frozen `audio_model.*` / `lip_model.*` entries (fairseq modules) are outside the replaced path and are
carried opaquely when present.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Tuple

import numpy as np
import torch

EMB_LEN = 1998  # model/diffusion.py:136 (hard-coded audio-token count of a 600-frame window)
FF_SIZE = 1024  # utils/model_util.py:64
KEYFRAME_STEP = 30  # model/diffusion.py:147
TCN_DILATIONS = (1, 2, 3, 1, 2, 3)  # model/diffusion.py:201-212


@dataclass(frozen=True)
class ModelDims:
    fmt: str  # "pose" | "face"
    C: int  # nfeats
    D: int  # latent
    L: int
    H: int
    FF: int
    cond_dim: int
    S2: int  # keyframe tokens (pose) or 0

    @property
    def dh(self) -> int:
        return self.D // self.H

    @property
    def n_film(self) -> int:
        return 4 if self.fmt == "pose" else 3


def model_dims(fmt: str, layers: int, heads: int, max_seq_length: int = 600) -> ModelDims:
    if fmt == "pose":
        s2 = len(range(0, max_seq_length, KEYFRAME_STEP))
        return ModelDims("pose", 104, 256, layers, heads, FF_SIZE, 1024, s2)
    if fmt == "face":
        return ModelDims("face", 256, 512, layers, heads, FF_SIZE, 1024 + 1014, 0)
    raise ValueError(fmt)


def _attn(prefix: str, D: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    return [
        (f"{prefix}.in_proj_weight", (3 * D, D), "w"),
        (f"{prefix}.in_proj_bias", (3 * D,), "b"),
        (f"{prefix}.out_proj.weight", (D, D), "w"),
        (f"{prefix}.out_proj.bias", (D,), "b"),
    ]


def _lin(prefix: str, o: int, i: int):
    return [(f"{prefix}.weight", (o, i), "w"), (f"{prefix}.bias", (o,), "b")]


def _ln(prefix: str, D: int):
    return [(f"{prefix}.weight", (D,), "g"), (f"{prefix}.bias", (D,), "b")]


def denoiser_param_spec(d: ModelDims) -> List[Tuple[str, Tuple[int, ...], str]]:
    """Ordered (name, shape, kind); kind in w(eight) b(ias) g(ain) e(mbedding) f(req buffer) k(resample kernel)."""
    D, C = d.D, d.C
    spec: List[Tuple[str, Tuple[int, ...], str]] = [
        ("null_cond_embed", (1, EMB_LEN, D), "e"),
        ("null_cond_hidden", (1, D), "e"),
    ]
    if d.fmt == "pose":
        spec.append(("null_pose_embed", (1, d.S2, D), "e"))
    spec.append(("rotary.freqs", (D // 2,), "f"))
    spec += _lin("time_mlp.1", 4 * D, D) + _lin("to_time_cond.0", D, 4 * D) + _lin("to_time_tokens.0", 2 * D, 4 * D)
    spec += _ln("norm_cond", D)
    spec.append(("audio_resampler.kernel", (1, 1, 41), "k"))
    spec += _lin("input_projection", D, C)
    if d.fmt == "pose":
        spec += _lin("frame_cond_projection", D, 104) + _ln("frame_norm_cond", D)
        chans = [(max(256, C), C), (C, max(256, C)), (C, C), (C, C), (C, C), (C, C)]
        for i, (co, ci) in enumerate(chans):
            spec += [(f"post_pose_layers.{i}.weight", (co, ci, 3), "w"), (f"post_pose_layers.{i}.bias", (co,), "b")]
        spec += [("final_conv.weight", (C, C, 1), "w"), ("final_conv.bias", (C,), "b")]
    else:
        for i in range(2):
            p = f"cond_encoder.{i}"
            spec += _attn(f"{p}.self_attn", D) + _lin(f"{p}.linear1", d.FF, D) + _lin(f"{p}.linear2", D, d.FF)
            spec += _ln(f"{p}.norm1", D) + _ln(f"{p}.norm2", D) + [(f"{p}.rotary.freqs", (D // 2,), "f")]
    spec += _lin("cond_projection", D, d.cond_dim)
    spec += _ln("non_attn_cond_projection.0", D) + _lin("non_attn_cond_projection.1", D, D)
    spec += _lin("non_attn_cond_projection.3", D, D)
    for n in range(d.L):
        p = f"seqTransDecoder.stack.{n}"
        spec += _attn(f"{p}.self_attn", D) + _attn(f"{p}.multihead_attn", D)
        spec += _lin(f"{p}.linear1", d.FF, D) + _lin(f"{p}.linear2", D, d.FF)
        spec += _ln(f"{p}.norm1", D) + _ln(f"{p}.norm2", D) + _ln(f"{p}.norm3", D)
        spec += _lin(f"{p}.film1.block.1", 2 * D, D) + _lin(f"{p}.film2.block.1", 2 * D, D)
        spec += _lin(f"{p}.film3.block.1", 2 * D, D)
        if d.fmt == "pose":
            spec += _attn(f"{p}.multihead_attn2", D) + _ln(f"{p}.norm2a", D) + _lin(f"{p}.film2a.block.1", 2 * D, D)
        spec.append((f"{p}.rotary.freqs", (D // 2,), "f"))
    spec += _lin("final_layer", C, D)
    return spec


def rotary_freqs(D: int) -> torch.Tensor:
    """theta_i = 10000^(-2i/D), i < D/2 -- model/modules/rotary_embedding_torch.py:99-101 (fp32 ops)."""
    return 1.0 / (10000 ** (torch.arange(0, D, 2)[: D // 2].float() / D))


def synthetic_state_dict(d: ModelDims, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic (numpy RandomState) weights: fan-in scaled matrices, NON-zero biases and LN affine
    so every bias/affine code path is exercised by the parity tests.  Platform independent."""
    rs = np.random.RandomState(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape, kind in denoiser_param_spec(d):
        if kind == "w":
            fan_in = int(np.prod(shape[1:]))
            a = rs.standard_normal(shape) / np.sqrt(fan_in)
        elif kind == "b":
            a = 0.05 * rs.standard_normal(shape)
        elif kind == "g":
            a = 1.0 + 0.1 * rs.standard_normal(shape)
        elif kind == "e":
            a = rs.standard_normal(shape)
        elif kind == "f":
            out[name] = rotary_freqs(d.D)
            continue
        elif kind == "k":
            a = rs.standard_normal(shape) / 41.0
        out[name] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    return out
