"""TEST INFRASTRUCTURE: deterministic synthetic inputs shared by oracle/make_golden.py and tests/.

Inputs are regenerated from numpy RandomState seeds (platform independent), so the committed golden
fixtures only need to hold the REFERENCE's outputs.  Shapes follow the collate contract
(data_loaders/tensors.py:33-86): x [B,C,1,T], keyframes [B,ceil(T/30),104], mask [B,1,1,T] bool.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np
import torch


@dataclass(frozen=True)
class Case:
    name: str
    fmt: str
    L: int
    H: int
    B: int
    T: int
    S: int                 # audio tokens (without the 2 time tokens)
    respacing: str = "ddim10"
    guidance: float = 2.0
    seed: int = 0
    wseed: int = 1
    masked: bool = False   # knock out some keyframes through y["mask"]
    gstep: float = 1.5     # per-row guidance scale = guidance + gstep * row (exercises the per-sample scale of cfg_sampler.py:33)


CASES: Dict[str, Case] = {c.name: c for c in [
    Case("pose_small", "pose", 2, 8, 2, 60, 198, masked=True),
    Case("pose_small_h4", "pose", 1, 4, 1, 45, 150, seed=3, wseed=4),
    Case("face_small", "face", 2, 8, 2, 64, 211, guidance=10.0, seed=5, wseed=6),
    Case("pose_full", "pose", 6, 8, 1, 600, 1998, seed=7, wseed=8),
    Case("face_cfg1", "face", 8, 8, 1, 64, 211, guidance=10.0, seed=9, wseed=10),   # BASELINE config 1 geometry
    Case("face_full", "face", 8, 8, 1, 600, 1998, guidance=10.0, seed=11, wseed=12),
    # smallest geometry that takes the fused row-chain arm (T >= 128, T % 8 == 0): __graft_entry__.smoke()
    Case("pose_smoke", "pose", 2, 8, 2, 160, 398, seed=15, wseed=16, masked=True),
    # the BENCHMARKED configuration (BASELINE configs[1]: pose, T=600, all 1000 steps, CFG) at a batch that takes the
    # loop's row-group cut (4 concurrent forwards, batch-row offsets b0 > 0)
    Case("pose_full_b4", "pose", 6, 8, 4, 600, 1998, respacing="", seed=13, wseed=14),      # guidance 2.0, 3.5, 5.0, 6.5 per row
    # the same with the benchmark's constant guidance 2.0 on every row (sample/generate.py:128-130 applies ONE guidance_param)
    Case("pose_full_b4_g2", "pose", 6, 8, 4, 600, 1998, respacing="", seed=13, wseed=14, gstep=0.0),
]}


def dims_of(case: Case):
    from audio2photoreal_b200.weights import model_dims
    return model_dims(case.fmt, case.L, case.H)


def make_inputs(case: Case, n_noise: int = 0) -> Dict[str, torch.Tensor]:
    rs = np.random.RandomState(1000 + case.seed)
    C = 104 if case.fmt == "pose" else 256
    cd = 1024 if case.fmt == "pose" else 2038
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32))
    out = {
        "x": f32(rs.standard_normal((case.B, C, 1, case.T))),
        "feats": f32(rs.standard_normal((case.B, case.S, cd))),
        "scale": f32(case.guidance + case.gstep * np.arange(case.B)),
        "times": torch.from_numpy(rs.randint(0, 1000, size=(case.B,)).astype(np.int64)),
    }
    nk = len(range(0, case.T, 30))
    out["keyframes"] = f32(rs.standard_normal((case.B, nk, 104)))
    mask = np.ones((case.B, 1, 1, case.T), dtype=bool)
    if case.masked:
        mask[0, ..., 30:] = False
    out["mask"] = torch.from_numpy(mask)
    if n_noise:
        out["noise_tape"] = [f32(rs.standard_normal((case.B, C, 1, case.T))) for _ in range(n_noise)]
    return out


def weights_of(case: Case):
    from audio2photoreal_b200.weights import synthetic_state_dict
    return synthetic_state_dict(dims_of(case), seed=case.wseed)


def layer_inputs(case, seed=11):
    """fixed (x, mem, t, mem2) for ONE decoder layer; regenerated from the seed by the tests"""
    D = dims_of(case).D
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(case.B, case.T, D, generator=g)
    mem = torch.randn(case.B, min(case.S, 64) + 2, D, generator=g)
    t = torch.randn(case.B, D, generator=g)
    mem2 = torch.randn(case.B, 20, D, generator=g) if case.fmt == "pose" else None
    return x, mem, t, mem2


def variant_inputs(kind: str):
    """inputs of the sampler-keyword variants case (oracle/make_golden.py golden_loop_variants): pose_small with
    clip_denoised=True + skip_timesteps=3 + a random init_image (DDIM) / const_noise + clip + skip_timesteps=2 (ancestral)"""
    case = CASES["pose_small"]
    resp, skip = ("ddim10", 3) if kind == "ddim" else ("10", 2)
    n = 10 - skip
    inp, sd = make_inputs(case, n_noise=n), weights_of(case)
    init = 0.5 * torch.from_numpy(np.random.RandomState(77).standard_normal(tuple(inp["x"].shape)).astype(np.float32))
    return case, resp, inp, sd, skip, init
