"""Noise schedule, timestep respacing and the per-step coefficient table.

Host-side, numpy fp64, once per sampler.  Everything here must be *bit-identical*
to what the reference builds because the sampler arithmetic downstream is
ill-conditioned at high t (sqrt(1/abar) ~ 2e4 at t=999).

Reference behaviour restated (not copied) from:
  diffusion/gaussian_diffusion.py:26-70    named beta schedules
  diffusion/gaussian_diffusion.py:149-186  derived fp64 tables
  diffusion/respace.py:21-74               space_timesteps (ddimN stride rule + section rule)
  diffusion/respace.py:86-100              kept-step betas: 1 - abar_i / abar_prev_kept
  diffusion/gaussian_diffusion.py:699-718  DDIM update (eta)      -> coeff columns 0..4
  diffusion/gaussian_diffusion.py:243-246,471-476  ancestral      -> coeff columns 5..7
  diffusion/gaussian_diffusion.py:1270     fp64 table entry -> .float() BEFORE any sqrt
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Iterable, List, Sequence, Set, Union

import numpy as np
import torch

BASE_STEPS = 1000

# column indices of the [N, 8] fp32 coefficient table consumed by the K3 kernel
COL_A, COL_B, COL_CX0, COL_CEPS, COL_SIGMA, COL_COEF1, COL_COEF2, COL_STD = range(8)


def _cosine_abar(t: float) -> float:
    return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2


def named_beta_schedule(name: str, n: int, scale_betas: float = 1.0) -> np.ndarray:
    """fp64 betas.  'cosine' uses python's math.cos per element (np.cos may differ by 1 ulp)."""
    if name == "linear":
        s = scale_betas * 1000 / n
        return np.linspace(s * 0.0001, s * 0.02, n, dtype=np.float64)
    if name == "cosine":
        out = np.empty(n, dtype=np.float64)
        for i in range(n):
            lo, hi = i / n, (i + 1) / n
            out[i] = min(1 - _cosine_abar(hi) / _cosine_abar(lo), 0.999)
        return out
    raise NotImplementedError(f"unknown beta schedule: {name}")


def space_timesteps(num_timesteps: int, section_counts: Union[str, Sequence[int]]) -> Set[int]:
    """Which of the base steps are kept.  'ddimN' -> first integer stride that yields N steps."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_timesteps):
                kept = range(0, num_timesteps, stride)
                if len(kept) == want:
                    return set(kept)
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(tok) for tok in section_counts.split(",")]
    n_sec = len(section_counts)
    base, extra = divmod(num_timesteps, n_sec)
    kept: List[int] = []
    start = 0
    for sec, count in enumerate(section_counts):
        size = base + (1 if sec < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return set(kept)


@dataclass
class DiffusionTables:
    """All fp64 tables the reference's GaussianDiffusion.__init__ derives from betas."""

    betas: np.ndarray

    def __post_init__(self) -> None:
        b = np.array(self.betas, dtype=np.float64)
        assert b.ndim == 1, "betas must be 1-D"
        assert (b > 0).all() and (b <= 1).all()
        self.betas = b
        self.num_timesteps = int(b.shape[0])
        alphas = 1.0 - b
        ac = np.cumprod(alphas, axis=0)
        self.alphas_cumprod = ac
        self.alphas_cumprod_prev = np.append(1.0, ac[:-1])
        self.alphas_cumprod_next = np.append(ac[1:], 0.0)
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.log_one_minus_alphas_cumprod = np.log(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)
        self.posterior_variance = b * (1.0 - self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_log_variance_clipped = np.log(
            np.append(self.posterior_variance[1], self.posterior_variance[1:])
        )
        self.posterior_mean_coef1 = b * np.sqrt(self.alphas_cumprod_prev) / (1.0 - ac)
        self.posterior_mean_coef2 = (1.0 - self.alphas_cumprod_prev) * np.sqrt(alphas) / (1.0 - ac)


def respaced_betas(base_betas: np.ndarray, use_timesteps: Iterable[int]):
    """(new_betas fp64, timestep_map) for the kept steps -- diffusion/respace.py:91-100."""
    keep = set(use_timesteps)
    base = DiffusionTables(base_betas)
    last = 1.0
    new_betas, tmap = [], []
    for i, abar in enumerate(base.alphas_cumprod):
        if i in keep:
            new_betas.append(1 - abar / last)
            last = abar
            tmap.append(i)
    return np.array(new_betas), tmap


def step_coefficients(tab: DiffusionTables, eta: float = 0.0, var_type: str = "FIXED_SMALL") -> np.ndarray:
    """[N, 8] fp32 table: one row per (respaced) step index i.

    Mirrors the reference's op ORDER in fp32 so the K3 kernel is bit-compatible:
    every table entry is cast to fp32 first (`_extract_into_tensor(...).float()`), all
    later sqrt/exp/mul happen in fp32 with torch CPU kernels (same libm as the oracle).
    """
    n = tab.num_timesteps
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float64)).float()
    a = f32(tab.sqrt_recip_alphas_cumprod)
    b = f32(tab.sqrt_recipm1_alphas_cumprod)
    ab = f32(tab.alphas_cumprod)
    abp = f32(tab.alphas_cumprod_prev)
    sigma = eta * torch.sqrt((1 - abp) / (1 - ab)) * torch.sqrt(1 - ab / abp)
    c_x0 = torch.sqrt(abp)
    c_eps = torch.sqrt(1 - abp - sigma**2)
    nonzero = torch.ones(n, dtype=torch.float32)
    nonzero[0] = 0.0
    coef1 = f32(tab.posterior_mean_coef1)
    coef2 = f32(tab.posterior_mean_coef2)
    if var_type == "FIXED_SMALL":
        logvar = f32(tab.posterior_log_variance_clipped)
    elif var_type == "FIXED_LARGE":
        logvar = f32(np.log(np.append(tab.posterior_variance[1], tab.betas[1:])))
    else:
        raise NotImplementedError(var_type)
    std_nz = nonzero * torch.exp(0.5 * logvar)
    out = torch.stack([a, b, c_x0, c_eps, nonzero * sigma, coef1, coef2, std_nz], dim=1)
    return out.contiguous().numpy()
