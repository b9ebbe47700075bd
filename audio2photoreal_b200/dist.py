"""Multi-GPU: the batch-of-samples x sequence-windows axis shards across ranks (one process per GPU,
torchrun); every row is an independent trajectory, so there is NO per-step communication and exactly
one collective: an all-gather of the final motion codes (NCCL over NVLink on GPUs, gloo in CPU tests).

Determinism: every rank seeds identically, draws the GLOBAL initial noise (and, for the stochastic samplers
-- ancestral / eta > 0 -- the GLOBAL per-step noise tape) and slices its own rows, or uses the in-kernel Philox
stream keyed by (seed, GLOBAL row index): the rows a rank computes do not depend on the partition.  On the exact
fp32 arm (split_terms = 0) the W-rank result is bit-identical to the 1-rank result; on the tensor-core arms it is
identical up to fp32 summation order (the attention's split-KV tail cuts keys by launch size), which
tests/test_gpu_parity.py::test_batch_rows_independent_and_deterministic bounds.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_rows: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row block of `rank`; the first n_rows % world ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_y(y: Dict, lo: int, hi: int, n_rows: int) -> Dict:
    """Slice every per-sample tensor of the reference's y dict (data_loaders/tensors.py:33-86)."""
    out = {}
    for k, v in y.items():
        out[k] = v[lo:hi] if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == n_rows else v
    return out


def global_noise(shape, seed: int, device) -> torch.Tensor:
    """The same [B,C,1,T] tensor on every rank (generator seeded identically; drawn on `device`)."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return torch.randn(*shape, device=device, generator=g)


def all_gather_rows(local: torch.Tensor, n_rows: int, group=None) -> torch.Tensor:
    """Single all-gather of the per-rank result rows -> [n_rows, ...] on every rank (ragged-safe)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_rows, world, r) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def global_noise_tape(n_steps: int, shape, seed: int, device) -> torch.Tensor:
    """The same [n_steps, B, C, 1, T] per-step noise on every rank (ancestral / eta > 0 samplers); ranks slice [:, lo:hi]."""
    g = torch.Generator(device=device)
    g.manual_seed(seed + 0x5EED)
    return torch.randn(n_steps, *shape, device=device, generator=g)


def sample_sharded(local_loop: Callable[..., torch.Tensor], shape, y: Dict, seed: int, device, group=None,
                   noise: Optional[torch.Tensor] = None, noise_tape_steps: int = 0) -> torch.Tensor:
    """Run `local_loop(local_shape, local_noise, local_y[, lo, hi[, local_tape]])` on this rank's rows and all-gather.

    `local_loop` is e.g. lambda s, n, yy, lo, hi: sampler.ddim_sample_loop(model, s, noise=n, clip_denoised=False,
    model_kwargs={"y": yy}).  For the stochastic samplers either pass `noise_tape_steps = n` (the global tape is drawn
    with the seeded generator and this rank's slice arrives as the 6th argument -> `noise_tape=`), or forward
    `row0=lo` with `noise_rng="philox"`; NEVER let every rank draw its own torch.randn_like tape from an identically
    seeded generator (all ranks would apply the same noise to different rows).  A 3-argument callable (deterministic
    eta = 0 DDIM) keeps working.  Works unchanged with world size 1 / no process group."""
    import inspect
    B = shape[0]
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank(group) if world > 1 else 0
    lo, hi = shard_range(B, world, rank)
    if noise is None:
        noise = global_noise(shape, seed, device)
    args = [(hi - lo,) + tuple(shape[1:]), noise[lo:hi].contiguous(), shard_y(y, lo, hi, B)]
    try:
        n_params = len([p for p in inspect.signature(local_loop).parameters.values()
                        if p.kind in (p.POSITIONAL_ONLY, p.POSITIONAL_OR_KEYWORD)])
    except (TypeError, ValueError):
        n_params = 3
    if n_params >= 5:
        args += [lo, hi]
    if noise_tape_steps:
        if n_params < 6:
            raise TypeError("noise_tape_steps needs local_loop(shape, noise, y, lo, hi, tape)")
        args.append(global_noise_tape(noise_tape_steps, shape, seed, device)[:, lo:hi].contiguous())
    local = local_loop(*args)
    return all_gather_rows(local, B, group)
