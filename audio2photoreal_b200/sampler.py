"""Sampler -- drop-in for the reference's `SpacedDiffusion` sampling API
(diffusion/respace.py:77-145, diffusion/gaussian_diffusion.py:525-665,815-936).

`ddim_sample_loop` / `p_sample_loop` keep the reference's full keyword sets.  When the model is an
a2p_b200 `CFGDenoiser` / `Denoiser` the whole reverse loop runs inside the CUDA library
(a2p_sample_loop: one CUDA graph per diffusion step, per-step scalars read on the device); for any
other callable the model is evaluated by the caller's PyTorch code and only the fused K3 epilogue
(a2p_sampler_step) runs here.  There is no CPU implementation.

Known reference behaviour kept on purpose:
  * `p_sample` is broken as shipped (undefined `noise`, gaussian_diffusion.py:476); this class
    implements the upstream-MDM form (noise = randn_like(x); const_noise -> noise[[0]] repeated).
  * `ddim_sample_loop` returns the LAST pred_xstart, not the last sample (:862).
  * DDIM draws randn_like(x) every step even at eta = 0 (:708), advancing the generator; with
    `advance_rng=True` (default) the same number of draws is made so later sampling stays in step.
"""
from __future__ import annotations

import ctypes as C
import enum
from typing import Iterable, Optional

import numpy as np
import torch

from . import _lib
from .denoiser import CFGDenoiser, Denoiser
from .schedule import (DiffusionTables, named_beta_schedule, respaced_betas, space_timesteps, step_coefficients)


class ModelMeanType(enum.Enum):
    PREVIOUS_X = enum.auto()
    START_X = enum.auto()
    EPSILON = enum.auto()


class ModelVarType(enum.Enum):
    LEARNED = enum.auto()
    FIXED_SMALL = enum.auto()
    FIXED_LARGE = enum.auto()
    LEARNED_RANGE = enum.auto()


class LossType(enum.Enum):
    MSE = enum.auto()
    RESCALED_MSE = enum.auto()
    KL = enum.auto()
    RESCALED_KL = enum.auto()


DDIM, ANCESTRAL = 0, 1


class Sampler(DiffusionTables):
    def __init__(self, use_timesteps: Iterable[int], *, betas, model_mean_type=ModelMeanType.START_X,
                 model_var_type=ModelVarType.FIXED_SMALL, loss_type=LossType.MSE, rescale_timesteps: bool = False,
                 lambda_vel: float = 0.0, data_format: str = "pose", model_path=None):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(betas)
        new_betas, self.timestep_map = respaced_betas(np.asarray(betas, dtype=np.float64), self.use_timesteps)
        DiffusionTables.__init__(self, new_betas)
        # enums may come from the reference's module when patched in: compare by name
        self.model_mean_type, self.model_var_type, self.loss_type = model_mean_type, model_var_type, loss_type
        if getattr(model_mean_type, "name", model_mean_type) != "START_X":
            raise NotImplementedError("the reference hard-codes x0-prediction (utils/model_util.py:80,97-99)")
        if getattr(model_var_type, "name", "") not in ("FIXED_SMALL", "FIXED_LARGE"):
            raise NotImplementedError("learned variances are not used by the reference's sampler configuration")
        if rescale_timesteps:
            raise NotImplementedError("rescale_timesteps is hard-coded False in the reference (utils/model_util.py:85)")
        self.rescale_timesteps = False
        self.lambda_vel, self.data_format, self.model_path = lambda_vel, data_format, model_path
        self._dev_tables = {}

    # ------------------------------------------------------------------ helpers
    def _tables_on(self, dev, eta: float):
        key = (dev, float(eta))
        if key not in self._dev_tables:
            co = step_coefficients(self, eta=eta, var_type=self.model_var_type.name)
            self._dev_tables[key] = (torch.from_numpy(co).to(dev), torch.tensor(self.timestep_map, dtype=torch.int64, device=dev))
        return self._dev_tables[key]

    @staticmethod
    def _device_of(model, device):
        if device is not None:
            return torch.device(device)
        return next(model.parameters()).device

    def _init_image(self, shape, noise, device, skip_timesteps, init_image):
        img = noise if noise is not None else torch.randn(*shape, device=device)
        if skip_timesteps and init_image is None:
            init_image = torch.zeros_like(img)
        n = self.num_timesteps - skip_timesteps
        if init_image is not None:
            # q_sample at the first index (gaussian_diffusion.py:215-233,628-632)
            i0 = n - 1
            a = torch.tensor(self.sqrt_alphas_cumprod[i0]).float().to(device)
            b = torch.tensor(self.sqrt_one_minus_alphas_cumprod[i0]).float().to(device)
            img = a * init_image + b * img
        return img.to(device=device, dtype=torch.float32).contiguous(), n

    def _unsupported(self, denoised_fn, cond_fn, randomize_class, cond_fn_with_grad):
        if cond_fn is not None or cond_fn_with_grad:
            raise NotImplementedError("classifier guidance (cond_fn) is not on the sampling path of the reference CLI")
        if denoised_fn is not None:
            raise NotImplementedError("denoised_fn is unused by the reference callers")
        if randomize_class:
            raise NotImplementedError("randomize_class is unused by the reference callers")

    # ------------------------------------------------------------------ the loop
    def _loop(self, kind, model, shape, noise, clip_denoised, model_kwargs, device, progress, eta, skip_timesteps,
              init_image, const_noise, noise_tape, advance_rng, use_graph, noise_rng="torch", seed=None, row0=0):
        assert isinstance(shape, (tuple, list))
        dev = self._device_of(model, device)
        if dev.type != "cuda":
            raise _lib.A2PError("a2p_b200 samplers run on CUDA only (no CPU fallback)")
        B, Cc, one, T = shape
        model_kwargs = model_kwargs or {}
        with torch.cuda.device(dev):
            x, n = self._init_image(shape, noise, dev, skip_timesteps, init_image)
            if noise is not None and x.data_ptr() == noise.data_ptr():
                x = x.clone()   # never sample in place on the caller's noise tensor
            coeffs, tsmap = self._tables_on(dev, eta)
            need_noise = kind == ANCESTRAL or eta != 0.0
            tape = None
            philox = need_noise and noise_rng == "philox"
            if noise_rng not in ("torch", "philox"):
                raise ValueError("noise_rng must be 'torch' (bit-parity tape) or 'philox' (in-kernel, statistically equivalent)")
            if philox:
                if const_noise or noise_tape is not None:
                    raise ValueError("noise_rng='philox' draws inside the kernel: const_noise / noise_tape need noise_rng='torch'")
                if seed is None:
                    seed = int(torch.randint(0, 2 ** 62, (1,)).item())
            if need_noise and not philox:
                if noise_tape is not None:
                    tape = noise_tape.to(dev, torch.float32).contiguous()
                    assert tape.shape == (n, B, Cc, one, T), "noise_tape must be [n_steps, B, C, 1, T]"
                else:
                    # same draw order as the reference: one randn_like(x) per step, first draw = step n-1
                    draws = [torch.randn_like(x) for _ in range(n)]
                    tape = torch.stack(draws, 0)
                if const_noise:
                    tape = tape[:, :1].expand(-1, B, -1, -1, -1).contiguous()
            elif advance_rng and noise_tape is None:
                for _ in range(n):
                    torch.randn_like(x)   # keeps the CUDA generator in step with gaussian_diffusion.py:708
            pred = torch.empty_like(x)
            lib = _lib.load()
            st = torch.cuda.current_stream(dev).cuda_stream
            inner = model.model if isinstance(model, CFGDenoiser) else model
            if isinstance(inner, Denoiser):
                y = model_kwargs.get("y", {})
                inner._ensure_bound(dev, max(T, 2000))
                inner.prepare(y, B, T, dev)
                cfg = isinstance(model, CFGDenoiser)
                scale = y["scale"].to(dev, torch.float32).contiguous() if cfg else None
                ws = inner._workspace(lib.a2p_workspace_bytes(C.byref(inner._cfg), B, T), dev)
                # sub-tables for skip_timesteps: loop runs indices n-1 .. 0
                if philox:
                    _lib.check(lib.a2p_sample_loop_rng(
                        inner._handle, kind, B, T, n, coeffs.data_ptr(), tsmap.data_ptr(),
                        scale.data_ptr() if scale is not None else None, x.data_ptr(), pred.data_ptr(), int(seed), int(row0),
                        int(bool(clip_denoised)), 3 if cfg else 1, int(bool(use_graph)), ws.data_ptr(), ws.numel(), st))
                else:
                    _lib.check(lib.a2p_sample_loop(
                        inner._handle, kind, B, T, n, coeffs.data_ptr(), tsmap.data_ptr(),
                        scale.data_ptr() if scale is not None else None, x.data_ptr(), pred.data_ptr(),
                        tape.data_ptr() if tape is not None else None, int(bool(clip_denoised)), 3 if cfg else 1,
                        int(bool(use_graph)), ws.data_ptr(), ws.numel(), st))
                self._last_keep = (x, pred, tape, scale)
                return x, pred
            # generic model: caller's PyTorch forward + fused K3 epilogue
            idx = range(n - 1, -1, -1)
            if progress:
                from tqdm.auto import tqdm
                idx = tqdm(idx)
            for k, i in enumerate(idx):
                ts = torch.full((B,), self.timestep_map[i], device=dev, dtype=torch.int64)
                with torch.no_grad():
                    out = model(x, ts, **model_kwargs).float().contiguous()
                if philox:
                    _lib.check(lib.a2p_sampler_step_rng(
                        kind, B, Cc, T, x.data_ptr(), out.data_ptr(), None, None, coeffs[i].data_ptr(), int(seed), k, int(row0),
                        int(bool(clip_denoised)), x.data_ptr(), pred.data_ptr(), st))
                else:
                    _lib.check(lib.a2p_sampler_step(
                        kind, B, Cc, T, x.data_ptr(), out.data_ptr(), None, None, coeffs[i].data_ptr(),
                        tape[k].data_ptr() if tape is not None else None, int(bool(clip_denoised)), x.data_ptr(),
                        pred.data_ptr(), st))
            return x, pred

    def ddim_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                         model_kwargs=None, device=None, progress=False, eta=0.0, skip_timesteps=0, init_image=None,
                         randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False,
                         noise_tape=None, advance_rng=True, use_graph=True, noise_rng="torch", seed=None, row0=0):
        """Returns the last pred_xstart [B,C,1,T] (gaussian_diffusion.py:862).  noise_rng='philox' (eta > 0 only): the
        per-step noise is drawn inside the kernel (statistically equivalent to th.randn_like, no tape)."""
        if dump_steps is not None:
            raise NotImplementedError()
        if const_noise == True:  # noqa: E712  (same check as the reference, :841)
            raise NotImplementedError()
        self._unsupported(denoised_fn, cond_fn, randomize_class, cond_fn_with_grad)
        _, pred = self._loop(DDIM, model, shape, noise, clip_denoised, model_kwargs, device, progress, eta,
                             skip_timesteps, init_image, False, noise_tape, advance_rng, use_graph, noise_rng, seed, row0)
        return pred

    def p_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None,
                      model_kwargs=None, device=None, progress=False, skip_timesteps=0, init_image=None,
                      randomize_class=False, cond_fn_with_grad=False, dump_steps=None, const_noise=False,
                      noise_tape=None, use_graph=True, noise_rng="torch", seed=None, row0=0):
        """Ancestral sampling; returns the final sample [B,C,1,T] (gaussian_diffusion.py:590).  noise_rng='torch' (default)
        reproduces the reference's randn_like draws bit for bit through a tape; 'philox' draws the noise inside the
        sampler kernel (Philox4x32-10 keyed by `seed`; `row0` = global index of the first batch row when sharded)."""
        if dump_steps is not None:
            raise NotImplementedError("dump_steps needs per-step host copies; not on the reference callers' path")
        self._unsupported(denoised_fn, cond_fn, randomize_class, cond_fn_with_grad)
        x, _ = self._loop(ANCESTRAL, model, shape, noise, clip_denoised, model_kwargs, device, progress, 0.0,
                          skip_timesteps, init_image, const_noise, noise_tape, False, use_graph, noise_rng, seed, row0)
        return x


    # ------------------------------------------------------------------ PLMS (gaussian_diffusion.py:938-1158)
    def _ext(self, arr, i, like):
        """_extract_into_tensor: fp64 table entry -> fp32 scalar tensor on the device (gaussian_diffusion.py:1260-1273)"""
        return torch.tensor(float(arr[i]), dtype=torch.float64).float().to(like.device)

    @torch.no_grad()
    def plms_sample_loop(self, model, shape, noise=None, clip_denoised=True, denoised_fn=None, cond_fn=None, model_kwargs=None,
                         device=None, progress=False, skip_timesteps=0, init_image=None, randomize_class=False,
                         cond_fn_with_grad=False, order=2):
        """Pseudo linear multistep sampling: the reference's third sampler (unused by its CLI).  The denoiser evaluations run in
        the CUDA library (one forward per call: this sampler needs x0 at TWO points in its first step and an eps history, so it
        is driven from the host); the multistep arithmetic is a handful of elementwise fp32 ops in the reference's order.
        Returns the final sample [B,C,1,T] (:1054)."""
        self._unsupported(denoised_fn, cond_fn, randomize_class, cond_fn_with_grad)
        if not int(order) or not 1 <= order <= 4:
            raise ValueError("order is invalid (should be int from 1-4).")
        dev = self._device_of(model, device)
        if dev.type != "cuda":
            raise _lib.A2PError("a2p_b200 samplers run on CUDA only (no CPU fallback)")
        model_kwargs = model_kwargs or {}
        B = shape[0]
        with torch.cuda.device(dev):
            img, n = self._init_image(shape, noise, dev, skip_timesteps, init_image)

            def model_out(x, i):
                ts = torch.full((B,), self.timestep_map[i], device=dev, dtype=torch.int64)      # _WrappedModel (respace.py:140-145)
                out = model(x, ts, **model_kwargs).float()
                if clip_denoised:
                    out = out.clamp(-1, 1)
                x0 = out.permute(0, 2, 1).unsqueeze(2)
                eps = (self._ext(self.sqrt_recip_alphas_cumprod, i, x) * x - x0) / self._ext(self.sqrt_recipm1_alphas_cumprod, i, x)
                return eps, x0

            old_eps = None
            idx = range(n - 1, -1, -1)
            if progress:
                from tqdm.auto import tqdm
                idx = tqdm(idx)
            for i in idx:
                abp = self._ext(self.alphas_cumprod_prev, i, img)
                eps, x0 = model_out(img, i)
                if order > 1 and old_eps is None:      # pseudo improved Euler (first step)
                    old_eps = [eps]
                    mean_pred = x0 * torch.sqrt(abp) + torch.sqrt(1 - abp) * eps
                    eps_2, _ = model_out(mean_pred, i - 1)
                    eps_prime = (eps + eps_2) / 2
                else:                                   # Adams-Bashforth
                    old_eps = (old_eps or []) + [eps]
                    cur = min(order, len(old_eps))
                    if cur == 1:
                        eps_prime = old_eps[-1]
                    elif cur == 2:
                        eps_prime = (3 * old_eps[-1] - old_eps[-2]) / 2
                    elif cur == 3:
                        eps_prime = (23 * old_eps[-1] - 16 * old_eps[-2] + 5 * old_eps[-3]) / 12
                    else:
                        eps_prime = (55 * old_eps[-1] - 59 * old_eps[-2] + 37 * old_eps[-3] - 9 * old_eps[-4]) / 24
                pred_prime = self._ext(self.sqrt_recip_alphas_cumprod, i, img) * img - self._ext(self.sqrt_recipm1_alphas_cumprod, i, img) * eps_prime
                mean_pred = pred_prime * torch.sqrt(abp) + torch.sqrt(1 - abp) * eps_prime
                if len(old_eps) >= order:
                    old_eps.pop(0)
                img = mean_pred if i != 0 else x0
            return img


def create_gaussian_diffusion(args) -> Sampler:
    """Mirror of utils/model_util.py:79-114 (1000 base steps, cosine default, x0-prediction, fixed-small sigma)."""
    steps = 1000
    betas = named_beta_schedule(args.noise_schedule, steps, 1.0)
    resp = args.timestep_respacing or [steps]
    return Sampler(
        use_timesteps=space_timesteps(steps, resp), betas=betas, model_mean_type=ModelMeanType.START_X,
        model_var_type=ModelVarType.FIXED_SMALL if args.sigma_small else ModelVarType.FIXED_LARGE,
        loss_type=LossType.MSE, rescale_timesteps=False, lambda_vel=getattr(args, "lambda_vel", 0.0),
        data_format=args.data_format, model_path=getattr(args, "save_dir", getattr(args, "model_path", None)))
