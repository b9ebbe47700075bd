"""Denoiser -- drop-in for the reference's `FiLMTransformer` (model/diffusion.py:82-403).

Same constructor arguments, attributes, `state_dict()` keys and `forward(x, times, y, cond_drop_prob)`
signature; the computation is split the B200 way:

  * step-INVARIANT conditioning (frozen audio encoders, cond_projection, face cond_encoder,
    mean-pool MLP, keyframe projection -- model/diffusion.py:285-336,372-381) runs ONCE per distinct
    `y` in PyTorch and is cached; the reference recomputes all of it inside every call (2x per step
    with CFG).
  * per-layer rotated-K / V projections of those memories are built once by the CUDA library
    (a2p_denoiser_set_conditioning) and reused by every diffusion step.
  * the step-dependent decoder stack runs in hand-written sm_100a kernels behind the C-ABI
    (include/a2p_b200.h).  There is no PyTorch / CPU fallback for it.
"""
from __future__ import annotations

import ctypes as C
import os
import math
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .weights import EMB_LEN, KEYFRAME_STEP, ModelDims, denoiser_param_spec, model_dims, synthetic_state_dict

_FROZEN_PREFIXES = ("audio_model.", "lip_model.", "transformer.", "tokenizer.")


class _Node(nn.Module):
    """Anonymous container so that parameter paths equal the reference's state_dict keys."""


def _install(root: nn.Module, path: str, tensor: torch.Tensor, buffer: bool) -> None:
    parts = path.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, _Node())
        mod = getattr(mod, p)
    if buffer:
        mod.register_buffer(parts[-1], tensor)
    else:
        mod.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


def _rope(x: torch.Tensor, freqs: torch.Tensor) -> torch.Tensor:
    """Full-width interleaved-pair rotation at positions 0..L-1 (rotary_embedding_torch.py:46-66,116-139)."""
    L = x.shape[-2]
    ang = torch.einsum("p,f->pf", torch.arange(L, device=x.device).type(freqs.dtype), freqs)
    ang = ang.repeat_interleave(2, dim=-1)
    x2 = x.reshape(*x.shape[:-1], -1, 2)
    rot = torch.stack((-x2[..., 1], x2[..., 0]), dim=-1).reshape(x.shape)
    return x * ang.cos() + rot * ang.sin()


class Denoiser(nn.Module):
    def __init__(
        self,
        args,
        nfeats: int,
        latent_dim: int = 512,
        ff_size: int = 1024,
        num_layers: int = 4,
        num_heads: int = 4,
        dropout: float = 0.1,
        cond_feature_dim: int = 4800,
        activation: Callable[[torch.Tensor], torch.Tensor] = F.gelu,
        use_rotary: bool = True,
        cond_mode: str = "audio",
        split_type: str = "train",
        device: str = "cuda",
        audio_model: Optional[nn.Module] = None,
        lip_model: Optional[nn.Module] = None,
        split_terms: Optional[int] = None,
        **kwargs,
    ) -> None:
        super().__init__()
        if not use_rotary:
            raise NotImplementedError("a2p_b200 implements the rotary variant only (--not_rotary is out of scope)")
        if activation is not F.gelu and getattr(activation, "__name__", "") != "gelu":
            raise NotImplementedError("a2p_b200 implements the exact-erf GELU feed-forward only")
        self.nfeats = nfeats
        self.cond_mode = cond_mode
        self.cond_feature_dim = cond_feature_dim
        self.add_frame_cond = args.add_frame_cond
        self.data_format = args.data_format
        self.split_type = split_type
        self.device = device
        self.seq_len = getattr(args, "max_seq_length", 600)
        self.dims: ModelDims = model_dims(self.data_format, num_layers, num_heads, self.seq_len)
        if (nfeats, latent_dim, ff_size) != (self.dims.C, self.dims.D, self.dims.FF):
            raise ValueError(f"geometry {(nfeats, latent_dim, ff_size)} is not the {self.data_format} model "
                             f"{(self.dims.C, self.dims.D, self.dims.FF)} (utils/model_util.py:49-76)")
        self.split_terms = int(split_terms) if split_terms is not None else (2 if self.data_format == "pose" else 3)
        self.resume_trans = None
        if self.data_format == "pose":
            self.step = KEYFRAME_STEP
            self.use_cm = True
        else:
            self.use_cm = False
        # parameters / buffers under the reference's names (random init: fan-in scaled, see weights.py)
        init = synthetic_state_dict(self.dims, seed=0)
        for name, shape, kind in denoiser_param_spec(self.dims):
            _install(self, name, init[name].clone(), buffer=kind in ("f", "k"))
        # frozen side models, set up where the reference's constructor sets them up (model/diffusion.py:140,147-157,
        # 226-277).  They stay PyTorch (north_star): explicit `audio_model=` / `lip_model=` arguments win, otherwise
        # the same loaders the reference calls are used when they are importable.
        self.audio_model = audio_model
        self.lip_model = lip_model
        self.setup_audio_models()
        if self.data_format == "pose":
            self.setup_guide_models(args)
        elif self.lip_model is None:
            self.setup_lip_models()
        self._frozen_state: Dict[str, torch.Tensor] = {}
        # runtime state
        self._handle: Optional[C.c_void_p] = None
        self._bound_sig = None
        self._cond_sig = None
        self._cond_keys = (None, None, None)
        self.cond_cache_hits = 0
        self._packed = self._ws = None
        self._kv = [None, None]
        self._keep = []
        self._max_pos = 0

    # ------------------------------------------------------------------ checkpoint contract
    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts reference checkpoints: frozen fairseq sub-modules that this object does not hold as
        nn.Modules are stashed instead of being reported as unexpected (utils/model_util.py:30-38)."""
        mine = {}
        for k, v in state_dict.items():
            top = k.split(".")[0]
            if k.startswith(_FROZEN_PREFIXES) and getattr(self, top, None) is None:
                self._frozen_state[k] = v
            else:
                mine[k] = v
        res = super().load_state_dict(mine, strict=strict, assign=assign)
        self._bound_sig = None
        return res

    def parameters_w_grad(self):
        return [p for p in self.parameters() if p.requires_grad]

    # ------------------------------------------------------------------ library plumbing
    def _sd(self) -> Dict[str, torch.Tensor]:
        return {k: v for k, v in self.state_dict(keep_vars=True).items() if not k.startswith(_FROZEN_PREFIXES)}

    def _ensure_bound(self, dev: torch.device, need_pos: int) -> None:
        if dev.type != "cuda":
            raise _lib.A2PError("a2p_b200 has no CPU path: move the model and inputs to a CUDA device")
        sd = self._sd()
        sig = (dev, tuple((t.data_ptr(), t._version) for t in sd.values()))
        if self._handle is not None and sig == self._bound_sig and need_pos <= self._max_pos:
            return
        lib = _lib.load()
        d = self.dims
        self._max_pos = max(need_pos, EMB_LEN + 2, self.seq_len)
        cfg = _lib.ModelCfg(fmt=0 if d.fmt == "pose" else 1, C=d.C, D=d.D, L=d.L, H=d.H, FF=d.FF, S2=d.S2,
                            max_pos=self._max_pos, split_terms=self.split_terms, reserved=0)
        self._cfg = cfg
        if self._handle is not None:
            lib.a2p_denoiser_destroy(self._handle)
            self._handle = None
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.a2p_denoiser_create(C.byref(h), C.byref(cfg)))
            self._handle = h
            tensors = {}
            for k, t in sd.items():
                if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                    raise _lib.A2PError(f"parameter {k} must be a contiguous fp32 tensor on {dev}")
                tensors[k] = t
            half = d.D // 2
            # timestep-embedding frequencies with the reference's own expression (model/utils.py:74-76)
            tensors["a2p.time_freqs"] = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).to(dev)
            table = (_lib.Weight * len(tensors))()
            self._keep = [tensors]
            for i, (k, t) in enumerate(tensors.items()):
                table[i].name = k.encode()
                table[i].ptr = t.data_ptr()
                table[i].numel = t.numel()
            nbytes = lib.a2p_packed_weight_bytes(C.byref(cfg))
            self._packed = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.a2p_denoiser_bind_weights(h, table, len(tensors), self._packed.data_ptr(), nbytes, st))
        self._bound_sig = sig
        self._cond_sig = None

    def _workspace(self, nbytes: int, dev) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return self._ws

    # ------------------------------------------------------------------ step-invariant conditioning (PyTorch, once)
    def setup_audio_models(self) -> None:
        """model/diffusion.py:270-271 + model/utils.py:18-26: frozen vq-wav2vec feature extractor through
        fairseq.checkpoint_utils (./assets/vq-wav2vec.pt).  Without an importable fairseq the model can only be fed
        precomputed features (y["audio_embed"]); raw audio then fails loudly in _audio_features.  The 48 kHz -> 16 kHz
        resampler is the `audio_resampler.kernel` buffer of the checkpoint contract (installed with the parameters)."""
        if self.audio_model is not None or self.cond_mode == "uncond":
            return
        try:
            import fairseq  # noqa: F401
        except ImportError:
            return
        cp_path = "./assets/vq-wav2vec.pt"
        audio_model, _, _ = fairseq.checkpoint_utils.load_model_ensemble_and_task([cp_path])
        audio_model = audio_model[0]
        for param in audio_model.parameters():
            param.requires_grad = False
        audio_model.eval()
        self.audio_model = audio_model

    def setup_lip_models(self) -> None:
        """model/diffusion.py:273-280: the frozen lip regressor (a reference-side PyTorch module, importable when this
        package is used inside a reference checkout as INTEGRATION.md describes) with ./assets/iter-0200000.pt."""
        try:
            from model.diffusion import Audio2LipRegressionTransformer  # reference checkout on sys.path
        except Exception:
            return
        import os
        cp_path = "./assets/iter-0200000.pt"
        if not os.path.exists(cp_path):
            return
        lip = Audio2LipRegressionTransformer()
        cp = torch.load(cp_path, map_location="cpu")
        lip.load_state_dict(cp["model_state_dict"])
        for param in lip.parameters():
            param.requires_grad = False
        self.lip_model = lip.eval()

    def setup_guide_models(self, args) -> None:
        """model/diffusion.py:226-268: at test time with --resume_trans the guide transformer and its VQ tokenizer
        are loaded next to the denoiser; sample/generate.py:51-71 calls model.transformer.generate and
        model.tokenizer.decode before the loop.  Built by the guide sampler of this package (guide.py)."""
        if self.split_type == "test" and getattr(args, "resume_trans", None) is not None:
            from .guide import load_guide_predictor
            self.resume_trans = args.resume_trans
            self.tokenizer, self.transformer = load_guide_predictor(args.resume_trans)

    def _resample_48k_16k(self, wave: torch.Tensor) -> torch.Tensor:
        """torchaudio.transforms.Resample(48000, 16000).forward with the module's `kernel` buffer (orig/gcd = 3,
        new/gcd = 1, width 19): pad, stride-3 FIR, crop to ceil(L/3) -- model/diffusion.py:286-287."""
        k = self.audio_resampler.kernel
        width = (k.shape[-1] - 3) // 2
        shape = wave.shape
        w = F.pad(wave.reshape(-1, shape[-1]), (width, width + 3))
        out = F.conv1d(w[:, None], k, stride=3).transpose(1, 2).reshape(w.shape[0], -1)
        out = out[..., : -(-shape[-1] // 3)]
        return out.reshape(shape[:-1] + out.shape[-1:])

    def encode_audio(self, raw_audio: torch.Tensor) -> torch.Tensor:
        """model/diffusion.py:285-293 -> [B, S, 1024]"""
        dev = next(self.parameters()).device
        a0 = self._resample_48k_16k(raw_audio[:, :, 0].to(dev, torch.float32))
        a1 = self._resample_48k_16k(raw_audio[:, :, 1].to(dev, torch.float32))
        with torch.no_grad():
            z0 = self.audio_model.feature_extractor(a0)
            z1 = self.audio_model.feature_extractor(a1)
            return torch.cat((z0, z1), dim=1).permute(0, 2, 1)

    def _audio_features(self, y, dev) -> torch.Tensor:
        """encode_audio (+ encode_lip for face) output [B,S,cond_dim] (model/diffusion.py:285-313), computed ONCE per
        distinct y (the reference recomputes it in every denoiser call).  `y["audio_embed"]` (precomputed wav2vec
        features: BASELINE's "synthetic wav2vec features") wins over raw audio."""
        if "audio_embed" in y:
            return y["audio_embed"].to(dev, torch.float32)
        if self.audio_model is None:
            raise _lib.A2PError("no frozen audio encoder available (fairseq is not installed): pass y['audio_embed'] "
                                "[B,S,%d] or construct Denoiser(audio_model=...)" % self.dims.cond_dim)
        raw = y["audio"].to(dev)
        emb = self.encode_audio(raw)
        if self.data_format == "face":
            if self.lip_model is None:
                raise _lib.A2PError("face model needs lip_model for raw audio; pass y['audio_embed'] instead")
            B = raw.shape[0]
            chunks = raw.reshape(B, -1, 1600, 2)[..., 0]
            lips = torch.cat([self.lip_model(chunks[:, i:i + 120]) for i in range(0, chunks.shape[1], 120)], dim=1)
            lips = lips.permute(0, 2, 3, 1).reshape(B, 338 * 3, -1)
            lips = F.interpolate(lips, size=emb.shape[1], mode="nearest-exact").permute(0, 2, 1)
            emb = torch.cat((emb, lips), dim=-1)
        return emb

    def _encoder_layer(self, x: torch.Tensor, sd, p: str) -> torch.Tensor:
        """TransformerEncoderLayerRotary, pre-LN (face cond_encoder; transformer_modules.py:69-102)."""
        D, H = self.dims.D, self.dims.H
        fr = sd[p + ".rotary.freqs"]
        h = F.layer_norm(x, (D,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5)
        hr = _rope(h, fr)
        W, b = sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"]
        q = F.linear(hr, W[:D], b[:D]); k = F.linear(hr, W[D:2 * D], b[D:2 * D]); v = F.linear(h, W[2 * D:], b[2 * D:])
        B, S, _ = q.shape
        sp = lambda t: t.view(B, S, H, D // H).transpose(1, 2)
        a = F.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(B, S, D)
        x = x + F.linear(a, sd[p + ".self_attn.out_proj.weight"], sd[p + ".self_attn.out_proj.bias"])
        h = F.layer_norm(x, (D,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)
        return x + F.linear(F.gelu(F.linear(h, sd[p + ".linear1.weight"], sd[p + ".linear1.bias"])),
                            sd[p + ".linear2.weight"], sd[p + ".linear2.bias"])

    def _encode_conditioning_torch(self, feats, pred, sd):
        """The conditioning encoders in PyTorch (A2P_COND_TORCH=1 and the parity test of the native encoders): cond_projection
        (-> face cond_encoder), mean-pool MLP, keyframe projection -- model/diffusion.py:355-381, 316-336.  One sample at a
        time: library GEMMs pick shape-dependent algorithms, and a row's result must not depend on how the batch is sharded."""
        d = self.dims
        D = d.D

        def per_sample(fn, t):
            return torch.cat([fn(t[i:i + 1]) for i in range(t.shape[0])], dim=0)

        def tokens_of(f1):
            tk = F.linear(f1, sd["cond_projection.weight"], sd["cond_projection.bias"])
            if d.fmt == "face":
                for i in range(2):
                    tk = self._encoder_layer(tk, sd, f"cond_encoder.{i}")
            return tk

        def hidden_of(tokens):
            hdn = F.layer_norm(tokens.mean(dim=-2), (D,), sd["non_attn_cond_projection.0.weight"],
                               sd["non_attn_cond_projection.0.bias"], 1e-5)
            hdn = F.silu(F.linear(hdn, sd["non_attn_cond_projection.1.weight"], sd["non_attn_cond_projection.1.bias"]))
            return F.linear(hdn, sd["non_attn_cond_projection.3.weight"], sd["non_attn_cond_projection.3.bias"])

        tok = per_sample(tokens_of, feats)
        hid = per_sample(hidden_of, tok)
        pose_c = None
        if pred is not None:
            ph = per_sample(lambda k1: F.linear(k1, sd["frame_cond_projection.weight"], sd["frame_cond_projection.bias"]), pred)
            pose_c = F.layer_norm(ph, (D,), sd["frame_norm_cond.weight"], sd["frame_norm_cond.bias"], 1e-5).contiguous()
        return tok, hid, pose_c

    @torch.no_grad()
    def prepare(self, y, batch_size: int, T: int, dev: torch.device) -> None:
        """Compute + upload the conditioning of BOTH branches unless `y` is unchanged since the last call."""
        d = self.dims
        key_t = y.get("audio_embed", y.get("audio")) if self.cond_mode != "uncond" else None
        kf = y.get("keyframes") if d.fmt == "pose" else None
        msk = y.get("mask") if d.fmt == "pose" else None
        if kf is not None:
            # pad the unknown keyframes in place on the caller's tensor like model/diffusion.py:318-320 -- only when it
            # changes something, so that an unchanged y keeps its tensor versions (and hits the cache below)
            new_mask = msk[..., :: self.step].reshape(kf.shape[0], -1).to(kf.device)
            if bool((kf[~new_mask] != 0).any()):
                kf[~new_mask] = 0.0
        # The cache key is (tensor identity, version, shape).  The key tensors are HELD (self._cond_keys) so that their
        # ids cannot be recycled by fresh tensors after the caller drops y (a recycled id + version 0 would silently
        # reuse the previous request's conditioning).
        keys = (key_t, kf, msk)
        sig = tuple((id(t), t._version, tuple(t.shape)) if torch.is_tensor(t) else None for t in keys)
        sig = sig + (batch_size, T, self._bound_sig is not None and id(self._packed))
        if sig == self._cond_sig and all(a is b for a, b in zip(keys, self._cond_keys)):
            self.cond_cache_hits += 1
            return
        sd = self._sd()
        lib = _lib.load()
        D = d.D
        if self.cond_mode == "uncond":
            feats = torch.zeros(batch_size, T, self.cond_feature_dim, device=dev)
        else:
            feats = self._audio_features(y, dev)
        S = feats.shape[1]
        if S > EMB_LEN:
            raise ValueError(f"{S} audio tokens exceed null_cond_embed's {EMB_LEN} rows (model/diffusion.py:136,378)")
        pred = None
        nk = 0
        if d.fmt == "pose":
            pred = y["keyframes"].detach().clone().to(dev, torch.float32).contiguous()   # unknown keyframes zeroed above
            nk = pred.shape[1]
        if os.environ.get("A2P_COND_TORCH"):
            tok, hid, pose_c = self._encode_conditioning_torch(feats, pred, sd)
        else:
            # native encoders (a2p_denoiser_encode_conditioning: exact-fp32 FFMA GEMMs / fp32 attention, batch-invariant per row)
            feats = feats.to(dev, torch.float32).contiguous()
            Fd = feats.shape[2]
            tok = torch.empty(batch_size, S, D, device=dev)
            hid = torch.empty(batch_size, D, device=dev)
            pose_c = torch.empty(batch_size, nk, D, device=dev) if pred is not None else None
            ews = self._workspace(lib.a2p_encode_workspace_bytes(C.byref(self._cfg), batch_size, S, Fd), dev)
            _lib.check(lib.a2p_denoiser_encode_conditioning(
                self._handle, batch_size, S, nk, Fd, feats.data_ptr(), pred.data_ptr() if pred is not None else None,
                tok.data_ptr(), hid.data_ptr(), pose_c.data_ptr() if pose_c is not None else None, ews.data_ptr(), ews.numel(),
                torch.cuda.current_stream(dev).cuda_stream))
        pose_u = sd["null_pose_embed"][:, :nk].contiguous() if d.fmt == "pose" else None
        sets = [
            (0, batch_size, tok.contiguous(), hid.contiguous(), pose_c),
            (1, 1, sd["null_cond_embed"][:, :S].contiguous(), sd["null_cond_hidden"].contiguous(), pose_u),
        ]
        st = torch.cuda.current_stream(dev).cuda_stream
        cfg = self._cfg
        for branch, bc, tokens, hidden, pose in sets:
            kvb = lib.a2p_kv_cache_bytes(C.byref(cfg), bc, S)
            if self._kv[branch] is None or self._kv[branch].numel() < kvb or self._kv[branch].device != dev:
                self._kv[branch] = torch.empty(kvb, dtype=torch.uint8, device=dev)
            wsb = lib.a2p_conditioning_workspace_bytes(C.byref(cfg), bc, S)
            ws = self._workspace(max(wsb, lib.a2p_workspace_bytes(C.byref(cfg), batch_size, T)), dev)
            _lib.check(lib.a2p_denoiser_set_conditioning(
                self._handle, branch, bc, S, nk, tokens.data_ptr(), hidden.data_ptr(),
                pose.data_ptr() if pose is not None else None, self._kv[branch].data_ptr(), self._kv[branch].numel(),
                ws.data_ptr(), ws.numel(), st))
        self._cond_keep = sets
        self._cond_sig = sig
        self._cond_keys = keys
        self._cond_S = S

    # ------------------------------------------------------------------ reference-compatible forward
    def _run(self, x: torch.Tensor, times: torch.Tensor, y, mask: int):
        if x.dim() not in (3, 4):
            raise ValueError("x must be [B,C,1,T] or [B,T,C]")
        dev = x.device
        layout = 0 if x.dim() == 4 else 1
        B = x.shape[0]
        T = x.shape[3] if layout == 0 else x.shape[1]
        self._ensure_bound(dev, max(T, EMB_LEN + 2))
        with torch.cuda.device(dev):
            self.prepare(y if y is not None else {}, B, T, dev)
            lib = _lib.load()
            x = x.contiguous().float()
            times = times.to(dev, torch.int64).contiguous()
            outs = [torch.empty(B, T, self.dims.C, device=dev) if mask & m else None for m in (1, 2)]
            ws = self._workspace(lib.a2p_workspace_bytes(C.byref(self._cfg), B, T), dev)
            st = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(lib.a2p_denoiser_forward(
                self._handle, B, T, x.data_ptr(), layout, times.data_ptr(), mask,
                outs[0].data_ptr() if outs[0] is not None else None,
                outs[1].data_ptr() if outs[1] is not None else None, ws.data_ptr(), ws.numel(), st))
        return outs

    @torch.no_grad()
    def forward(self, x: torch.Tensor, times: torch.Tensor, y: Optional[dict] = None, cond_drop_prob: float = 0.0):
        """Same contract as FiLMTransformer.forward (model/diffusion.py:338-403): returns [B,T,C].
        Inference keep-masks are deterministic (model/utils.py:83-87), so cond_drop_prob must be 0 or 1."""
        if cond_drop_prob not in (0, 0.0, 1, 1.0):
            raise NotImplementedError("stochastic conditioning dropout is a training feature (out of scope)")
        mask = 1 if cond_drop_prob == 0 else 2
        outs = self._run(x, times, y, mask)
        return outs[0] if mask == 1 else outs[1]

    def launch_count(self) -> int:
        return int(_lib.load().a2p_launch_count(self._handle)) if self._handle is not None else 0

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().a2p_denoiser_destroy(self._handle)
        except Exception:
            pass


class CFGDenoiser(nn.Module):
    """Drop-in for `ClassifierFreeSampleModel` (model/cfg_sampler.py:17-33): both branches run as ONE
    2B-row batch through the library instead of two sequential model calls."""

    def __init__(self, model: Denoiser):
        super().__init__()
        self.model = model
        self.nfeats = model.nfeats
        self.cond_mode = model.cond_mode
        self.add_frame_cond = model.add_frame_cond
        if self.add_frame_cond is not None:
            if getattr(model, "resume_trans", None) is not None:
                self.transformer = model.transformer
                self.tokenizer = model.tokenizer
            self.step = model.step

    @torch.no_grad()
    def forward(self, x, timesteps, y=None):
        out, out_uncond = self.model._run(x, timesteps, y, 3)
        return out_uncond + (y["scale"].to(out.device).view(-1, 1, 1) * (out - out_uncond))
