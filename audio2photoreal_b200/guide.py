"""Guide-keyframe predictor of the body model: runs ONCE per sample before the diffusion loop.

  * `GuideSampler` -- drop-in for the reference's `GuideTransformer` at inference (model/guide.py:26-222): same checkpoint
    keys, same `generate(condition, sequence_length, layers, n_sequences, ...)` signature and nucleus sampling, but the
    autoregressive loop is O(n) instead of O(n^2): the reference re-encodes the raw audio and re-runs the whole prefix for
    every one of the 80 tokens (model/guide.py:149,200); here the audio memory (resample -> frozen vq-wav2vec -> TCN ->
    projection -> norm), its per-layer rotated-K / V projections and the FiLM vectors are computed once, and every step
    appends one row to a per-layer self-attention K/V cache.  Causality makes this exact: position n of the reference's
    masked forward depends only on positions <= n.
  * `VQDecoder` -- drop-in for `TemporalVertexCodec.decode` (model/vqvae.py:381-392,454-464,508-521): residual-codebook
    lookup + the causal dilated conv decoder, same checkpoint keys (`ckpt["net"]`).
  * `load_guide_predictor(resume_trans)` -- what `FiLMTransformer.setup_guide_predictor` does (model/diffusion.py:240-268).
  * `decode_and_save(...)` -- the de-normalise + `results.npy` contract of sample/generate.py:98-152,289-292.

This is host-side PyTorch by design (north_star: "Host code stays Python/PyTorch for tensor plumbing and the frozen
wav2vec/VQ encoders"): 80 sequential micro-steps of <= 81 tokens against a 798-row memory, far off the hot path.
"""
from __future__ import annotations

import json
import math
import os
from typing import Callable, Dict, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

_FROZEN = ("audio_model.",)


def _rope(x: torch.Tensor, freqs: torch.Tensor, pos0: int = 0) -> torch.Tensor:
    """Interleaved-pair rotation over the full width at positions pos0 .. pos0+L-1
    (model/modules/rotary_embedding_torch.py:46-66,116-139)."""
    L = x.shape[-2]
    pos = torch.arange(pos0, pos0 + L, device=x.device).type(freqs.dtype)
    ang = torch.einsum("p,f->pf", pos, freqs).repeat_interleave(2, dim=-1)
    x2 = x.reshape(*x.shape[:-1], -1, 2)
    rot = torch.stack((-x2[..., 1], x2[..., 0]), dim=-1).reshape(x.shape)
    return x * ang.cos() + rot * ang.sin()


def _install(root: nn.Module, path: str, tensor: torch.Tensor) -> None:
    parts = path.split(".")
    mod = root
    for p in parts[:-1]:
        if not hasattr(mod, p):
            mod.add_module(p, nn.Module())
        mod = getattr(mod, p)
    mod.register_buffer(parts[-1], tensor)


class _Frozen(nn.Module):
    """Holds a reference checkpoint's tensors as buffers under the reference's names (state_dict keys identical)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], skip=()):
        super().__init__()
        self._names = []
        for k, v in state_dict.items():
            if k.startswith(tuple(skip)):
                continue
            _install(self, k, v.detach().clone())
            self._names.append(k)

    def t(self, name: str) -> torch.Tensor:
        mod = self
        for p in name.split("."):
            mod = getattr(mod, p)
        return mod


def nucleus_sample(logits: torch.Tensor, top_p: float, draw: Optional[Callable[[torch.Tensor], torch.Tensor]] = None) -> torch.Tensor:
    """model/guide.py:203-217: softmax, sort, keep the smallest prefix with cumulative mass >= top_p (the first token is
    always kept), renormalise, sample one index.  `draw(probs)` defaults to Categorical(probs).sample()."""
    one_hot = F.softmax(logits, dim=-1)
    sorted_probs, indices = torch.sort(one_hot, dim=-1, descending=True)
    cumulative = torch.cumsum(sorted_probs, dim=-1)
    nucleus = cumulative < top_p
    nucleus = torch.cat([nucleus.new_ones(nucleus.shape[:-1] + (1,)), nucleus[..., :-1]], dim=-1)
    sorted_probs = sorted_probs.masked_fill(~nucleus, 0.0)
    sorted_probs = sorted_probs / sorted_probs.sum(-1, keepdim=True)
    idx = draw(sorted_probs) if draw is not None else torch.distributions.Categorical(sorted_probs).sample()
    return indices.gather(-1, idx.unsqueeze(-1))


class GuideSampler(_Frozen):
    def __init__(self, state_dict: Dict[str, torch.Tensor], tokens: int, num_heads: int = 4, audio_model: Optional[nn.Module] = None):
        super().__init__(state_dict, skip=_FROZEN)
        self.tokens = tokens
        self.num_heads = num_heads
        self.dim = self.t("token_embedding.weight").shape[1]
        self.num_layers = 1 + max(int(k.split(".")[2]) for k in state_dict if k.startswith("seqTransDecoder.stack."))
        self.conv_ids = sorted({int(k.split(".")[1]) for k in state_dict if k.startswith("pre_audio.") and k.endswith(".weight")})
        self.audio_model = audio_model
        if audio_model is None and any(k.startswith("audio_model.") for k in state_dict):
            # the frozen vq-wav2vec lives in the checkpoint (model/guide.py:118-119 -> model/utils.py:18-26)
            try:
                import fairseq  # noqa: F401
                am, _, _ = fairseq.checkpoint_utils.load_model_ensemble_and_task(["./assets/vq-wav2vec.pt"])
                am = am[0]
                am.load_state_dict({k[len("audio_model."):]: v for k, v in state_dict.items() if k.startswith("audio_model.")}, strict=False)
                for p in am.parameters():
                    p.requires_grad = False
                self.audio_model = am.eval()
            except ImportError:
                pass

    # ------------------------------------------------------------------ one-time audio memory
    def _resample(self, wave: torch.Tensor) -> torch.Tensor:
        k = self.t("audio_resampler.kernel")
        width = (k.shape[-1] - 3) // 2
        shape = wave.shape
        w = F.pad(wave.reshape(-1, shape[-1]), (width, width + 3))
        out = F.conv1d(w[:, None], k, stride=3).transpose(1, 2).reshape(w.shape[0], -1)[..., : -(-shape[-1] // 3)]
        return out.reshape(shape[:-1] + out.shape[-1:])

    def encode_audio(self, raw_audio: torch.Tensor) -> torch.Tensor:
        """model/guide.py:121-129"""
        dev = self.t("final_layer.weight").device
        if self.audio_model is None:
            raise RuntimeError("GuideSampler has no frozen audio encoder (fairseq is not importable): pass audio_model=")
        a0 = self._resample(raw_audio[:, :, 0].to(dev, torch.float32))
        a1 = self._resample(raw_audio[:, :, 1].to(dev, torch.float32))
        with torch.no_grad():
            z0 = self.audio_model.feature_extractor(a0)
            z1 = self.audio_model.feature_extractor(a1)
        return torch.cat((z0, z1), dim=1).permute(0, 2, 1)

    def _pre_audio(self, x: torch.Tensor) -> torch.Tensor:
        """the (un-padded) dilated conv stack of setup_audio_models, eval mode (model/guide.py:84-119): [B,C,S] -> [B,C,S']"""
        dil = (1, 2, 3, 1, 2, 3)
        n_conv = len(self.conv_ids)
        for j, i in enumerate(self.conv_ids):
            w, b = self.t(f"pre_audio.{i}.weight"), self.t(f"pre_audio.{i}.bias")
            if j == n_conv - 1 and w.shape[-1] == 1:
                x = F.conv1d(x, w, b)
            else:
                x = F.leaky_relu(F.conv1d(x, w, b, dilation=dil[j % 6]), 0.2)
        return x

    def _ln(self, x, name):
        return F.layer_norm(x, (x.shape[-1],), self.t(name + ".weight"), self.t(name + ".bias"), 1e-5)

    @torch.no_grad()
    def memory(self, condition: torch.Tensor):
        """everything that does not depend on the generated tokens (model/guide.py:149-171)"""
        cond_embed = self.encode_audio(condition)
        cond_tokens = self._pre_audio(cond_embed.permute(0, 2, 1)).permute(0, 2, 1)
        cond_tokens = F.linear(cond_tokens, self.t("cond_projection.weight"), self.t("cond_projection.bias"))
        hdn = self._ln(cond_tokens.mean(dim=-2), "non_attn_cond_projection.0")
        hdn = F.silu(F.linear(hdn, self.t("non_attn_cond_projection.1.weight"), self.t("non_attn_cond_projection.1.bias")))
        cond_hidden = F.linear(hdn, self.t("non_attn_cond_projection.3.weight"), self.t("non_attn_cond_projection.3.bias"))
        mem = self._ln(cond_tokens, "norm_cond")
        D, H = self.dim, self.num_heads
        freqs = self.t("rotary.freqs")
        mem_r = _rope(mem, freqs)
        mish_t = F.mish(cond_hidden)
        layers = []
        B, S, _ = mem.shape
        sp = lambda t: t.view(B, -1, H, D // H).transpose(1, 2)
        for n in range(self.num_layers):
            p = f"seqTransDecoder.stack.{n}."
            W, b = self.t(p + "multihead_attn.in_proj_weight"), self.t(p + "multihead_attn.in_proj_bias")
            film = [F.linear(mish_t, self.t(p + f"film{i}.block.1.weight"), self.t(p + f"film{i}.block.1.bias")).unsqueeze(1).chunk(2, dim=-1)
                    for i in (1, 2, 3)]
            layers.append({"k": sp(F.linear(mem_r, W[D:2 * D], b[D:2 * D])), "v": sp(F.linear(mem, W[2 * D:], b[2 * D:])), "film": film,
                           "sk": None, "sv": None})
        return layers

    def _step(self, tok: torch.Tensor, pos: int, layers) -> torch.Tensor:
        """logits of the token at position `pos` given the cached prefix (model/guide.py:140-172 for ONE new row)."""
        D, H = self.dim, self.num_heads
        freqs = self.t("rotary.freqs")
        x = F.embedding(tok, self.t("token_embedding.weight"))                     # [B,1,D]
        B = x.shape[0]
        sp = lambda t: t.view(B, -1, H, D // H).transpose(1, 2)
        for n, lc in enumerate(layers):
            p = f"seqTransDecoder.stack.{n}."
            h = self._ln(x, p + "norm1")
            hr = _rope(h, freqs, pos)
            W, b = self.t(p + "self_attn.in_proj_weight"), self.t(p + "self_attn.in_proj_bias")
            q, k, v = sp(F.linear(hr, W[:D], b[:D])), sp(F.linear(hr, W[D:2 * D], b[D:2 * D])), sp(F.linear(h, W[2 * D:], b[2 * D:]))
            lc["sk"] = k if lc["sk"] is None else torch.cat((lc["sk"], k), dim=2)
            lc["sv"] = v if lc["sv"] is None else torch.cat((lc["sv"], v), dim=2)
            a = F.scaled_dot_product_attention(q, lc["sk"], lc["sv"]).transpose(1, 2).reshape(B, 1, D)
            a = F.linear(a, self.t(p + "self_attn.out_proj.weight"), self.t(p + "self_attn.out_proj.bias"))
            sc, sh = lc["film"][0]
            x = x + ((sc + 1) * a + sh)
            h = _rope(self._ln(x, p + "norm2"), freqs, pos)
            W, b = self.t(p + "multihead_attn.in_proj_weight"), self.t(p + "multihead_attn.in_proj_bias")
            a = F.scaled_dot_product_attention(sp(F.linear(h, W[:D], b[:D])), lc["k"], lc["v"]).transpose(1, 2).reshape(B, 1, D)
            a = F.linear(a, self.t(p + "multihead_attn.out_proj.weight"), self.t(p + "multihead_attn.out_proj.bias"))
            sc, sh = lc["film"][1]
            x = x + ((sc + 1) * a + sh)
            h = self._ln(x, p + "norm3")
            a = F.linear(F.gelu(F.linear(h, self.t(p + "linear1.weight"), self.t(p + "linear1.bias"))),
                         self.t(p + "linear2.weight"), self.t(p + "linear2.bias"))
            sc, sh = lc["film"][2]
            x = x + ((sc + 1) * a + sh)
        return F.linear(x, self.t("final_layer.weight"), self.t("final_layer.bias"))[:, -1, :]

    @torch.no_grad()
    def logits_for(self, tokens: torch.Tensor, condition: torch.Tensor) -> torch.Tensor:
        """teacher-forced logits [B, n, tokens] of `GuideTransformer.forward(tokens, condition)` via the cached path"""
        layers = self.memory(condition)
        return torch.stack([self._step(tokens[:, i:i + 1], i, layers) for i in range(tokens.shape[1])], dim=1)

    @torch.no_grad()
    def generate(self, condition: torch.Tensor, sequence_length: int, layers: int, n_sequences: int = 1, max_key_len: int = 8,
                 max_seq_len: int = 240, top_p: float = 0.94, draw=None) -> torch.Tensor:
        """model/guide.py:174-222 -> [n_sequences, sequence_length * layers] int64 tokens (start token dropped)"""
        assert max_key_len == int(max_seq_len / 30), "currently only running for 1fps"
        cache = self.memory(condition)
        dev = self.t("final_layer.weight").device
        tok = torch.zeros(n_sequences, 1, dtype=torch.int64, device=dev) + self.tokens
        out = []
        for i in range(sequence_length * layers):
            tok = nucleus_sample(self._step(tok, i, cache), top_p, draw)
            out.append(tok)
        return torch.cat(out, dim=-1).contiguous()


class VQDecoder(_Frozen):
    """`TemporalVertexCodec` at inference: only `decode` (and the attributes the callers read)."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], n_vertices: int, latent_dim: int, categories: int, residual_depth: int):
        super().__init__(state_dict)
        self.latent_dim, self.categories, self.residual_depth, self.n_clusters = latent_dim, categories, residual_depth, categories
        self.n_vertices = n_vertices
        self.receptive_field = 8

    @torch.no_grad()
    def decode(self, q: torch.Tensor) -> torch.Tensor:
        """[B, T, residual_depth] (or [N, residual_depth]) codes -> [B, T, n_vertices] (model/vqvae.py:381-392,454-464,508-521)"""
        reformat = q.dim() > 2
        if reformat:
            B, T, _ = q.shape
            q = q.reshape(-1, self.residual_depth)
        enc = torch.tensor(0.0, device=q.device)
        for i in range(q.shape[1]):
            enc = enc + F.embedding(q[:, i], self.t(f"quantizer.layers.{i}._codebook.embed"))   # project_out is Identity (dim == codebook dim)
        if reformat:
            enc = enc.reshape(B, T, -1)
        x = F.pad(enc.permute(0, 2, 1).contiguous(), (self.receptive_field - 1, 0))
        for i, dil in ((0, 1), (2, 2), (4, 3), (6, 1)):
            x = F.leaky_relu(F.conv1d(x, self.t(f"decoder.dec.{i}.weight"), self.t(f"decoder.dec.{i}.bias"), dilation=dil), 0.2)
        x = F.conv1d(x, self.t("decoder.dec.8.weight"), self.t("decoder.dec.8.bias"))
        return x.permute(0, 2, 1)


def setup_tokenizer(resume_pth: str, device=None) -> VQDecoder:
    """model/vqvae.py:18-34"""
    with open(os.path.join(os.path.dirname(resume_pth), "args.json")) as f:
        a = json.load(f)
    ckpt = torch.load(resume_pth, map_location="cpu")
    tok = VQDecoder(ckpt["net"], a["nb_joints"], a["output_emb_width"], a["code_dim"], a["depth"])
    dev = device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu")
    return tok.to(dev).eval()


def load_guide_predictor(cp_path: str, device=None, audio_model: Optional[nn.Module] = None):
    """model/diffusion.py:240-268: <dir>/args.json names the tokenizer checkpoint and the transformer geometry."""
    cp_dir = cp_path.split("checkpoints/iter-")[0]
    with open(f"{cp_dir}/args.json") as f:
        trans_args = json.load(f)
    tokenizer = setup_tokenizer(trans_args["resume_pth"], device)
    cp = torch.load(cp_path, map_location="cpu")
    sd = cp["model_state_dict"]
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("seqTransDecoder.stack."))
    assert n_layers == trans_args["layers"], (n_layers, trans_args["layers"])
    assert sd["token_embedding.weight"].shape == (tokenizer.n_clusters + 1, trans_args["dim"])
    guide = GuideSampler(sd, tokens=tokenizer.n_clusters, audio_model=audio_model)
    guide = guide.to(next(tokenizer.buffers()).device).eval()
    return tokenizer, guide


# ------------------------------------------------------------------ N3: de-normalise + results.npy
def inv_transform(data, data_type: str, stats: Dict[str, np.ndarray]):
    """Social.inv_transform (data_loaders/data.py:71-91) with the `data_stats.pth` dictionary of the dataset."""
    if data_type == "pose":
        std, mean = stats["pose_std_flat"], stats["pose_mean"]
    elif data_type == "face":
        std, mean = stats["code_std_flat"], stats["code_mean"]
    elif data_type == "audio":
        std, mean = stats["audio_std_flat"], stats["audio_mean"]
    else:
        raise ValueError(f"Data type not supported: {data_type}")
    if torch.is_tensor(data):
        return data * torch.as_tensor(std, dtype=data.dtype, device=data.device) + torch.as_tensor(mean, dtype=data.dtype, device=data.device)
    return data * std + mean


def results_block(samples, audio, gt, lengths, keyframes) -> Dict[str, np.ndarray]:
    """the dictionary sample/generate.py:146-152 saves with np.save(<output_dir>/results.npy, block) (:289-292)"""
    cat = lambda xs: np.concatenate([np.asarray(x.cpu() if torch.is_tensor(x) else x) for x in xs], axis=0)
    return {"motions": cat(samples), "audio": cat(audio), "gt": cat(gt), "lengths": cat(lengths), "keyframes": cat(keyframes)}


def save_results(path: str, block: Dict[str, np.ndarray]) -> None:
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    np.save(path, block)
