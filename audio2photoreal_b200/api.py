"""Factory functions mirroring the reference's utils/model_util.py so `sample/generate.py:_setup_model`
(and demo/demo.py) can be pointed at the B200 path without edits, plus `patch_reference()`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .denoiser import CFGDenoiser, Denoiser
from .sampler import Sampler, create_gaussian_diffusion


def default_split_terms(args) -> int:
    """Arithmetic arm of the CUDA path.  `args.split_terms` (or A2P_SPLIT_TERMS) wins; otherwise the fastest arm that
    meets rtol 1e-3 / atol 1e-4 against the reference on that model: two bf16 planes (16 mantissa bits, 3 tensor-core
    products per MAC, fused row-chain kernels) for pose, three planes (6 products, ~fp32) for face, whose O(10) outputs
    leave no margin at two planes (tests/test_gpu_tc_arm.py).  0 selects the exact-fp32 FFMA arm."""
    import os
    v = getattr(args, "split_terms", None)
    if v is None and os.environ.get("A2P_SPLIT_TERMS"):
        v = int(os.environ["A2P_SPLIT_TERMS"])
    if v is None:
        v = 2 if args.data_format == "pose" else 3
    return int(v)


def get_model_args(args, split_type: str) -> dict:
    """utils/model_util.py:49-76."""
    if args.data_format == "face":
        nfeat, lfeat = 256, 512
    elif args.data_format == "pose":
        nfeat, lfeat = 104, 256
    else:
        raise ValueError(args.data_format)
    if not hasattr(args, "num_audio_layers"):
        args.num_audio_layers = 3
    return {
        "args": args, "nfeats": nfeat, "latent_dim": lfeat, "ff_size": 1024, "num_layers": args.layers,
        "num_heads": args.heads, "dropout": 0.1, "cond_feature_dim": 512 * 2, "activation": F.gelu,
        "use_rotary": not args.not_rotary, "cond_mode": "uncond" if args.unconstrained else "audio",
        "split_type": split_type, "num_audio_layers": args.num_audio_layers, "device": args.device,
        "split_terms": default_split_terms(args),
    }


def create_model_and_diffusion(args, split_type: str):
    """utils/model_util.py:41-46 -> (Denoiser, Sampler)."""
    model = Denoiser(**get_model_args(args, split_type=split_type)).to(torch.float32)
    return model, create_gaussian_diffusion(args)


def load_model(model, state_dict) -> None:
    """utils/model_util.py:30-38: no unexpected keys; only guide transformer / tokenizer may be missing."""
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert len(unexpected) == 0, unexpected
    assert all(k.startswith(("transformer.", "tokenizer.", "audio_model.", "lip_model.")) for k in missing), missing


_patched = []   # (module, attribute, original) triples of the last patch_reference()


def patch_reference() -> None:
    """Swap the B200 classes into an importable reference checkout (it must be on sys.path):
    afterwards `python -m sample.generate ...` builds Denoiser / CFGDenoiser / Sampler.  `unpatch_reference()` undoes it."""
    import utils.model_util as mu  # reference module
    import model.cfg_sampler as cs

    def swap(mod, name, new):
        if getattr(mod, name, None) is not new:
            _patched.append((mod, name, getattr(mod, name)))
            setattr(mod, name, new)

    swap(mu, "create_model_and_diffusion", create_model_and_diffusion)
    swap(mu, "load_model", load_model)
    swap(mu, "create_gaussian_diffusion", create_gaussian_diffusion)
    swap(cs, "ClassifierFreeSampleModel", CFGDenoiser)
    try:
        import sample.generate as gen
    except Exception:
        return
    swap(gen, "create_model_and_diffusion", create_model_and_diffusion)
    swap(gen, "load_model", load_model)
    swap(gen, "ClassifierFreeSampleModel", CFGDenoiser)


def unpatch_reference() -> None:
    while _patched:
        mod, name, orig = _patched.pop()
        setattr(mod, name, orig)


__all__ = ["Denoiser", "CFGDenoiser", "Sampler", "create_model_and_diffusion", "create_gaussian_diffusion",
           "load_model", "get_model_args", "patch_reference", "unpatch_reference"]
