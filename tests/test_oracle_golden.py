"""The oracle (oracle/a2p_oracle.py) against the REFERENCE's outputs committed under tests/golden/.
CPU only; sized to finish in about a minute."""
import os

import numpy as np
import pytest
import torch

from oracle import a2p_oracle as O
from oracle.cases import CASES, layer_inputs, make_inputs, weights_of


def _check(got, ref, atol=2e-5, rtol=1e-5):
    ref = torch.as_tensor(ref).double()
    scale = max(1.0, ref.abs().max().item())     # fp32-vs-fp32 noise floor scales with the output magnitude
    assert torch.allclose(got.double(), ref, atol=atol * scale, rtol=rtol), (got.double() - ref).abs().max().item()


@pytest.mark.parametrize("name", ["pose_small", "pose_small_h4", "face_small"])
def test_forward_matches_reference(golden_dir, name):
    case, g = CASES[name], np.load(os.path.join(golden_dir, f"fwd_{name}.npz"))
    inp, sd = make_inputs(case), weights_of(case)
    a = (sd, case.fmt, case.H, inp["x"], inp["times"], inp["feats"], inp["keyframes"], inp["mask"])
    _check(O.denoiser_forward(*a, 0.0), g["cond"])
    _check(O.denoiser_forward(*a, 1.0), g["uncond"])
    _check(O.cfg_forward(*a, inp["scale"]), g["cfg"], atol=2e-4)


@pytest.mark.parametrize("name", ["pose_small", "face_small"])
def test_decoder_layer_matches_reference(golden_dir, name):
    """one FiLMTransformerDecoderLayer (transformer_modules.py:190-217) of the reference on fixed (x, mem, t, mem2)"""
    import hashlib
    case, g = CASES[name], np.load(os.path.join(golden_dir, f"layer_{name}.npz"))
    x, mem, t, mem2 = layer_inputs(case)
    assert hashlib.sha1(x.numpy().tobytes()).hexdigest() == str(g["x_sha1"])
    _check(O.decoder_layer(x, mem, t, mem2, weights_of(case), "seqTransDecoder.stack.1", case.H), g["out"])


def test_forward_full_size_pose(golden_dir):
    case, g = CASES["pose_full"], np.load(os.path.join(golden_dir, "fwd_pose_full.npz"))
    inp, sd = make_inputs(case), weights_of(case)
    _check(O.denoiser_forward(sd, case.fmt, case.H, inp["x"], inp["times"], inp["feats"], inp["keyframes"], inp["mask"], 0.0),
           g["cond"])


@pytest.mark.parametrize("name,resp,kind,eta", [("pose_small", "ddim10", "ddim", 0.0), ("pose_small", "ddim10", "ddim", 0.5),
                                                ("pose_small", "10", "ancestral", 0.0), ("face_small", "ddim10", "ddim", 0.0)])
def test_loops_match_reference(golden_dir, name, resp, kind, eta):
    case = CASES[name]
    od = O.OracleDiffusion(resp)
    inp, sd = make_inputs(case, n_noise=od.num_timesteps), weights_of(case)
    fn = lambda x, ts: O.cfg_forward(sd, case.fmt, case.H, x, ts, inp["feats"], inp["keyframes"], inp["mask"], inp["scale"])
    tag = f"{kind}_{name}_{resp}" + (f"_eta{eta}" if eta else "")
    ref = np.load(os.path.join(golden_dir, f"loop_{tag}.npz"))["result"]
    got = od.ddim_sample_loop(fn, inp["x"], eta=eta, noise_tape=inp["noise_tape"]) if kind == "ddim" else \
        od.p_sample_loop(fn, inp["x"], inp["noise_tape"])
    _check(got, ref, atol=3e-4, rtol=1e-4)


@pytest.mark.parametrize("kind", ["ddim", "ancestral"])
def test_sampler_variants_match_reference(golden_dir, kind):
    """clip_denoised / skip_timesteps / init_image (DDIM) and const_noise + clip + skip (ancestral) against the
    reference's own loops (oracle/make_golden.py golden_loop_variants)."""
    from oracle.cases import variant_inputs
    case, resp, inp, sd, skip, init = variant_inputs(kind)
    od = O.OracleDiffusion(resp)
    fn = lambda x, ts: O.cfg_forward(sd, case.fmt, case.H, x, ts, inp["feats"], inp["keyframes"], inp["mask"], inp["scale"])
    ref = np.load(os.path.join(golden_dir, "loop_variants_pose_small.npz"))[kind]
    if kind == "ddim":
        got = od.ddim_sample_loop(fn, inp["x"], clip_denoised=True, skip_timesteps=skip, init_image=init)
    else:
        got = od.p_sample_loop(fn, inp["x"], inp["noise_tape"], const_noise=True, clip_denoised=True, skip_timesteps=skip)
    _check(got, ref, atol=3e-4, rtol=1e-4)


@pytest.mark.parametrize("order", [2, 4])
def test_plms_matches_reference(golden_dir, order):
    """plms_sample_loop (gaussian_diffusion.py:938-1158) restated by the oracle vs the reference's output"""
    case = CASES["pose_small"]
    od = O.OracleDiffusion("ddim10")
    inp, sd = make_inputs(case), weights_of(case)
    fn = lambda x, ts: O.cfg_forward(sd, case.fmt, case.H, x, ts, inp["feats"], inp["keyframes"], inp["mask"], inp["scale"])
    ref = np.load(os.path.join(golden_dir, "loop_plms_pose_small.npz"))[f"order{order}"]
    _check(od.plms_sample_loop(fn, inp["x"], order=order), ref, atol=3e-4, rtol=1e-4)


def test_oracle_fp64_agrees_with_fp32():
    case = CASES["pose_small"]
    inp, sd = make_inputs(case), weights_of(case)
    a = (sd, case.fmt, case.H, inp["x"], inp["times"], inp["feats"], inp["keyframes"], inp["mask"], 0.0)
    _check(O.denoiser_forward(*a, dtype=torch.float64).float(), O.denoiser_forward(*a).numpy(), atol=2e-5)
