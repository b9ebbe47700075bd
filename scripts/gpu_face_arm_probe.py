"""face model: error of every arithmetic arm against the reference's golden outputs (fwd_face_full: L=8, D=512, T=600, S=1998, g=10;
loop_ddim_face_cfg1_ddim10: BASELINE configs[0] geometry)"""
import os, sys
from argparse import Namespace
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.cases import CASES, make_inputs, weights_of
from audio2photoreal_b200.api import CFGDenoiser, create_model_and_diffusion, load_model

def build(case, resp, terms):
    args = Namespace(data_format="face", add_frame_cond=None, max_seq_length=600, layers=case.L, heads=case.H, not_rotary=False,
                     unconstrained=False, device="cuda", timestep_respacing=resp, noise_schedule="cosine", sigma_small=True,
                     lambda_vel=0.0, model_path="x", resume_trans=None, split_terms=terms)
    model, sampler = create_model_and_diffusion(args, "test")
    load_model(model, weights_of(case))
    model = model.cuda().eval()
    return model, CFGDenoiser(model), sampler

def stats(got, ref, what):
    got, ref = got.double().cpu(), torch.as_tensor(ref).double()
    d = (got - ref).abs()
    bad = d > 1e-4 + 1e-3 * ref.abs()
    print(f"  {what:28s} max|d|={d.max():.3e} rms={d.pow(2).mean().sqrt():.3e} outside strict={bad.double().mean():.4%}  |ref|max={ref.abs().max():.2f}", flush=True)

for terms in (0, 3, 2):
    print(f"split_terms={terms}")
    case = CASES["face_full"]; g = np.load("tests/golden/fwd_face_full.npz"); inp = make_inputs(case)
    model, cfg, _ = build(case, "ddim10", terms)
    y = {"audio_embed": inp["feats"].cuda(), "keyframes": inp["keyframes"].clone(), "mask": inp["mask"], "scale": inp["scale"].cuda()}
    x, t = inp["x"].cuda(), inp["times"].cuda()
    stats(model(x, t, y, cond_drop_prob=0.0), g["cond"], "fwd_face_full cond")
    stats(model(x, t, y, cond_drop_prob=1.0), g["uncond"], "fwd_face_full uncond")
    stats(cfg(x, t, y), g["cfg"], "fwd_face_full cfg (g=10)")
    del model, cfg; torch.cuda.empty_cache()
    case = CASES["face_cfg1"]; inp = make_inputs(case)
    model, cfg, sampler = build(case, "ddim10", terms)
    y = {"audio_embed": inp["feats"].cuda(), "keyframes": inp["keyframes"].clone(), "mask": inp["mask"], "scale": inp["scale"].cuda()}
    ref = np.load("tests/golden/loop_ddim_face_cfg1_ddim10.npz")["result"]
    stats(sampler.ddim_sample_loop(cfg, tuple(inp["x"].shape), noise=inp["x"].cuda(), clip_denoised=False, model_kwargs={"y": y}), ref, "loop face_cfg1 ddim10")
    del model, cfg; torch.cuda.empty_cache()
