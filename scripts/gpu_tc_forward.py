"""TC arm (split_terms 2/3) forward + ddim10 parity vs golden, and per-kernel timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from argparse import Namespace
from audio2photoreal_b200.api import create_model_and_diffusion, load_model, CFGDenoiser
from oracle.cases import CASES, make_inputs, weights_of
GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

def args_of(case, resp, terms):
    return Namespace(data_format=case.fmt, add_frame_cond=1 if case.fmt == "pose" else None, max_seq_length=600,
                     layers=case.L, heads=case.H, not_rotary=False, unconstrained=False, device="cuda",
                     timestep_respacing=resp, noise_schedule="cosine", sigma_small=True, lambda_vel=0.0,
                     model_path="x", resume_trans=None, split_terms=terms)

def report(tag, got, ref):
    got, ref = got.double().cpu(), torch.as_tensor(ref).double()
    d = (got - ref).abs()
    viol = (d > 1e-4 + 1e-3 * ref.abs()).double().mean().item()
    print(f"{tag}: max|d|={d.max().item():.3e} |ref|max={ref.abs().max().item():.3f} viol={100*viol:.4f}%", flush=True)

for terms in (3, 2):
    for name in ["pose_small", "face_small", "pose_full"]:
        case = CASES[name]
        inp = make_inputs(case)
        model, diff = create_model_and_diffusion(args_of(case, "ddim10", terms), "test")
        load_model(model, weights_of(case))
        model = model.cuda().eval()
        y = {"audio_embed": inp["feats"].cuda(), "keyframes": inp["keyframes"].clone(), "mask": inp["mask"], "scale": inp["scale"].cuda()}
        g = np.load(os.path.join(GOLD, f"fwd_{name}.npz"))
        x, t = inp["x"].cuda(), inp["times"].cuda()
        c = model(x, t, y, cond_drop_prob=0.0); report(f"terms={terms} {name}/cond", c, g["cond"])
        cfg = CFGDenoiser(model)
        o = cfg(x, t, y); report(f"terms={terms} {name}/cfg", o, g["cfg"])
        lp = os.path.join(GOLD, f"loop_ddim_{name}_ddim10.npz")
        if os.path.exists(lp):
            torch.cuda.synchronize(); t0 = time.time()
            res = diff.ddim_sample_loop(cfg, tuple(inp["x"].shape), noise=x, clip_denoised=False, model_kwargs={"y": y})
            torch.cuda.synchronize()
            report(f"terms={terms} {name}/ddim10 ({time.time()-t0:.2f}s)", res, np.load(lp)["result"])
print("DONE")
