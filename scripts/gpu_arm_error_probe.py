"""Where does the tensor-core arms' error come from?  One CFG forward of the pose_full_b4 case at three timesteps on the
exact-fp32 arm (reference for this probe) and on the split arms with components switched (env knobs are read at handle
creation / first use, so every variant runs in a fresh process: this script re-invokes itself)."""
import os, subprocess, sys
from argparse import Namespace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

VARIANTS = [("exact", 0, {}), ("P2 default", 2, {}), ("P2 unfused GEMMs (A2P_NO_CHAIN)", 2, {"A2P_NO_CHAIN": "1"}),
            ("P2 first-gen attention (A2P_ATTN2=0)", 2, {"A2P_ATTN2": "0"}), ("P2 no split-KV", 2, {"A2P_NO_SPLIT_KV": "1"}),
            ("P2 no short-attn", 2, {"A2P_NO_ATTN_SHORT": "1"}), ("P3", 3, {})]

if len(sys.argv) > 1:
    import numpy as np, torch
    from oracle.cases import CASES, make_inputs, weights_of
    from audio2photoreal_b200.api import CFGDenoiser, create_model_and_diffusion, load_model
    idx = int(sys.argv[1])
    name, terms, _ = VARIANTS[idx]
    case = CASES["pose_full_b4"]
    inp = make_inputs(case)
    args = Namespace(data_format="pose", add_frame_cond=1, max_seq_length=600, layers=case.L, heads=case.H, not_rotary=False,
                     unconstrained=False, device="cuda", timestep_respacing="", noise_schedule="cosine", sigma_small=True,
                     lambda_vel=0.0, model_path="x", resume_trans=None, split_terms=terms)
    model, _ = create_model_and_diffusion(args, "test")
    load_model(model, weights_of(case))
    model = model.cuda().eval()
    cfg = CFGDenoiser(model)
    y = {"audio_embed": inp["feats"].cuda(), "keyframes": inp["keyframes"].clone(), "mask": inp["mask"], "scale": inp["scale"].cuda()}
    outs = []
    for t in (900, 500, 10):
        g = torch.Generator().manual_seed(t)
        x = (torch.randn(case.B, 104, 1, case.T, generator=g) * (1.0 if t > 100 else 0.3)).cuda()
        outs.append(cfg(x, torch.full((case.B,), t, device="cuda"), y).double().cpu().numpy())
    np.save(f"/tmp/probe_{idx}.npy", np.stack(outs))
    sys.exit(0)

import numpy as np
for i, (name, terms, env) in enumerate(VARIANTS):
    e = dict(os.environ); e.update(env)
    r = subprocess.run([sys.executable, __file__, str(i)], env=e, capture_output=True, text=True)
    if r.returncode != 0:
        print(name, "FAILED", r.stderr[-400:])
ref = np.load("/tmp/probe_0.npy")
print(f"reference = exact-fp32 arm; |out| max {np.abs(ref).max():.3f}; columns per timestep (900, 500, 10): max|d|, rms d, mean signed d")
for i, (name, terms, env) in enumerate(VARIANTS[1:], 1):
    try:
        o = np.load(f"/tmp/probe_{i}.npy")
    except Exception:
        continue
    d = o - ref
    print(f"{name:42s} " + "  ".join(f"{np.abs(d[k]).max():.2e} {np.sqrt((d[k] ** 2).mean()):.2e} {d[k].mean():+.1e}" for k in range(3)))
