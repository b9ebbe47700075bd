"""1000-step B=4 CFG loop (BASELINE configs[1] geometry) on every arithmetic arm against the reference's golden output:
error statistics per arm (max |d|, fraction outside rtol 1e-3 / atol 1e-4, percentiles) -> gpurun_out/r2_loop1000_arms.txt"""
import os, sys, time
from argparse import Namespace
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.cases import CASES, make_inputs, weights_of
from audio2photoreal_b200.api import CFGDenoiser, create_model_and_diffusion, load_model

NAME = sys.argv[1] if len(sys.argv) > 1 else "pose_full_b4"
case = CASES[NAME]
ref = torch.from_numpy(np.load(f"tests/golden/loop_ddim_{NAME}_full.npz")["result"]).double()
inp = make_inputs(case)
print(f"case {NAME}: guidance scales {inp['scale'].tolist()}")
out = {}
for terms in (0, 3, 2):
    args = Namespace(data_format="pose", add_frame_cond=1, max_seq_length=600, layers=case.L, heads=case.H, not_rotary=False,
                     unconstrained=False, device="cuda", timestep_respacing="", noise_schedule="cosine", sigma_small=True,
                     lambda_vel=0.0, model_path="x", resume_trans=None, split_terms=terms)
    model, sampler = create_model_and_diffusion(args, "test")
    load_model(model, weights_of(case))
    model = model.cuda().eval()
    cfg = CFGDenoiser(model)
    y = {"audio_embed": inp["feats"].cuda(), "keyframes": inp["keyframes"].clone(), "mask": inp["mask"], "scale": inp["scale"].cuda()}
    t0 = time.time()
    res = sampler.ddim_sample_loop(cfg, tuple(inp["x"].shape), noise=inp["x"].cuda(), clip_denoised=False, model_kwargs={"y": y})
    torch.cuda.synchronize()
    dt = time.time() - t0
    r = res.double().cpu()
    out[terms] = r
    d = (r - ref).abs()
    bad = d > 1e-4 + 1e-3 * ref.abs()
    q = torch.quantile(d.flatten()[:: 7].float(), torch.tensor([0.5, 0.99, 0.9999]))
    rows = [(d[b].max().item(), int(bad[b].sum())) for b in range(d.shape[0])]
    print(f"terms={terms}: per-row (max|d|, #outside): " + ", ".join(f"({a:.2e}, {n})" for a, n in rows))
    print(f"terms={terms}: {dt:.1f}s  max|d|={d.max():.3e}  outside={bad.double().mean():.5%} ({int(bad.sum())} of {bad.numel()})  "
          f"median/p99/p99.99 |d| = {q[0]:.2e}/{q[1]:.2e}/{q[2]:.2e}  |ref|max={ref.abs().max():.3f}", flush=True)
    del model, cfg, sampler
    torch.cuda.empty_cache()
d02 = (out[0] - out[2]).abs()
print(f"exact-fp32 arm vs split-2 arm: max|d|={d02.max():.3e}; exact vs split-3: {(out[0]-out[3]).abs().max():.3e}")
