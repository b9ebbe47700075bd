#!/usr/bin/env python
"""bench.py -- motion-frames/sec of the full reverse-diffusion sampling loop (BASELINE.json metric).

One "step" = ONE pass of the hot path over one batch = one full sampling loop:
    body (pose) diffusion, 1000 steps (timestep_respacing ''), T = 600 frames, 104-dim pose,
    batch 8 per GPU (BASELINE configs[1]), CFG guidance 2.0 (sample/generate.py always wraps the
    model in ClassifierFreeSampleModel), random-init denoiser + synthetic wav2vec features.
value   = N*B*T / t_loop with inputs already resident in HBM (conditioning precompute is inside the loop time)
e2e     = same metric through the public API (Sampler.ddim_sample_loop) with HOST (pinned) inputs and a
          device->host read of the result inside the timed region
Scaling is weak: every rank owns B independent rows (no data-path collective; one all-gather of the result).

  python bench.py [--gpus N --steps K --warmup W]         this framework
  python bench.py --impl reference ...                      the reference's algorithm on the host CPU cores
                                                            (oracle port; /root/reference cannot travel)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from argparse import Namespace

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "motion-frames/sec full p_sample_loop (body, 1000 steps, T=600)"
UNIT = "frames/s"
WORKLOAD = dict(fmt="pose", layers=6, heads=8, T=600, S=1998, C=104, B=8, guidance=2.0, respacing="")


SPLIT_TERMS = 2


def model_args(respacing):
    return Namespace(split_terms=SPLIT_TERMS, data_format="pose", add_frame_cond=1, max_seq_length=600, layers=WORKLOAD["layers"],
                     heads=WORKLOAD["heads"], not_rotary=False, unconstrained=False, device="cuda",
                     timestep_respacing=respacing, noise_schedule="cosine", sigma_small=True, lambda_vel=0.0,
                     model_path="synthetic", resume_trans=None)


def synth_inputs(B, T, S, seed, pin=False):
    g = torch.Generator().manual_seed(seed)
    y = {
        "audio_embed": torch.randn(B, S, 1024, generator=g),
        "keyframes": torch.randn(B, len(range(0, T, 30)), 104, generator=g),
        "mask": torch.ones(B, 1, 1, T, dtype=torch.bool),
        "scale": torch.full((B,), WORKLOAD["guidance"]),
    }
    noise = torch.randn(B, WORKLOAD["C"], 1, T, generator=g)
    if pin and torch.cuda.is_available():
        y = {k: v.pin_memory() for k, v in y.items()}
        noise = noise.pin_memory()
    return y, noise


def flops_per_sample_forward(T=600, S=2000, S2=20, D=256, L=6, FF=1024, C=104):
    """SURVEY.md 8d formulas (2*MAC, cached-K/V convention)."""
    sa = 6 * T * D * D + 4 * T * T * D + 2 * T * D * D
    ca = 2 * T * D * D + 8 * D * D + 4 * T * S * D + 2 * T * D * D
    ca2 = 2 * T * D * D + 4 * T * S2 * D + 2 * T * D * D
    ffn = 4 * T * D * FF
    film = 4 * 4 * D * D
    io = 4 * T * C * D
    lens = [T + 24 - 2, T + 24 - 6, T + 24 - 12, T + 24 - 14, T + 24 - 18, T]
    ch = [(104, 256), (256, 104), (104, 104), (104, 104), (104, 104), (104, 104)]
    conv = sum(2 * ln * ci * co * 3 for ln, (ci, co) in zip(lens, ch)) + 2 * T * C * C
    return L * (sa + ca + ca2 + ffn + film) + io + conv + 32 * D * D


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


CPU_SAMPLE_STEPS = 10  # diffusion steps the in-line CPU arm times (about 10 s on 32 threads)
CPU_SAMPLE_ROWS = 2   # rows of the B=8 batch the CPU arm actually evaluates (per-row cost is independent of the batch)


def cpu_port_frames_per_s(n_diff_steps, threads):
    """The reference's algorithm on the host CPU (oracle port) on a BOUNDED sample of the same workload:
    `n_diff_steps` CFG diffusion steps of CPU_SAMPLE_ROWS of the 8 rows (T=600, S=1998, conditioning recomputed
    every call exactly like the reference), scaled to 8 rows x 1000 steps."""
    from oracle import a2p_oracle as O
    from audio2photoreal_b200.weights import model_dims, synthetic_state_dict
    torch.set_num_threads(threads)
    w = WORKLOAD
    sd = synthetic_state_dict(model_dims("pose", w["layers"], w["heads"]), seed=1)
    y, noise = synth_inputs(CPU_SAMPLE_ROWS, w["T"], w["S"], seed=10)
    fn = lambda x, ts, n: O.cfg_forward(sd, "pose", w["heads"], x[:n], ts[:n], y["audio_embed"][:n], y["keyframes"][:n],
                                        y["mask"][:n], y["scale"][:n])
    ts = torch.full((CPU_SAMPLE_ROWS,), 500, dtype=torch.long)
    with torch.no_grad():
        fn(noise, ts, 1)  # warm-up: one CFG call on one row
        t0 = time.perf_counter()
        x = noise
        for i in range(n_diff_steps):
            out = fn(x, ts - i, CPU_SAMPLE_ROWS)
            x = 0.9 * x + 0.1 * out.permute(0, 2, 1).unsqueeze(2)   # sampler arithmetic is negligible next to the model
        dt = time.perf_counter() - t0
    per_step_full_batch = dt / n_diff_steps * (w["B"] / CPU_SAMPLE_ROWS)
    return w["B"] * w["T"] / (per_step_full_batch * 1000), dt


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = min(os.cpu_count() or 1, 32)   # torch CPU GEMMs of this size stop scaling (and regress) beyond ~32 threads
    K_diff = 5
    vals = []
    for _ in range(max(1, a.steps)):
        v, dt = cpu_port_frames_per_s(K_diff, threads)
        vals.append(v)
    v = float(np.median(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * WORKLOAD["B"] * WORKLOAD["T"] / v, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "pose ddim/1000-step CFG g=2.0 B=8 T=600 L=6 D=256 (BASELINE configs[1])"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{K_diff} of 1000 diffusion steps of {CPU_SAMPLE_ROWS} of the 8 batch rows (CFG, T=600, S=1998), "
                                   f"scaled x1000/{K_diff} x 8/{CPU_SAMPLE_ROWS}; torch CPU fp32, conditioning recomputed per call "
                                   f"like the reference"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="a2p", choices=["a2p", "reference"])
    ap.add_argument("--diffusion-steps", type=int, default=1000, help="debug only; anything but 1000 is not the benchmark")
    ap.add_argument("--batch", type=int, default=WORKLOAD["B"], help="debug only; per-GPU batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cfg", action="store_true",
                    help="secondary measurement (SURVEY 8d, config 2 'no-CFG variant'): the bare denoiser, one forward per step; "
                         "not the benchmark line")
    ap.add_argument("--dump-out", default=None, help="debug only: save the result of the last timed loop as .npy")
    ap.add_argument("--split-terms", type=int, default=2, help="0: exact-fp32 FFMA arm; 2 (default, fused chain kernels) | 3: split-bf16 tcgen05 arms")
    a = ap.parse_args()
    global SPLIT_TERMS
    SPLIT_TERMS = a.split_terms
    if a.impl == "reference":
        return run_reference_arm(a)

    import torch.distributed as dist
    from audio2photoreal_b200 import _lib
    from audio2photoreal_b200.api import CFGDenoiser, create_model_and_diffusion, load_model
    from audio2photoreal_b200.dist import all_gather_rows
    from audio2photoreal_b200.weights import synthetic_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOAD
    B, T, S = a.batch, w["T"], w["S"]
    resp = "" if a.diffusion_steps == 1000 else f"ddim{a.diffusion_steps}"
    model, sampler = create_model_and_diffusion(model_args(resp), "test")
    load_model(model, synthetic_state_dict(model.dims, seed=1))
    model = model.to(dev).eval()
    cfg = model if a.no_cfg else CFGDenoiser(model)
    nbr = 1 if a.no_cfg else 2          # denoiser evaluations per row and step
    n_diff = sampler.num_timesteps
    shape = (B, w["C"], 1, T)

    # per-rank inputs (rank-dependent seed: independent rows on every GPU = weak scaling)
    y_host, noise_host = synth_inputs(B, T, S, seed=10 + rank, pin=True)
    y_dev = {k: v.to(dev) for k, v in y_host.items()}
    noise_dev = noise_host.to(dev)

    def loop_resident():
        yy = dict(y_dev)
        model._cond_sig = None          # drop the conditioning cache: the one-time precompute is part of every loop
        res = sampler.ddim_sample_loop(cfg, shape, noise=noise_dev, clip_denoised=False, model_kwargs={"y": yy},
                                       advance_rng=False)
        return all_gather_rows(res, B * world)     # the one collective of the path (no-op at world size 1)

    def loop_e2e():
        yy = {k: v.to(dev, non_blocking=True) for k, v in y_host.items()}
        nz = noise_host.to(dev, non_blocking=True)
        model._cond_sig = None
        res = sampler.ddim_sample_loop(cfg, shape, noise=nz, clip_denoised=False, model_kwargs={"y": yy}, advance_rng=False)
        return all_gather_rows(res, B * world).cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, wall * 1e3) if fn is loop_e2e else ms    # e2e includes the host-side result copy
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        barrier()
        return t.item() / k, out

    for _ in range(a.warmup):
        loop_resident()
    launches0 = model.launch_count()
    clocks = ClockSampler(local)
    clocks.start()
    ms_step, out = timed(loop_resident, a.steps)
    clk = clocks.stop()
    launches = (model.launch_count() - launches0)
    if a.dump_out and rank == 0:
        np.save(a.dump_out, out.float().cpu().numpy())
    loop_e2e()
    ms_e2e, out_h = timed(loop_e2e, a.steps)
    assert out.shape[0] == B * world and torch.isfinite(out).all()

    frames = world * B * T
    value = frames / (ms_step / 1e3)
    e2e_value = frames / (ms_e2e / 1e3)
    h2d = sum(v.numel() * v.element_size() for v in y_host.values()) + noise_host.numel() * 4
    d2h = out_h.numel() * 4

    roofline, cpu_base = None, None
    if rank == 0:
        # ---- live per-kernel measurement (CUDA events around every launch of one denoiser evaluation)
        lib = _lib.load()
        ncat = 9
        ms_cat = (C.c_float * ncat)()
        n_cat = (C.c_int64 * ncat)()
        x_btc = torch.randn(B, T, w["C"], device=dev)
        ts = torch.full((B,), 500, device=dev, dtype=torch.int64)
        ws = model._workspace(lib.a2p_workspace_bytes(C.byref(model._cfg), B, T), dev)
        acc = np.zeros(ncat)
        reps = 5
        # the fused arm cuts a CFG step into concurrent forwards: {cond, uncond} x groups of batch rows (engine.cu,
        # sample_loop_impl).  Profile the launches the loop really makes -- every (branch, row group) forward, every launch
        # timed ALONE (in the loop the units overlap, so the per-kernel times add up to more than a step)
        groups = int(lib.a2p_loop_row_groups(model._handle, B, T))      # 0: one stacked forward for both branches
        units = [(3, 0, B)] if groups == 0 else [(mk, B * g // groups, B * (g + 1) // groups - B * g // groups)
                                                 for g in range(groups) for mk in (1, 2)]
        if a.no_cfg:
            groups, units = 0, [(1, 0, B)]
        two_branch = groups > 0
        rows_per_launch = nbr * B if groups == 0 else units[0][2]
        n_tot = np.zeros(ncat, dtype=np.int64)
        for i in range(reps + 1):
            for mk, b0, bs in units:
                _lib.check(lib.a2p_profile_forward_rows(model._handle, B, b0, bs, T, x_btc[b0:].data_ptr(), ts[b0:].data_ptr(), mk,
                                                        ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream,
                                                        ms_cat, n_cat, ncat))
                if i:
                    acc += np.array(list(ms_cat))
                if i == 1:
                    n_tot += np.array(list(n_cat), dtype=np.int64)
        acc /= reps
        n_cat = [int(v) for v in n_tot]
        names = ["cond_gemm", "ln_rope", "attn_proj_gemm", "attn_self", "attn_cross_audio", "attn_cross_keyframe", "ffn_gemm",
                 "io_tcn_gemm", "misc"]
        # kernel families: the three attention categories are launches of ONE kernel (umma_attn2_kernel), likewise the
        # chain categories; the dominant kernel is the family with the largest share of the step, reported through its
        # biggest category (the audio cross-attention launch / the FFN chain launches)
        fam_attn, fam_chain = acc[3] + acc[4] + acc[5], acc[2] + acc[6]
        dom = (4 if fam_attn >= fam_chain else int(np.argmax([0, 0, acc[2], 0, 0, 0, acc[6]]))) if SPLIT_TERMS == 2 else int(np.argmax(acc))
        R = nbr * B        # rows of one step (both branches); per-launch figures below divide by the launch counts
        D, L = 256, w["layers"]
        # per-launch algorithmic FLOPs of each category (attention cores exactly; linears = category total / launches)
        lin_proj = (6 * T * D * D + 2 * T * D * D + 4 * T * D * D + 4 * T * D * D) * R * L
        lin_ffn = 4 * T * D * 1024 * R * L
        if SPLIT_TERMS == 2:   # fused chain arm: PROJ = {sa_out+q, ca_out+q} per layer, FFN = {out+ffn1, ffn2+next qkv} per layer
            lin_proj, lin_ffn = 8 * T * D * D * R * L, (2 * T * D * D + 4 * T * D * 1024 + 6 * T * D * D) * R * L
        alg = {3: 4 * T * T * D * R * L / max(1, n_cat[3]), 4: 4 * T * (S + 2) * D * R * L / max(1, n_cat[4]),
               5: 4 * T * 20 * D * R * L / max(1, n_cat[5]),
               2: lin_proj / max(1, n_cat[2]), 6: lin_ffn / max(1, n_cat[6])}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops", 1590.0)
        peak_src = "measured" if "bf16_tflops" in peaks else "fallback"
        per_launch_ms = acc[dom] / max(1, n_cat[dom])
        flops_launch = alg.get(dom, 0)
        achieved = flops_launch / (per_launch_ms * 1e-3) / 1e12 if per_launch_ms > 0 else 0.0
        traffic = None     # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
            if tr.get("kernel") == names[dom] and B == WORKLOAD["B"] and tr.get("rows_per_launch", 2 * B) == rows_per_launch:
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        roofline = {"bound": "tensor", "kernel": names[dom], "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": achieved / peak_tf, "traffic": traffic, "peak_source": f"bf16_tflops burst, of {peak_src}",
                    "ms_per_launch": per_launch_ms, "launches_per_forward": int(n_cat[dom]),
                    "forward_ms_by_kernel": {n: round(float(v), 4) for n, v in zip(names, acc)},
                    "kernel_family_ms": {"attention(self+audio+keyframe)": round(float(fam_attn), 4), "chain(proj+ffn)": round(float(fam_chain), 4)},
                    "split_terms": SPLIT_TERMS,
                    "launch_shape": ("bare denoiser, no CFG: one forward of B rows per step" if a.no_cfg else f"{rows_per_launch} rows of one CFG branch per launch ({len(units)} concurrent forwards per step), each launch "
                                     "timed alone; in the loop the forwards overlap" if two_branch else "both CFG branches (2B rows) per launch"),
                    "note": ("algorithmic FLOPs (one product per MAC) over measured time; the split-bf16 arm spends %d tensor-core "
                             "products per MAC for fp32-level parity" % {0: 0, 1: 1, 2: 3, 3: 6}[SPLIT_TERMS])}
        if not a.no_cpu_baseline:
            threads = min(os.cpu_count() or 1, 32)
            v, dt = cpu_port_frames_per_s(CPU_SAMPLE_STEPS, threads)
            cpu_base = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": f"{CPU_SAMPLE_STEPS} of 1000 diffusion steps of {CPU_SAMPLE_ROWS} of the 8 batch rows (CFG, T=600, S=1998) "
                                  f"in {dt:.1f}s, scaled x1000/{CPU_SAMPLE_STEPS} x 8/{CPU_SAMPLE_ROWS}"}
        f_fwd = flops_per_sample_forward()
        # whole-step view beside the per-launch one: the concurrent forwards share the machine, so the step as a whole sustains
        # more than any single launch timed alone
        step_tf = f_fwd * nbr * B * n_diff / (ms_step * 1e-3) / 1e12
        roofline["step_level"] = {"achieved": step_tf, "frac": step_tf / peak_tf, "concurrent_forwards": len(units),
                                  "note": "algorithmic FLOPs of a whole loop / loop time on this GPU"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if SPLIT_TERMS == 0 else f"bf16x{SPLIT_TERMS} split (fp32-equivalent), fp32 accumulate",
            "data": "synthetic",
            "config": {"workload": f"pose body diffusion, {n_diff} steps, T={T}, C=104, batch {B}/GPU, {'NO CFG (bare denoiser)' if a.no_cfg else 'CFG g=2.0'}, "
                                   f"L=6 D=256 H=8, synthetic wav2vec features [B,{S},1024] (BASELINE configs[1])",
                       "global_batch": B * world, "parallelism": f"batch-sharded x{world}, 1 all-gather",
                       "l2": "inputs_larger_than_l2 (K/V caches %d MB + activations per step)" % (B * 25),
                       "algorithmic_gflop_per_loop": f_fwd * nbr * B * world * n_diff / 1e9},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e},
            "roofline": roofline, "cpu_baseline": cpu_base,
            "model_tflops": f_fwd * nbr * B * world * n_diff / (ms_step * 1e-3) / 1e12,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
