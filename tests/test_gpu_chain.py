"""-m gpu: the fused row-chain kernel (csrc/umma_chain.cuh) against an fp64 torch restatement of the same chain
(transformer_modules.py:190-217: out_proj + FiLM + residual -> LayerNorm -> RoPE -> in_proj; model/diffusion.py:364,397).
Error budget: split-bf16 x2 operands keep 16 mantissa bits (2^-17 relative per operand), so every stage must agree with
fp64 to a few 1e-5 of the output scale."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

CASES = {
    # name:        M     T    K0   N1   film ln rope gelu vjob scale_ncols
    "sa_out_q":   (9600, 600, 256, 256, 1, 1, 1, 0, 0, 0),
    "ffn2_qkv":   (9600, 600, 1024, 512, 1, 1, 1, 0, 1, 256),
    "inproj_qkv": (9600, 600, 104, 512, 0, 1, 1, 0, 1, 256),
    "out_ffn1":   (9600, 600, 256, 1024, 1, 1, 0, 1, 0, 0),
    "ffn2_final": (9600, 600, 1024, 104, 1, 0, 0, 0, 0, 0),
    "ragged":     (328, 164, 256, 256, 1, 1, 1, 0, 1, 0),   # last tile 72 rows, a sample boundary inside a tile
}


def _lib():
    from audio2photoreal_b200 import _lib
    lib = _lib.load_testing()
    vp, i32, sz, f32 = C.c_void_p, C.c_int, C.c_size_t, C.c_float
    lib.a2p_test_chain_scratch_bytes.argtypes = [i32] * 4
    lib.a2p_test_chain_scratch_bytes.restype = sz
    lib.a2p_test_chain.argtypes = [i32] * 9 + [f32, i32] + [vp] * 15 + [sz, i32, C.POINTER(f32), vp]
    lib.a2p_test_chain.restype = i32
    lib.a2p_test_chain_set_mode.argtypes = [i32]
    lib.a2p_test_chain_set_mode.restype = None
    lib.a2p_test_chain_set_nsplit.argtypes = [i32]
    lib.a2p_test_chain_set_nsplit.restype = None
    return _lib, lib


def run_case(name, iters=0, seed=0, mode=0, nsplit=0, M=None):
    """returns dict(stage -> (max abs err, max |ref|)) and the kernel time in ms (iters > 0)"""
    M_, T, K0, N1, film_mode, ln_mode, rope, gelu, vjob, scale_ncols = CASES[name]
    M = M or M_
    L, lib = _lib()
    lib.a2p_test_chain_set_mode(mode)      # 0 library default, 1 one CTA per tile, 2 CTA pairs (cta_group::2)
    lib.a2p_test_chain_set_nsplit(nsplit)  # > 0: every tile worked on by min(nsplit, accumulator halves) CTAs (CTA pairs)
    g = torch.Generator(device="cuda").manual_seed(seed)
    dev = "cuda"
    rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev) * sc).contiguous()
    A0, W0, b0 = rn(M, K0), rn(256, K0, sc=K0 ** -0.5), rn(256, sc=0.1)
    film = rn((M + T - 1) // T, 512, sc=0.3)
    x0 = rn(M, 256, sc=2.0) + 0.5
    lnw, lnb = 1.0 + rn(256, sc=0.1), rn(256, sc=0.1)
    freqs = (10000.0 ** (-torch.arange(0, 256, 2, device=dev).float() / 256)).contiguous()
    W1, b1 = rn(N1, 256, sc=1 / 16), rn(N1, sc=0.1)
    W2, b2 = rn(256, 256, sc=1 / 16), rn(256, sc=0.1)
    out_scale = 0.25504
    M8 = (M + 7) // 8 * 8
    Cp = torch.zeros(2, M, N1, device=dev, dtype=torch.bfloat16)
    Vt = torch.zeros(2, 256, M8, device=dev, dtype=torch.bfloat16)
    nb = lib.a2p_test_chain_scratch_bytes(M, K0, N1, T)
    scratch = torch.zeros(nb, device=dev, dtype=torch.uint8)
    x = x0.clone()
    ms = C.c_float(0.0)
    st = torch.cuda.current_stream().cuda_stream
    call = lambda xx, it: L.check_testing(lib.a2p_test_chain(
        M, T, K0, N1, film_mode, ln_mode, rope, gelu, vjob, out_scale, scale_ncols, A0.data_ptr(), W0.data_ptr(), b0.data_ptr(),
        film.data_ptr(), xx.data_ptr(), lnw.data_ptr(), lnb.data_ptr(), freqs.data_ptr(), W1.data_ptr(), b1.data_ptr(),
        W2.data_ptr(), b2.data_ptr(), Cp.data_ptr(), Vt.data_ptr(), scratch.data_ptr(), nb, it, C.byref(ms), st))
    call(x, 0)
    torch.cuda.synchronize()
    # ---- fp64 restatement
    d = torch.float64
    acc = A0.to(d) @ W0.to(d).T + b0.to(d)
    if film_mode:
        s_idx = torch.arange(M, device=dev) // T
        sc, sh = film[s_idx, :256].to(d), film[s_idx, 256:].to(d)
        xr = x0.to(d) + ((sc + 1) * acc + sh)
    else:
        xr = acc
    if ln_mode:
        mu = xr.mean(-1, keepdim=True)
        var = ((xr - mu) ** 2).mean(-1, keepdim=True)
        h = (xr - mu) / torch.sqrt(var + 1e-5) * lnw.to(d) + lnb.to(d)
    else:
        h = xr
    hr = h
    if rope:
        pos = (torch.arange(M, device=dev) % T).float()
        ang = (pos[:, None] * freqs[None, :]).to(d)          # fp32 product like the table kernel
        cs, sn = torch.cos(ang), torch.sin(ang)
        he, ho = h[:, 0::2], h[:, 1::2]
        hr = torch.stack([he * cs - ho * sn, ho * cs + he * sn], -1).reshape(M, 256)
    c1 = hr @ W1.to(d).T + b1.to(d)
    if gelu:
        c1 = torch.nn.functional.gelu(c1)
    else:
        ncol = scale_ncols if scale_ncols else N1
        c1[:, :ncol] *= out_scale
    res = {"x": ((x.to(d) - xr).abs().max().item(), xr.abs().max().item()),
           "Cp": ((Cp.to(d).sum(0) - c1).abs().max().item(), c1.abs().max().item())}
    if vjob:
        vt = (h @ W2.to(d).T + b2.to(d)).T
        res["Vt"] = ((Vt.to(d).sum(0)[:, :M] - vt).abs().max().item(), vt.abs().max().item())
    t = None
    if iters > 0:
        call(x0.clone(), iters)
        t = ms.value
    lib.a2p_test_chain_set_nsplit(0)
    lib.a2p_test_chain_set_mode(0)
    return res, t


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("name", list(CASES))
def test_chain_vs_fp64(name, mode):
    res, _ = run_case(name, mode=mode)
    for stage, (err, scale) in res.items():
        tol = (2e-5 if stage == "x" else 4e-5) * max(1.0, scale)
        assert err <= tol, f"{name}/{stage}: max|d|={err:.3e} (|ref|max={scale:.3f}, tol {tol:.1e})"


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("nsplit", [2, 3, 4])
@pytest.mark.parametrize("name", list(CASES))
def test_chain_n_split_vs_fp64(name, nsplit, mode):
    """N split (small batches): every tile is worked on by nsplit CTAs (CTA pairs) that share GEMM0 / E_A redundantly and own
    disjoint accumulator halves of GEMM1 / the V job -- incl. parts that hold only V halves, a V job cut between two parts
    (ragged: 2 + 2 halves over 3 parts) and the residual stream written to a second buffer."""
    res, _ = run_case(name, mode=mode, nsplit=nsplit, M=2400 if CASES[name][0] == 9600 else None)
    for stage, (err, scale) in res.items():
        tol = (2e-5 if stage == "x" else 4e-5) * max(1.0, scale)
        assert err <= tol, f"{name}/{stage} nsplit={nsplit}: max|d|={err:.3e} (|ref|max={scale:.3f}, tol {tol:.1e})"


def test_chain_n_split_is_bit_identical():
    """the parts of a tile repeat the same GEMM0 / E_A arithmetic: the split changes WHO computes a column, not its value"""
    import numpy as np
    outs = []
    for nsplit in (0, 2, 3):
        M, T, K0, N1, film_mode, ln_mode, rope, gelu, vjob, scale_ncols = CASES["ffn2_qkv"]
        M = 1200
        L, lib = _lib()
        lib.a2p_test_chain_set_mode(1)
        lib.a2p_test_chain_set_nsplit(nsplit)
        g = torch.Generator(device="cuda").manual_seed(3)
        rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device="cuda") * sc).contiguous()
        A0, W0, b0 = rn(M, K0), rn(256, K0, sc=K0 ** -0.5), rn(256, sc=0.1)
        film = rn((M + T - 1) // T, 512, sc=0.3)
        x = rn(M, 256, sc=2.0) + 0.5
        lnw, lnb = 1.0 + rn(256, sc=0.1), rn(256, sc=0.1)
        freqs = (10000.0 ** (-torch.arange(0, 256, 2, device="cuda").float() / 256)).contiguous()
        W1, b1 = rn(N1, 256, sc=1 / 16), rn(N1, sc=0.1)
        W2, b2 = rn(256, 256, sc=1 / 16), rn(256, sc=0.1)
        M8 = (M + 7) // 8 * 8
        Cp = torch.zeros(2, M, N1, device="cuda", dtype=torch.bfloat16)
        Vt = torch.zeros(2, 256, M8, device="cuda", dtype=torch.bfloat16)
        nb = lib.a2p_test_chain_scratch_bytes(M, K0, N1, T)
        scratch = torch.zeros(nb, device="cuda", dtype=torch.uint8)
        ms = C.c_float(0.0)
        L.check_testing(lib.a2p_test_chain(
            M, T, K0, N1, film_mode, ln_mode, rope, gelu, vjob, 0.25504, scale_ncols, A0.data_ptr(), W0.data_ptr(), b0.data_ptr(),
            film.data_ptr(), x.data_ptr(), lnw.data_ptr(), lnb.data_ptr(), freqs.data_ptr(), W1.data_ptr(), b1.data_ptr(),
            W2.data_ptr(), b2.data_ptr(), Cp.data_ptr(), Vt.data_ptr(), scratch.data_ptr(), nb, 0, C.byref(ms),
            torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        lib.a2p_test_chain_set_nsplit(0)
        lib.a2p_test_chain_set_mode(0)
        outs.append((x.cpu().numpy(), Cp.view(torch.int16).cpu().numpy(), Vt.view(torch.int16).cpu().numpy()))
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            assert np.array_equal(a, b)


if __name__ == "__main__":
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    names = [a for a in sys.argv[1:] if not a.startswith("--")] or list(CASES)
    for M in (9600, 2400, 1200):          # 32 / 8 / 4 sample rows per launch: 75 / 19 / 10 tiles
        for mode in (1, 2):
            for nsplit in ((0,) if M == 9600 else (0, 2, 3, 4)):
                for name in names:
                    if name == "ragged":
                        continue
                    res, t = run_case(name, iters=20, mode=mode, nsplit=nsplit, M=M)
                    print(f"M={M} mode{mode} nsplit{nsplit}", name, {k: f"{e:.2e}/{s:.2f}" for k, (e, s) in res.items()},
                          f"{1e3 * t:.1f} us" if t else "", flush=True)
