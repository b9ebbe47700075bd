"""On-GPU unit test + timing of the split-bf16 tcgen05 attention against fp64 torch and the FFMA kernel."""
import ctypes as C, os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio2photoreal_b200 import _lib

lib = _lib.load_testing()
vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
lib.a2p_test_tc_attention_scratch_bytes.argtypes = [i32] * 5
lib.a2p_test_tc_attention_scratch_bytes.restype = sz
lib.a2p_test_tc_attention.argtypes = [i32] * 7 + [vp] * 7 + [sz, i32, C.POINTER(C.c_float), vp]
lib.a2p_test_simt_attention.argtypes = [i32] * 6 + [vp] * 6 + [i32, C.POINTER(C.c_float), vp]

def run(terms, R, T, D, dh, S, nx, iters=5, qscale=1.0):
    g = torch.Generator(device="cuda").manual_seed(R * 1000 + T + S)
    Q = torch.randn(R, T, D, device="cuda", generator=g) * qscale
    K = torch.randn(R, S, D, device="cuda", generator=g)
    V = torch.randn(R, S, D, device="cuda", generator=g)
    Kx = torch.randn(R, max(nx, 1), D, device="cuda", generator=g)
    Vx = torch.randn(R, max(nx, 1), D, device="cuda", generator=g)
    O = torch.full((R, T, D), float("nan"), device="cuda")
    nb = lib.a2p_test_tc_attention_scratch_bytes(R, T, D, S, nx)
    scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
    ms = C.c_float()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check_testing(lib.a2p_test_tc_attention(terms, R, T, D, dh, S, nx, Q.data_ptr(), K.data_ptr(), V.data_ptr(), Kx.data_ptr(),
                                         Vx.data_ptr(), O.data_ptr(), scratch.data_ptr(), nb, iters, C.byref(ms), st))
    H = D // dh
    Kf = torch.cat([K, Kx[:, :nx]], 1) if nx else K
    Vf = torch.cat([V, Vx[:, :nx]], 1) if nx else V
    sp = lambda t: t.double().view(R, -1, H, dh).transpose(1, 2)
    att = torch.softmax(sp(Q) @ sp(Kf).transpose(-1, -2) / math.sqrt(dh), -1) @ sp(Vf)
    ref = att.transpose(1, 2).reshape(R, T, D)
    err = (O.double() - ref).abs().max().item()
    O2 = torch.empty_like(O)
    ms2 = C.c_float()
    _lib.check_testing(lib.a2p_test_simt_attention(R, T, D, dh, S, nx, Q.data_ptr(), K.data_ptr(), V.data_ptr(), Kx.data_ptr(), Vx.data_ptr(),
                                           O2.data_ptr(), iters, C.byref(ms2), st))
    err2 = (O2.double() - ref).abs().max().item()
    fl = 4.0 * R * T * (S + nx) * D
    print(f"terms={terms} R={R} T={T} D={D} dh={dh} S={S}+{nx}: tc max|d|={err:.3e} {ms.value*1e3:.1f}us {fl/ms.value/1e9:.1f} TF/s alg | "
          f"ffma max|d|={err2:.3e} {ms2.value*1e3:.1f}us {fl/ms2.value/1e9:.1f} TF/s", flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "small"):
        run(2, 1, 128, 64, 32, 64, 0, iters=1)       # one CTA, one key block, one group (2 heads)
        run(2, 1, 128, 64, 64, 64, 0, iters=1)       # dh = 64
        run(2, 1, 128, 64, 32, 200, 0, iters=1)      # several key blocks + ragged tail
        run(2, 2, 100, 256, 32, 77, 2, iters=1)      # ragged T, extra keys
        run(3, 2, 100, 256, 32, 77, 2, iters=1)
        run(1, 2, 100, 256, 32, 77, 2, iters=1)
        run(2, 2, 200, 512, 64, 211, 2, iters=1)     # face geometry
    if which in ("all", "big"):
        for terms in (2, 3):
            run(terms, 16, 600, 256, 32, 1998, 2)      # audio cross-attention, B=8 CFG
            run(terms, 16, 600, 256, 32, 600, 0)       # self-attention
            run(terms, 16, 600, 256, 32, 20, 0)        # keyframe cross-attention
            run(terms, 4, 600, 512, 64, 1998, 2)       # face
        run(2, 16, 600, 256, 32, 1998, 2, qscale=4.0)  # peaky softmax
    if which == "attn2poly":
        for terms in (21, 22, 23):     # 22 / 23: 1 / 2 of every 4 exponentials on the FMA pipe
            run(terms, 16, 600, 256, 32, 1998, 2)
            run(terms, 16, 600, 256, 32, 600, 0)
    if which == "sk":              # A/B of attention2 variant 8 (terms 27) against the default (terms 21)
        lib.a2p_test_attn2_set_persist.argtypes = [C.c_int]
        lib.a2p_test_attn2_set_persist(1)
        for terms in (21, 27, 21, 27):
            for R, S, nx in ((4, 1998, 2), (4, 600, 0), (8, 1998, 2), (32, 1998, 2), (32, 600, 0)):
                run(terms, R, 600, 256, 32, S, nx)
        run(27, 16, 600, 256, 32, 1998, 2, qscale=4.0)
    if which == "lo":              # lo-plane variants of the probability split: 21 rounding adds (default), 25 truncation, 26 cvt.rn.bf16x2
        for terms in (21, 25, 26, 21, 25, 26):
            for R, S, nx in ((4, 1998, 2), (4, 600, 0), (32, 1998, 2), (32, 600, 0)):
                run(terms, R, 600, 256, 32, S, nx)
        run(25, 16, 600, 256, 32, 1998, 2, qscale=4.0)
        run(26, 16, 600, 256, 32, 1998, 2, qscale=4.0)
    if which == "persist":         # persistent-CTA schedule A/B at the loop's launch shapes
        lib.a2p_test_attn2_set_persist.argtypes = [C.c_int]
        for on in (0, 1):
            lib.a2p_test_attn2_set_persist(on)
            print("persist", on)
            for R, S, nx in ((4, 1998, 2), (4, 600, 0), (8, 1998, 2), (8, 600, 0), (32, 1998, 2), (32, 600, 0)):
                run(21, R, 600, 256, 32, S, nx)
    if which == "decouple":        # attention2 timings at the loop's launch shapes (used for the r02 per-head MMA scheduling A/B, profiles/experiments/)
        for R, S, nx in ((4, 1998, 2), (4, 600, 0), (32, 1998, 2), (32, 600, 0), (2, 77, 2), (1, 200, 0)):
            T = 600 if R > 2 else 100
            run(21, R, T, 256, 32, S, nx)
    if which == "prof":
        run(int(sys.argv[2]), 16, 600, 256, 32, 1998, 2, iters=1)
    if which in ("all", "attn2"):
        # second-generation kernel (umma_attention2.cuh): terms 20 = P planes in shared memory, 21 = in tensor memory
        for terms in (20, 21):
            run(terms, 1, 128, 64, 32, 64, 0, iters=1)
            run(terms, 1, 128, 64, 32, 200, 0, iters=1)
            run(terms, 2, 100, 256, 32, 77, 2, iters=1)
            run(terms, 16, 600, 256, 32, 1998, 2)
            run(terms, 16, 600, 256, 32, 600, 0)
            run(terms, 16, 600, 256, 32, 20, 0)
            run(terms, 16, 600, 256, 32, 1998, 2, qscale=4.0)
        run(2, 16, 600, 256, 32, 1998, 2)
    print("DONE")
