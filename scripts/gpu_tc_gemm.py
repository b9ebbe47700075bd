"""On-GPU unit test + timing of the split-bf16 tcgen05 GEMM against an fp64 torch reference."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio2photoreal_b200 import _lib

lib = _lib.load_testing()
vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
lib.a2p_test_tc_gemm_scratch_bytes.argtypes = [i32] * 4
lib.a2p_test_tc_gemm_scratch_bytes.restype = sz
lib.a2p_test_tc_gemm.argtypes = [i32] * 6 + [vp, vp, vp, vp, vp, sz, i32, C.POINTER(C.c_float), vp]
lib.a2p_test_sgemm.argtypes = [i32] * 5 + [vp, vp, vp, vp, i32, C.POINTER(C.c_float), vp]

def run(terms, M, N, K, taps=1, dil=0, iters=10):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(taps, N, K, device="cuda", generator=g) / (K ** 0.5)
    bias = torch.randn(N, device="cuda", generator=g)
    Cc = torch.full((M, N), float("nan"), device="cuda")
    nb = lib.a2p_test_tc_gemm_scratch_bytes(M, N, K, taps)
    scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
    ms = C.c_float()
    st = torch.cuda.current_stream().cuda_stream
    _lib.check_testing(lib.a2p_test_tc_gemm(terms, M, N, K, taps, dil, A.data_ptr(), W.data_ptr(), bias.data_ptr(), Cc.data_ptr(),
                                    scratch.data_ptr(), nb, iters, C.byref(ms), st))
    Ad, Wd = A.double(), W.double()
    ref = bias.double().expand(M, N).clone()
    for j in range(taps):
        sh = (taps - 1 - j) * dil
        Ash = torch.zeros_like(Ad)
        if sh == 0:
            Ash = Ad
        else:
            Ash[sh:] = Ad[:-sh]
        ref += Ash @ Wd[j].T
    err = (Cc.double() - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    fl = 2.0 * M * N * K * taps
    # fp32 FFMA comparison
    C2 = torch.empty(M, N, device="cuda")
    Ws = W.permute(1, 0, 2).reshape(N, taps * K).contiguous()
    ms2 = C.c_float()
    _lib.check_testing(lib.a2p_test_sgemm(M, N, K, taps, dil, A.data_ptr(), Ws.data_ptr(), bias.data_ptr(), C2.data_ptr(), iters, C.byref(ms2), st))
    err2 = (C2.double() - ref).abs().max().item()
    print(f"terms={terms} M={M} N={N} K={K} taps={taps}: tc max|d|={err:.3e} (rel {rel:.2e}) {ms.value*1e3:.1f}us "
          f"{fl/ms.value/1e9:.1f} TF/s alg | ffma max|d|={err2:.3e} {ms2.value*1e3:.1f}us {fl/ms2.value/1e9:.1f} TF/s", flush=True)
    return rel

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    run(1, 128, 128, 64, iters=1)          # single tile, single k-block
    run(1, 256, 256, 256)
    run(2, 256, 256, 256)
    run(3, 256, 256, 256)
    run(3, 200, 104, 104)                  # ragged M, N=104, K tail (zero fill)
    run(3, 300, 104, 104, taps=3, dil=2)   # conv taps
    for terms in (1, 2, 3):
        run(terms, 9600, 768, 256)
        run(terms, 9600, 1024, 256)
        run(terms, 9600, 256, 1024)
        run(terms, 38400, 256, 256)
    print("DONE")
