import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio2photoreal_b200 import _lib
# trace build of the testing library: python audio2photoreal_b200/csrc/build.py --trace
lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(_lib.__file__)), "liba2p_b200_trace.so"))
lib.a2p_test_last_error.restype = C.c_char_p
def _check(rc):
    if rc: raise RuntimeError(lib.a2p_test_last_error().decode())
vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
lib.a2p_test_tc_attention_scratch_bytes.argtypes = [i32] * 5
lib.a2p_test_tc_attention_scratch_bytes.restype = sz
lib.a2p_test_tc_attention.argtypes = [i32] * 7 + [vp] * 7 + [sz, i32, C.POINTER(C.c_float), vp]
lib.a2p_test_attn2_set_persist.argtypes = [C.c_int]
lib.a2p_test_attn2_set_persist(0)          # one CTA per work item: block indices of the trace are those of one tile
terms = int(sys.argv[1]) if len(sys.argv) > 1 else 2
R, T, D, dh, S, nx = 16, 600, 256, 32, 1998, 2
g = torch.Generator(device="cuda").manual_seed(0)
Q, K, V = (torch.randn(R, n, D, device="cuda", generator=g) for n in (T, S, S))
Kx, Vx = (torch.randn(R, 2, D, device="cuda", generator=g) for _ in range(2))
O = torch.zeros(R, T, D, device="cuda")
nb = lib.a2p_test_tc_attention_scratch_bytes(R, T, D, S, nx)
scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
ms = C.c_float()
_check(lib.a2p_test_tc_attention(terms, R, T, D, dh, S, nx, Q.data_ptr(), K.data_ptr(), V.data_ptr(), Kx.data_ptr(), Vx.data_ptr(),
                                     O.data_ptr(), scratch.data_ptr(), nb, -1, C.byref(ms), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
tr = O.view(-1)[: 64 * 32 * 2].view(torch.int64).view(64, 32).cpu()
t0 = tr[8, 0].item()
cols = [(0, "h0:top"), (1, "h0:s_full"), (2, "h0:ld+pv"), (3, "h0:max"), (4, "h0:exp+st"), (5, "h0:arr"),
        (16, "h1:top"), (17, "h1:s_full"), (18, "h1:ld+pv"), (19, "h1:max"), (20, "h1:exp+st"), (21, "h1:arr"),
        (8, "mma:top"), (9, "mma:kv"), (10, "mma:w_p0"), (11, "mma:p0"), (12, "mma:w_p1"), (13, "mma:p1")]
print(f"# A2P_ATTN_SKEW_NS={os.environ.get('A2P_ATTN_SKEW_NS', '0')}  cycles since head 0 entered block 8; one row per 64-key block")
print("iter " + " ".join(f"{n:>10}" for _, n in cols))
for i in range(8, 26):
    print(f"{i:4d} " + " ".join(f"{tr[i, k].item() - t0:10d}" for k, _ in cols))
for h, c in ((0, 0), (1, 16)):
    d = tr[12:28, c] - tr[11:27, c]
    print(f"head {h}: mean cycles per block {d.float().mean().item():.0f}")
print("head 1 lag behind head 0 at block top:", [int(tr[i, 16].item() - tr[i, 0].item()) for i in range(8, 26)])
