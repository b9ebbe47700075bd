import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio2photoreal_b200 import _lib
lib = _lib.load_testing()
vp, i32, sz = C.c_void_p, C.c_int, C.c_size_t
lib.a2p_test_tc_attention_scratch_bytes.argtypes = [i32] * 5
lib.a2p_test_tc_attention_scratch_bytes.restype = sz
lib.a2p_test_tc_attention.argtypes = [i32] * 7 + [vp] * 7 + [sz, i32, C.POINTER(C.c_float), vp]
terms = int(sys.argv[1]) if len(sys.argv) > 1 else 2
R, T, D, dh, S, nx = 16, 600, 256, 32, 1998, 2
g = torch.Generator(device="cuda").manual_seed(0)
Q, K, V = (torch.randn(R, n, D, device="cuda", generator=g) for n in (T, S, S))
Kx, Vx = (torch.randn(R, 2, D, device="cuda", generator=g) for _ in range(2))
O = torch.zeros(R, T, D, device="cuda")
nb = lib.a2p_test_tc_attention_scratch_bytes(R, T, D, S, nx)
scratch = torch.empty(nb, dtype=torch.uint8, device="cuda")
ms = C.c_float()
_lib.check_testing(lib.a2p_test_tc_attention(terms, R, T, D, dh, S, nx, Q.data_ptr(), K.data_ptr(), V.data_ptr(), Kx.data_ptr(), Vx.data_ptr(),
                                     O.data_ptr(), scratch.data_ptr(), nb, -1, C.byref(ms), torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
tr = O.view(-1)[: 64 * 16 * 2].view(torch.int64).view(64, 16).cpu()
t0 = tr[8, 0].item()
names = ["sm:top", "sm:s_full", "sm:ldtm", "sm:max/alpha", "sm:exp+st", "sm:wait+arr", "sm:consume", "-",
         "mma:top", "mma:kv_full", "mma:w_p0", "mma:p0", "mma:w_p1", "mma:p1"] if terms >= 20 else \
        ["sm:top", "sm:s_full", "sm:ldtm", "sm:max/alpha", "sm:p_empty", "sm:exp+sts", "sm:fence+arr", "sm:consume",
         "mma:top", "mma:kv_full", "mma:S_issued", "mma:PV_issued"]
print("iter " + " ".join(f"{n:>13}" for n in names))
for i in range(8, 28):
    print(f"{i:4d} " + " ".join(f"{tr[i, k].item() - t0:13d}" for k in range(len(names))))
d = tr[20:50, 0] - tr[19:49, 0]
print("mean cycles per iteration (softmax thread):", d.float().mean().item())
