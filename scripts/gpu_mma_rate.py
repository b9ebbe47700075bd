import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio2photoreal_b200 import _lib
lib = _lib.load_testing()
lib.a2p_test_mma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(1, dtype=torch.int64, device="cuda")
for a_tmem in (0, 1):
    for N in (32, 64, 128, 256):
        for n in (64, 1024):
            _lib.check_testing(lib.a2p_test_mma_rate(N, a_tmem, n, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            print(f"A_from_{'TMEM' if a_tmem else 'SMEM'} N={N:3d} n_mma={n:5d}: {out.item()/n:7.1f} cycles/MMA (total {out.item()})", flush=True)
print("DONE")
