import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audio2photoreal_b200 import _lib
lib = _lib.load_testing()
lib.a2p_test_mma_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
out = torch.zeros(2, dtype=torch.int64, device="cuda")
if len(sys.argv) > 1 and sys.argv[1] == "tmem":
    lib.a2p_test_tmem_ldst_rate.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    print("# warp-side TMEM traffic: n_warps warps x n_ops x (32 lanes x 32 columns x 4 B = 4 KB) tcgen05.ld / tcgen05.st 32x32b.x32")
    for store in (0, 1):
        for nw in (1, 2, 4, 8):
            for n in (64, 512):
                _lib.check_testing(lib.a2p_test_tmem_ldst_rate(store, n, nw, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
                torch.cuda.synchronize()
                cyc = out[0].item()
                print(f"{'st' if store else 'ld'} warps={nw} n_ops={n:4d}: {cyc:8d} cycles  {cyc / n:7.1f} cycles per instruction and warp  "
                      f"{nw * n * 4096 / cyc:7.1f} B/clk per SM", flush=True)
    print("DONE"); sys.exit(0)
for a_tmem in (0, 1):
    for N in (32, 64, 128, 256):
        for n in (64, 1024):
            _lib.check_testing(lib.a2p_test_mma_rate(N, a_tmem, n, out.data_ptr(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            print(f"A_from_{'TMEM' if a_tmem else 'SMEM'} N={N:3d} n_mma={n:5d}: {out.item()/n:7.1f} cycles/MMA (total {out.item()})", flush=True)
print("DONE")
