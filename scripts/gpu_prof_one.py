"""Tiny driver for ncu captures: runs ONE configuration of a tcgen05 kernel a few times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
which = sys.argv[1]
if which == "gemm":
    from scripts.gpu_tc_gemm import run
    run(int(sys.argv[2]), 9600, 768, 256, iters=2)
elif which == "attn":
    from scripts.gpu_tc_attn import run
    run(int(sys.argv[2]), 16, 600, 256, 32, 1998, 2, iters=2)
