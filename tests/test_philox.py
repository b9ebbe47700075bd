"""In-kernel noise of the sampler epilogue: Philox4x32-10 known-answer vectors on the host build of the
kernel's block function (CPU), and on the GPU the N(0,1) stream drawn by K3: moments, determinism, shard invariance and
agreement with a numpy restatement of counter -> Box-Muller."""
import ctypes as C

import numpy as np
import pytest
import torch

KAT = [  # Random123 kat_vectors: philox4x32-10  (counter x4, key x2 -> output x4)
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def _philox_np(c, k):
    c = [np.uint64(v) for v in c]; k = [np.uint64(v) for v in k]
    M0, M1, W0, W1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0x9E3779B9), np.uint64(0xBB67AE85), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ k[0], p1 & MASK, (p0 >> np.uint64(32)) ^ c[3] ^ k[1], p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return [int(v) for v in c]


def test_philox_known_answers_host_build():
    from audio2photoreal_b200 import _lib
    lib = _lib.load_testing()
    lib.a2p_test_philox.argtypes = [C.c_uint32] * 6 + [C.POINTER(C.c_uint32)]
    lib.a2p_test_philox.restype = None
    out = (C.c_uint32 * 4)()
    for ctr, key, want in KAT:
        lib.a2p_test_philox(*ctr, *key, out)
        assert tuple(out) == want
        assert tuple(_philox_np(ctr, key)) == want


@pytest.mark.gpu
def test_k3_philox_noise_stream():
    from audio2photoreal_b200 import _lib
    lib = _lib.load()
    B, Cc, T, seed = 4, 104, 600, 0x1234567890ABCDEF
    dev = "cuda"
    z = torch.zeros(B, Cc, 1, T, device=dev)
    x0 = torch.zeros(B, T, Cc, device=dev)
    co = torch.zeros(8, device=dev); co[7] = 1.0          # ancestral: out = coef1*x0 + coef2*x_t + sigma*noise  ->  noise
    st = torch.cuda.current_stream().cuda_stream

    def draw(nrows, row0, it):
        out, pred = torch.empty(nrows, Cc, 1, T, device=dev), torch.empty(nrows, Cc, 1, T, device=dev)
        _lib.check(lib.a2p_sampler_step_rng(1, nrows, Cc, T, z[:nrows].data_ptr(), x0[:nrows].data_ptr(), None, None, co.data_ptr(),
                                            seed, it, row0, 0, out.data_ptr(), pred.data_ptr(), st))
        torch.cuda.synchronize()
        return out
    a, a2, b = draw(B, 0, 3), draw(B, 0, 3), draw(B, 0, 4)
    assert torch.equal(a, a2) and not torch.equal(a, b)                       # deterministic per (seed, iteration)
    assert torch.equal(draw(2, 2, 3), a[2:4])                                 # rows 2..3 of the batch, drawn by "another rank"
    n = a.numel()
    assert abs(a.mean().item()) < 4 / n ** 0.5 and abs(a.var().item() - 1) < 0.02 and abs((a ** 4).mean().item() - 3) < 0.15
    # numpy restatement of element (b=1, c=5, t=0..7)
    for t in range(8):
        e = (1 * Cc + 5) * T + t
        r = _philox_np((e & 0xFFFFFFFF, e >> 32, 3, 0), (seed & 0xFFFFFFFF, seed >> 32))
        u1, u2 = (r[0] + 1.0) * 2.0 ** -32, r[1] * 2.0 ** -32
        ref = np.sqrt(-2.0 * np.log(u1)) * np.cos(2 * np.pi * u2)
        assert abs(a[1, 5, 0, t].item() - ref) < 2e-5 * max(1.0, abs(ref))
