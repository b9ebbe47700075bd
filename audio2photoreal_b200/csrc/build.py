"""Build liba2p_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "liba2p_b200.so")                    # the product: C-ABI of include/a2p_b200.h
TEST_LIB = os.path.join(PKG, "liba2p_b200_testing.so")       # test / measurement hooks (include/a2p_b200_testing.h)
TARGETS = {LIB: ["engine.cu"], TEST_LIB: ["testing.cu"]}
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


STAMP = LIB + ".srchash"      # sha256 of the sources the library was built from (mtimes do not survive a snapshot copy)


def source_hash() -> str:
    import hashlib
    deps = sorted(os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cu", ".cuh", ".h")))
    inc = os.path.join(os.path.dirname(PKG), "include")
    deps += sorted(os.path.join(inc, f) for f in os.listdir(inc) if f.endswith(".h"))
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _stale() -> bool:
    if not (os.path.exists(LIB) and os.path.exists(TEST_LIB) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile both libraries (concurrently) unless the stamp matches the sources; returns the product library's path."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    procs = []
    for out, srcs in TARGETS.items():
        cmd = [nvcc] + NVCC_FLAGS + [os.path.join(HERE, s) for s in srcs] + ["-o", out]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((out, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)))
    for out, pr in procs:
        so, se = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed for {out}:\n" + so + se)
        if verbose:
            print(se)
    with open(STAMP, "w") as f:
        f.write(source_hash())
    return LIB


def build_trace() -> str:
    """liba2p_b200_trace.so: the testing library with the clock64 timeline of umma_attn2_kernel compiled in (diagnostics only:
    scripts/gpu_attn_trace.py); not part of build() and never loaded by the product or the tests."""
    out = os.path.join(PKG, "liba2p_b200_trace.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-DA2P_ATTN2_TRACE=1", os.path.join(HERE, "testing.cu"), "-o", out]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out


if __name__ == "__main__":
    if "--trace" in sys.argv:
        print(build_trace())
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
