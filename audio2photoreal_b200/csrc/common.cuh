// Shared helpers for the a2p_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string>

namespace a2p {

// thread-local last-error string surfaced through a2p_last_error()
inline std::string& last_error() {
  static thread_local std::string s;
  return s;
}

#define A2P_FAIL(...)                                   \
  do {                                                  \
    char _buf[512];                                     \
    snprintf(_buf, sizeof(_buf), __VA_ARGS__);          \
    a2p::last_error() = _buf;                           \
    return 1;                                           \
  } while (0)

#define A2P_CUDA(expr)                                                                       \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) A2P_FAIL("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#define A2P_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r) return _r;         \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Addressing of a per-sample operand that exists once per CFG branch:
//   rows [0, rows_per_branch) -> base[0] + r * stride[0]      (cond; stride 0 = shared)
//   rows [rows_per_branch, ..) -> base[1] + (r - rpb) * stride[1]  (uncond)
struct BranchPtr {
  const float* base[2];
  long long stride[2];
  int rows_per_branch;
  __device__ __forceinline__ const float* at(int r) const {
    int br = (r >= rows_per_branch) ? 1 : 0;
    int rr = r - br * rows_per_branch;
    return base[br] + (long long)rr * stride[br];
  }
};

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl() may start while its stream
// predecessor is still draining; it must call pdl_wait() before touching anything the predecessor wrote.  Only kernels
// containing pdl_wait() may be launched with the attribute.  pdl_trigger() allows the stream successor to be scheduled once
// every CTA of the grid has issued it (or exited).  Round 1 triggered at kernel entry: the successor's CTAs then sat on SMs
// (227 KB of shared memory each) spinning in pdl_wait() for the whole duration of this kernel -- neutral on a single
// forward, 10 % slower with concurrent forwards (profiles/r01t).  The trigger now sits where a CTA's producer warp has
// requested its last inbound tile, i.e. a few microseconds before the CTA ends: only the successor's prologue (barrier
// init, tensor-memory allocation, descriptor prefetch) and the launch latency overlap this kernel's tail: B = 8 loop +4.6 %, B = 4 +3 %, B = 32 unchanged (SM-time bound) -- default on
// (profiles/r02_chain_nsplit_pdl.txt).  A2P_PDL=0 switches the attribute off.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifndef A2P_PDL_DEFAULT
#define A2P_PDL_DEFAULT 1
#endif
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("A2P_PDL"); v = e ? (atoi(e) != 0) : A2P_PDL_DEFAULT; }
  return v == 1;
}

// Scheduling priority of the NEXT launch_pdl() launches (0 = the stream's own).  When two forwards run as concurrent
// streams, pending CTAs of a higher-priority kernel are placed before pending CTAs of a lower-priority one.
inline int& launch_priority() { static int v = 0; return v; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (launch_priority() != 0) {
    attr[n].id = cudaLaunchAttributePriority;
    attr[n].val.priority = launch_priority();
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float mishf(float x) {
  // x * tanh(softplus(x)); softplus threshold 20 like torch
  float sp = (x > 20.f) ? x : log1pf(expf(x));
  return x * tanhf(sp);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

}  // namespace a2p
