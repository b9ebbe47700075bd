// Shared helpers for the a2p_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
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

__device__ __forceinline__ float mishf(float x) {
  // x * tanh(softplus(x)); softplus threshold 20 like torch
  float sp = (x > 20.f) ? x : log1pf(expf(x));
  return x * tanhf(sp);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

}  // namespace a2p
