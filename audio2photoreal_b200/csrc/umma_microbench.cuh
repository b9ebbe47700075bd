// Measurement kernel: sustained cost (SM cycles) of one tcgen05.mma (M = 128, K = 16, bf16) as a function of N and of
// where the A operand lives (shared memory descriptor vs tensor memory).  Used to size the tiles of the GEMM and
// attention kernels (results in profiles/).
#pragma once
#include "umma.cuh"

namespace a2p {

template <int N, int A_TMEM>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n_mma, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { umma::mbar_init(&bar, 1); umma::fence_barrier_init(); }
  if (warp == 1) umma::tmem_alloc<512>(&slot);
  umma::fence_proxy_async();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tm = slot;
  if (warp == 0) {
    constexpr uint32_t idesc = umma::idesc_bf16_f32(128, N);
    const uint32_t loA = umma::desc_lo(umma::smem_u32(smem)), loB = umma::desc_lo(umma::smem_u32(smem + 16384));
    long long t0 = 0;
    if (umma::elect_one()) {
      t0 = clock64();
      for (int i = 0; i < n_mma; ++i) {
        const int k = i & 3;
        if (A_TMEM) {
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
              ::"r"(tm), "r"(tm + 256 + 8 * k), "l"(umma::desc_make(loB + 2 * k)), "r"(idesc), "r"(1u)
              : "memory");
        } else {
          umma::mma_bf16(tm, umma::desc_make(loA + 2 * k), umma::desc_make(loB + 2 * k), idesc, 1u);
        }
      }
      umma::mma_commit(&bar);
    }
    __syncwarp();
    umma::mbar_wait(&bar, 0);
    if (umma::elect_one()) out_cycles[0] = clock64() - t0;
    __syncwarp();
  }
  __syncthreads();
  if (warp == 1) { umma::fence_after(); umma::tmem_dealloc<512>(tm); }
}

template <int N, int A_TMEM>
inline int run_mma_rate(int n_mma, long long* out, cudaStream_t st) {
  A2P_CUDA(cudaFuncSetAttribute(mma_rate_kernel<N, A_TMEM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  mma_rate_kernel<N, A_TMEM><<<1, 128, 64 * 1024, st>>>(n_mma, out);
  A2P_CUDA(cudaGetLastError());
  return 0;
}

inline int launch_mma_rate(int N, int a_tmem, int n_mma, long long* out, cudaStream_t st) {
#define A2P_CASE(NN) if (N == NN) return a_tmem ? run_mma_rate<NN, 1>(n_mma, out, st) : run_mma_rate<NN, 0>(n_mma, out, st);
  A2P_CASE(32) A2P_CASE(64) A2P_CASE(128) A2P_CASE(256)
#undef A2P_CASE
  A2P_FAIL("mma_rate: N must be 32, 64, 128 or 256");
}

}  // namespace a2p
