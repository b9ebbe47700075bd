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

// tcgen05.ld / tcgen05.st rate of the WARPS: every participating warp moves `n_ops` times 32 lanes x 32 columns x 4 B = 4 KB between its
// TMEM quarter and registers (32x32b.x32).  warps_per_quarter = 1 or 2 (the attention kernel has two softmax warps per quarter).
template <int ST>
__global__ void __launch_bounds__(384, 1) tmem_ldst_rate_kernel(int n_ops, int n_warps, long long* out_cycles) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) umma::tmem_alloc<512>(&slot);
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (warp >= 4 && warp < 4 + n_warps) {
    const uint32_t addr = tm + ((uint32_t)((warp & 3) * 32) << 16) + ((warp - 4) >> 2) * 128;
    float v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = (float)c;
    asm volatile("bar.sync 1, %0;" ::"r"(n_warps * 32) : "memory");
    t0 = clock64();
    for (int i = 0; i < n_ops; ++i) {
      if (ST) {
        uint32_t* r = reinterpret_cast<uint32_t*>(v);
        asm volatile(
            "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
            "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
            "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
            ::"r"(addr + (i & 3) * 32), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
              "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
              "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
              "r"(r[30]), "r"(r[31])
            : "memory");
      } else {
        umma::tmem_ld32(addr + (i & 3) * 32, v);
      }
    }
    if (ST) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    else umma::tmem_ld_wait();
    t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) acc += v[c];
    if (acc == 12345.678f) out_cycles[1] = 1;     // keep the loads alive
    if (threadIdx.x == 128) out_cycles[0] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { umma::fence_after(); umma::tmem_dealloc<512>(tm); }
}

inline int launch_tmem_ldst_rate(int store, int n_ops, int n_warps, long long* out, cudaStream_t st) {
  if (n_warps < 1 || n_warps > 8) A2P_FAIL("tmem_ldst_rate: 1..8 warps");
  if (store) tmem_ldst_rate_kernel<1><<<1, 384, 0, st>>>(n_ops, n_warps, out);
  else tmem_ldst_rate_kernel<0><<<1, 384, 0, st>>>(n_ops, n_warps, out);
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
