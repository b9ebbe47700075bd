// sm_100a primitives used by the tensor-core kernels: mbarrier, TMA (cp.async.bulk.tensor), TMEM
// allocation, tcgen05.mma / commit / ld, UMMA shared-memory + instruction descriptors, and the
// error-compensated fp32 -> split-bf16 conversion.  Inline PTX only (no CUTLASS); the bit layouts were
// cross-checked against cute/arch/mma_sm100_desc.hpp.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace a2p {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU box.  ~4e9 cycles (a couple of seconds) then trap.
// The slow path is ONE out-of-line copy: the kernels have dozens of wait sites and are instruction-fetch sensitive
// (every launch starts with a cold instruction cache; profiles/r01k).
__device__ __noinline__ void mbar_wait_slow(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("a2p: mbarrier wait timed out (block %d thread %d parity %u)\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(bar, parity);
}
// call-free variant for kernels that use setmaxnreg (a call into a function that needs more registers than a shrunk
// warp owns cannot be allocated): bounded spin, trap without a message
__device__ __forceinline__ void mbar_wait_nc(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// L2 eviction-priority policies for TMA loads: weights are re-read by every CTA of a launch and again on the next
// diffusion step (evict_last); K/V-cache and activation tiles stream through once per launch (evict_first).
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void tma_load_3d_hint(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}

// ------------------------------------------------------------------ TMEM
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "n"(NCOLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS));
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread i = TMEM lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ UMMA descriptors
// K-major operand tile in the canonical SWIZZLE_128B layout written by TMA (rows of 128 B, 8-row groups of
// 1024 B): start address (>>4), LBO = 1 (unused for swizzled K-major), SBO = 1024 B (>>4 = 64),
// version = 1 (Blackwell), layout_type = 2 (SWIZZLE_128B).  Tile base must be 1024-B aligned.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// The same descriptor split into 32-bit halves so that advancing along K (or to another tile) is ONE 32-bit add
// on the low word: hi = SBO | version | layout (constant), lo = (addr >> 4) | LBO.  smem < 256 KB, so the
// 14-bit address field never carries.
constexpr uint32_t kDescHiSw128 = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
__device__ __forceinline__ uint64_t desc_make(uint32_t lo) { return ((uint64_t)kDescHiSw128 << 32) | (uint64_t)lo; }

// same for 64-byte rows (SWIZZLE_64B: 8-row groups of 512 B), used when the K extent of a tile is 32 bf16
__device__ __forceinline__ uint64_t smem_desc_sw64(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}
// kind::f16 instruction descriptor: D = fp32, A = B = bf16, both K-major, shape M x N (K = 16 per instruction)
__host__ __device__ constexpr uint32_t idesc_bf16_f32(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ error-compensated split
// x = p0 + p1 (+ p2) with p_i bf16, round-to-nearest at every step: 16 (24) significant bits.
template <int P>
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16* out) {
  float r = x;
#pragma unroll
  for (int i = 0; i < P; ++i) {
    __nv_bfloat16 b = __float2bfloat16_rn(r);
    out[i] = b;
    r = r - __bfloat162float(b);
  }
}

// Pairwise split WITHOUT conversion instructions: on sm_100 both F2F and F2FP (cvt.rn.bf16x2.f32) issue on the
// XU pipe (~8 lanes/clk/SM, shared with MUFU.EX2), which is the limiter of the softmax warps.  Rounding is done
// with integer adds on the ALU pipe instead: u + 0x8000 rounds the upper 16 bits half-up (a tie happens with
// probability 2^-16 and costs half a bf16 ulp of the LAST plane), PRMT packs the two upper halves.
// out[t] holds plane t of (a, b) packed as bf16x2 (a in the low half = lower address).
template <int P>
__device__ __forceinline__ void split_bf16_pair(float a, float b, uint32_t* out) {
#pragma unroll
  for (int t = 0; t < P; ++t) {
    const uint32_t ua = __float_as_uint(a) + 0x8000u, ub = __float_as_uint(b) + 0x8000u;
    out[t] = __byte_perm(ua, ub, 0x7632);
    if (t + 1 < P) {
      a = a - __uint_as_float(ua & 0xFFFF0000u);
      b = b - __uint_as_float(ub & 0xFFFF0000u);
    }
  }
}
// Truncating variant for THREE planes: plane t = upper 16 bits of the running residual (PRMT picks them directly),
// residual = x - (x & 0xFFFF0000) is exact.  Three truncated planes keep 3 x 7 = 21+ explicit mantissa bits
// (|err| <= 2^-21 |x|, ~fp32 for softmax probabilities) for 2 ALU + 2 FMA-pipe ops per value instead of 7.
__device__ __forceinline__ void split_bf16_pair_trunc3(float a, float b, uint32_t* out) {
  const uint32_t a0 = __float_as_uint(a), b0 = __float_as_uint(b);
  out[0] = __byte_perm(a0, b0, 0x7632);
  a = a - __uint_as_float(a0 & 0xFFFF0000u);
  b = b - __uint_as_float(b0 & 0xFFFF0000u);
  const uint32_t a1 = __float_as_uint(a), b1 = __float_as_uint(b);
  out[1] = __byte_perm(a1, b1, 0x7632);
  a = a - __uint_as_float(a1 & 0xFFFF0000u);
  b = b - __uint_as_float(b1 & 0xFFFF0000u);
  out[2] = __byte_perm(__float_as_uint(a), __float_as_uint(b), 0x7632);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2(x) for x <= 0 on the FMA/ALU pipes (no MUFU): round-to-nearest range reduction x = n + f, |f| <= 0.5,
// degree-6 Taylor of 2^f (|rel err| < 1.3e-7), exponent patched in with an integer add.  Used for a fraction of the
// softmax elements because MUFU.EX2 (8 lanes/clk/SM on B200) is the bottleneck of the dh = 32 attention.
__device__ __forceinline__ float ex2_poly(float x) {
  const float t = fmaxf(x, -125.f);
  const float z = t + 12582912.f;                       // 1.5 * 2^23: the low mantissa bits of z hold round(t)
  const float f = t - (z - 12582912.f);
  float p = 1.5403530e-4f;
  p = fmaf(p, f, 1.3333558e-3f);
  p = fmaf(p, f, 9.6181291e-3f);
  p = fmaf(p, f, 5.5504109e-2f);
  p = fmaf(p, f, 2.4022651e-1f);
  p = fmaf(p, f, 6.9314718e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(z) << 23));
}

}  // namespace umma
}  // namespace a2p
