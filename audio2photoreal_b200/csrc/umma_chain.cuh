// Fused "row chain" kernel for the D = 256 (pose) denoiser on sm_100a, split-bf16 x2 operands (3 tensor-core
// products per MAC, fp32 accumulation in TMEM).  One CTA owns 128 consecutive rows of the residual stream and runs,
// without leaving the SM:
//
//   GEMM0    acc0[128,256]  = A0[128,K0] * W0[256,K0]^T          A0 / W0 tiles streamed from global by TMA
//   E_A      x = x + (film_scale + 1) * (acc0 + b0) + film_shift (or x = acc0 + b0); the old x tile arrives by TMA,
//            the new one leaves by TMA store; h = LayerNorm(x) (two-pass, fp32), optional full-width RoPE (table rows
//            arrive by TMA), split into two bf16 planes written IN PLACE over x in tensor memory
//   GEMM1    acc1[128,N1]   = h[128,256] * W1[N1,256]^T          A operand from TENSOR MEMORY, 128 columns at a time
//   E_B      bias, optional scale / erf-GELU, split into planes -> global (Q|K planes, FFN hidden planes, ...)
//   V job    V[128,256]     = h'[128,256] * W2[256,256]^T        h' = un-rotated LayerNorm output (x tile re-read by
//            TMA); the epilogue stores V TRANSPOSED (the V^T planes of the PV product of the attention kernel)
//
// This replaces, per decoder layer, 4 LayerNorm(+RoPE) launches and 9 GEMM launches of the unfused arm by 4 launches
// (transformer_modules.py:190-217: out_proj+FiLM+residual -> norm -> rotate -> in_proj of the NEXT block).
//
// Everything that comes from global memory arrives through ONE ring of eleven 16 KB shared-memory slots filled by the
// TMA warp in a fixed, data-independent order (A0 / W0 tiles, x chunks, RoPE-table chunks, W1 tiles, x chunks again,
// W2 tiles); consumers (the MMA warp or an epilogue warpgroup) find their slot by position in that order.  No
// epilogue instruction waits on a global load: v0 of this kernel did (x, FiLM, LayerNorm weights, RoPE table with
// plain loads and no L1 -- the whole 227 KB is shared memory) and spent 25 of its 44 us there (profiles/r01i).
//
// Roles (384 threads): warp 0 TMA producer | warp 1 MMA issuer | warp 2 TMEM allocator | warps 4-11 two epilogue
// warpgroups (thread = TMEM lane = row).  E_A: the warpgroups split the 256 columns (row statistics are exchanged
// through shared memory).  E_B: they alternate 128-column accumulator halves, so the epilogue of half i overlaps
// the MMAs of half i+1.  TMEM: columns [0,256) acc0 -> x -> planes (chunk c of 32 columns holds plane 0 in its first
// 16 columns, plane 1 in the last 16: 8 columns per 16-element k-step), [256,384) / [384,512) the GEMM1 accumulators.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"

namespace a2p {

struct ChainParams {
  int M, T;                  // rows (samples * T); RoPE position = row % T, FiLM sample = row / T.  T >= 128.
  int K0;                    // GEMM0 reduction length (any multiple of 8; TMA zero-fills the tail of the last 64-chunk)
  const float* bias0;        // [256]
  int film_mode;             // 1: x += (scale + 1) * (acc0 + b0) + shift    0: x = acc0 + b0
  const float* film; long long film_ld; int film_scale_off, film_shift_off;
  int ln_mode;               // 1: h = LayerNorm(x) * ln_w + ln_b            0: h = x
  const float* ln_w; const float* ln_b;
  int rope;                  // rotate h before GEMM1 (the V job always uses the un-rotated h)
  int N1;                    // GEMM1 output columns (<= 1024)
  const float* bias1; float out_scale; int scale_ncols;   // out_scale applies to columns < scale_ncols (0 = all)
  int gelu;
  __nv_bfloat16* Cp; long long cp_plane_stride, ldcp; int remap_rps, remap_pad;
  int vjob; const float* bias2; __nv_bfloat16* Vt; long long vt_plane_stride, ldvt;
  int nsplit;                // > 1: every 128-row tile is worked on by nsplit CTAs (CTA pairs in pair mode), see "N split" below
  int eb_tma;                // E_B writes the output planes with TMA stores from a bf16 staging tile (set by the launcher)
  long long* trace;          // optional [64] clock64 timeline of CTA 0 (A2P_CHAIN_TRACE=1 in the test hook)
};

constexpr int CH_NWG = 4;                           // epilogue warpgroups (thread = TMEM lane = row; a warpgroup owns 64 of the 256 columns in E_A)
constexpr int CH_THREADS = 128 + 128 * CH_NWG;
constexpr int CH_NS = 11;                          // ring slots
constexpr int CH_TILE = 16384;                     // one slot: [128 rows][128 B], SWIZZLE_128B
constexpr int CH_STG_BYTES = 4 * CH_NWG * 2048;    // per epilogue warp: [32 rows][16 cols] fp32, 16-byte unit q of row r stored at q ^ ((r >> 1) & 3)
// parameter block (floats): bias0[256] | film[2 samples][scale 256 | shift 256] | ln_w[256] | ln_b[256] | bias1[1024] | bias2[256]
constexpr int CH_PB_BIAS0 = 0, CH_PB_FILM = 256, CH_PB_LNW = 1280, CH_PB_LNB = 1536, CH_PB_BIAS1 = 1792, CH_PB_BIAS2 = 2816;
constexpr int CH_PB_FLOATS = 3072;
constexpr int CH_RED_BYTES = CH_NWG * 512;          // [warpgroups][128 rows] fp32
constexpr int CH_SMEM_BYTES = CH_NS * CH_TILE + CH_STG_BYTES + CH_PB_FLOATS * 4 + CH_RED_BYTES + 512 + 1024;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16u(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T  (A: lane = row, 8 columns per 16 bf16 of K)
__device__ __forceinline__ void mma_bf16_tmem_a(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, uint32_t dst_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(m)), "r"(umma::smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void sts128u(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float4 lds128(uint32_t saddr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// exact-erf GELU (torch's default, transformer_modules.py:265) with the Abramowitz-Stegun 7.1.26 erf: max abs error
// 4.7e-7 over [-12, 12] (torch's own fp32 GELU is 1.2e-6 from fp64), 2 MUFU + ~13 FMA-pipe ops instead of erff().
__device__ __forceinline__ float gelu_as(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  float pl = 1.061405429f;
  pl = fmaf(pl, t, -1.453152027f);
  pl = fmaf(pl, t, 1.421413741f);
  pl = fmaf(pl, t, -0.284496736f);
  pl = fmaf(pl, t, 0.254829592f);
  pl *= t;
  const float e = umma::ex2_approx(-(z * z) * 1.4426950408889634f);
  const float er = copysignf(fmaf(-pl, e, 1.0f), x);
  const float hx = 0.5f * x;
  return fmaf(hx, er, hx);
}

// ---- packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 on sm_100): one issue slot for two elements.  The epilogues run at
// ~0.25 instructions per cycle per warp with two warps per scheduler, so their time is proportional to the instruction count.
struct f2 { float x, y; };
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 r;
  asm("{\n\t.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 r;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ f2 bc2(float v) { return f2{v, v}; }

// gelu_as on a pair (same formula, same constants: results identical to the scalar version up to fma contraction)
__device__ __forceinline__ f2 gelu_as2(f2 x) {
  const f2 z = mul2(f2{fabsf(x.x), fabsf(x.y)}, bc2(0.70710678118654752440f));
  const f2 den = fma2(bc2(0.3275911f), z, bc2(1.0f));
  f2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  f2 pl = fma2(bc2(1.061405429f), t, bc2(-1.453152027f));
  pl = fma2(pl, t, bc2(1.421413741f));
  pl = fma2(pl, t, bc2(-0.284496736f));
  pl = fma2(pl, t, bc2(0.254829592f));
  pl = mul2(pl, t);
  const f2 ze = mul2(mul2(z, z), bc2(-1.4426950408889634f));
  const f2 e = f2{umma::ex2_approx(ze.x), umma::ex2_approx(ze.y)};
  const f2 er0 = fma2(f2{-pl.x, -pl.y}, e, bc2(1.0f));
  const f2 er = f2{copysignf(er0.x, x.x), copysignf(er0.y, x.y)};
  const f2 hx = mul2(x, bc2(0.5f));
  return fma2(hx, er, hx);
}

// activations: plane 0 rounded to bf16, plane 1 = exact residual truncated (|x - p0 - p1| <= 2^-17 |x|, unbiased because the
// residual has either sign): 8 ALU/FMA-pipe ops per pair
__device__ __forceinline__ void split_act_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ua = __float_as_uint(a) + 0x8000u, ub = __float_as_uint(b) + 0x8000u;
  hi = __byte_perm(ua, ub, 0x7632);
  // residual against the NEGATED plane-0 values (sign bit flipped in the same logic op) as one packed add
  const f2 r = add2(f2{a, b}, f2{__uint_as_float((ua & 0xFFFF0000u) ^ 0x80000000u), __uint_as_float((ub & 0xFFFF0000u) ^ 0x80000000u)});
  lo = __byte_perm(__float_as_uint(r.x), __float_as_uint(r.y), 0x7632);
}

// ---- CTA-pair mode (CL = 2): the two CTAs of a cluster (one TPC) own neighbouring 128-row tiles and run every GEMM as
// `tcgen05.mma.cta_group::2` instructions of shape M = 256 issued by the leader (cluster rank 0): each CTA supplies ITS rows of
// A (shared memory or tensor memory) and HALF of the weight tile (N / 2 rows of W), and gets its 128 rows of D in its own
// tensor memory.  A CTA therefore streams half of every weight matrix: the per-SM ingress that bounds GEMM0 / GEMM1
// (profiles/r01w_source_chain_*: 46 % of the warp samples in operand waits; at B = 32 all 148 SMs pull W from L2 at once)
// drops from A + W to A + W / 2.  Protocol:
//   * both producers walk the same slot sequence; a tile consumed by the MMA warp signals the LEADER's full barrier
//     (`cp.async.bulk.tensor...cta_group::2` with the barrier address mapped to rank 0), which expects the bytes of both CTAs;
//   * the leader's commits are multicast to the same barrier of both CTAs (slot releases, accumulator-full signals);
//   * barriers the MMA warp waits on besides the ring (planes ready, accumulator drained) collect the arrivals of both CTAs'
//     epilogue warps (the peer arrives remotely).
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_smem_addr), "r"(rank));
  return ra;
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t rank) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_rank(umma::smem_u32(bar), rank)) : "memory");
}
// arrive on a barrier the MMA warp (leader CTA) waits on: local for the leader, remote for the peer
template <int CL>
__device__ __forceinline__ void arrive_at_leader(uint64_t* bar, uint32_t crank) {
  if (CL == 2 && crank != 0) mbar_arrive_remote(bar, 0);
  else umma::mbar_arrive(bar);
}
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {   // arrive on the same barrier of BOTH CTAs when the MMAs complete
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(umma::smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
template <int CL>
__device__ __forceinline__ void chain_commit(uint64_t* bar) {
  if (CL == 2) mma_commit_pair(bar);
  else umma::mma_commit(bar);
}
// D[tmem] (+)= A * B^T over the CTA pair (M = 256: 128 rows per CTA; every CTA holds N / 2 rows of B)
__device__ __forceinline__ void mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_bf16_tmem_a_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int CL>
__device__ __forceinline__ void chain_mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CL == 2) mma_bf16_pair(d, a, b, idesc, acc);
  else umma::mma_bf16(d, a, b, idesc, acc);
}
template <int CL>
__device__ __forceinline__ void chain_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (CL == 2) mma_bf16_tmem_a_pair(d, a_tmem, b, idesc, acc);
  else mma_bf16_tmem_a(d, a_tmem, b, idesc, acc);
}
// TMA tile load of a CTA pair: data into THIS CTA's shared memory, bytes counted on the barrier `bar_cluster_addr`
// (a shared::cluster address: the leader's full barrier)
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* m, uint32_t bar_cluster_addr, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(umma::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot_in_smem) {  // whole warp, the same warp of BOTH CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(umma::smem_u32(slot_in_smem)), "n"(NCOLS));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS));
}

// Epilogue-thread context of chain_emit_planes (kept out of line: it runs twice in kernels with a V job, and the chain
// kernel is instruction-fetch sensitive -- every launch starts with a cold instruction cache).
struct ChainEpiCtx {
  uint32_t tmem_row;       // tmem_base + lane quarter
  uint32_t slots_u32, row_off, pb_u32;
  uint64_t* s_full; uint64_t* s_empty;
  int wg, trow, ln_mode;
  float mean, rstd;
  uint32_t peer;
};

// planes of (rotated) LayerNorm(x), written in place over x in tensor memory.  seq_x >= 0: the x chunks come from that
// ring position (V job: tile re-read by TMA); otherwise from tensor memory.  rot: RoPE with table chunks at seq_tab.
template <int CL>
__device__ __forceinline__ void chain_emit_planes(const ChainEpiCtx c, int rot, int seq_x, int seq_tab) {
  const int rx = c.trow & 7;
#pragma unroll 1
  for (int cc = 0; cc < 8 / CH_NWG; ++cc) {
    const int ch = c.wg * (8 / CH_NWG) + cc;
    float v[32];
    if (seq_x >= 0) {
      const int qx = seq_x + cc * CH_NWG + c.wg;
      const int sx = qx % CH_NS;
      umma::mbar_wait(&c.s_full[sx], (qx / CH_NS) & 1);
      const uint32_t srow = c.slots_u32 + sx * CH_TILE + c.row_off;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float4 t = lds128(srow + ((u ^ rx) << 4));
        v[4 * u] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
      }
    } else {
      umma::tmem_ld32(c.tmem_row + ch * 32, v);
      umma::tmem_ld_wait();
    }
    // the shared-memory loads of a group are issued before its first dependent instruction (volatile asm keeps program order)
    if (c.ln_mode) {
      const uint32_t pw_ = c.pb_u32 + (CH_PB_LNW + ch * 32) * 4, pb_ = c.pb_u32 + (CH_PB_LNB + ch * 32) * 4;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float4 ww[4], bb[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { ww[k] = lds128(pw_ + (hf * 4 + k) * 16); bb[k] = lds128(pb_ + (hf * 4 + k) * 16); }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int u = hf * 4 + k;
          const f2 h01 = fma2(mul2(add2(f2{v[4 * u], v[4 * u + 1]}, bc2(-c.mean)), bc2(c.rstd)), f2{ww[k].x, ww[k].y}, f2{bb[k].x, bb[k].y});
          const f2 h23 = fma2(mul2(add2(f2{v[4 * u + 2], v[4 * u + 3]}, bc2(-c.mean)), bc2(c.rstd)), f2{ww[k].z, ww[k].w}, f2{bb[k].z, bb[k].w});
          v[4 * u] = h01.x; v[4 * u + 1] = h01.y; v[4 * u + 2] = h23.x; v[4 * u + 3] = h23.y;
        }
      }
    }
    if (rot) {
      const int qt = seq_tab + cc * CH_NWG + c.wg;
      const int st = qt % CH_NS;
      umma::mbar_wait(&c.s_full[st], (qt / CH_NS) & 1);
      const uint32_t trw = c.slots_u32 + st * CH_TILE + c.row_off;
      float4 cs[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) cs[u] = lds128(trw + ((u ^ rx) << 4));   // (cos0, sin0, cos1, sin1)
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float h0 = v[4 * u], h1 = v[4 * u + 1], h2 = v[4 * u + 2], h3 = v[4 * u + 3];
        const f2 r01 = fma2(f2{h0, h1}, bc2(cs[u].x), mul2(f2{-h1, h0}, bc2(cs[u].y)));
        const f2 r23 = fma2(f2{h2, h3}, bc2(cs[u].z), mul2(f2{-h3, h2}, bc2(cs[u].w)));
        v[4 * u] = r01.x; v[4 * u + 1] = r01.y; v[4 * u + 2] = r23.x; v[4 * u + 3] = r23.y;
      }
    }
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {     // 8 pairs at a time: 8 + 8 packed registers live
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) split_act_pair(v[hf * 16 + 2 * e], v[hf * 16 + 2 * e + 1], hi[e], lo[e]);
      tmem_st8u(c.tmem_row + ch * 32 + hf * 8, hi);
      tmem_st8u(c.tmem_row + ch * 32 + 16 + hf * 8, lo);
    }
  }
  if (seq_x >= 0 || rot) {          // release the slots this warpgroup has finished reading (one barrier for its chunks)
    asm volatile("bar.sync %0, 128;" ::"r"(2 + c.wg) : "memory");
    if (c.trow == 0) {
#pragma unroll 1
      for (int cc = 0; cc < 8 / CH_NWG; ++cc) {
        if (seq_x >= 0) umma::mbar_arrive(&c.s_empty[(seq_x + cc * CH_NWG + c.wg) % CH_NS]);
        if (rot) umma::mbar_arrive(&c.s_empty[(seq_tab + cc * CH_NWG + c.wg) % CH_NS]);
      }
    }
  }
  tmem_st_wait();
  umma::fence_before();
}

#define CH_TRACE(slot, cond) do { if (p.trace && blockIdx.x == 0 && (cond)) p.trace[slot] = clock64(); } while (0)

template <int CL>
__global__ void __launch_bounds__(CH_THREADS, 1)
umma_chain_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmW0,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                  const __grid_constant__ CUtensorMap tmXin, const __grid_constant__ CUtensorMap tmXout,
                  const __grid_constant__ CUtensorMap tmTab, const __grid_constant__ CUtensorMap tmC,
                  const __grid_constant__ CUtensorMap tmVt, ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sSlots = smem;
  float* sStg = reinterpret_cast<float*>(sSlots + CH_NS * CH_TILE);
  float* sPB = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sStg) + CH_STG_BYTES);
  float* sRed = sPB + CH_PB_FLOATS;
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sRed) + CH_RED_BYTES);
  uint64_t* s_full = bars;             // [11]
  uint64_t* s_empty = bars + 11;       // [11]
  uint64_t* acc0_full = bars + 22;
  uint64_t* a_ready = bars + 23;       // 256 arrivals
  uint64_t* acc1_full = bars + 24;     // [2]
  uint64_t* acc1_empty = bars + 26;    // [2] 128 arrivals
  uint64_t* a_reads_done = bars + 28;
  uint64_t* a2_ready = bars + 29;      // 256 arrivals
  uint64_t* x_stored = bars + 30;      // the x tile written by E_A is globally visible (store warp)
  uint64_t* x_written = bars + 31;     // [8] 128 arrivals each: x chunk at ring position seqEA + j is in its staging slot
  uint64_t* m_full = bars + 39;        // [11] pair mode, leader only: "the tiles of BOTH CTAs for this MMA use of slot s have landed"
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 39 + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = CL == 2 ? cluster_ctarank() : 0u, peer = crank ^ 1u;
  const int kc0 = ceil_div(p.K0, 64);
  const int NH1 = ceil_div(p.N1, 128);
  const int n_acc = NH1 + (p.vjob ? 2 : 0);
  // ---- N split (small batches: a launch of 10-40 tiles leaves most of the 148 SMs idle, and its duration is the serial
  // GEMM0 -> E_A -> GEMM1 / E_B time of ONE tile).  Every tile is given to nsplit "parts" = consecutive CTAs (consecutive CTA pairs
  // in pair mode); each part runs GEMM0 + E_A redundantly (idle SMs are free; bit-identical work) and then only ITS range
  // [h_lo, h_hi) of the n_acc 128-column accumulator halves of GEMM1 / the V job.  The parts never talk to each other: the
  // residual stream is read from x_in and written to a DIFFERENT buffer x_out (by the parts that own V halves, which re-read
  // their own store; by part 0 when there is no V job), so no part can observe another part's update of the tile.
  const int nsp = p.nsplit > 1 ? p.nsplit : 1;
  const int unit_ = CL == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int tile_ = unit_ / nsp, part = unit_ - tile_ * nsp;
  const int m0 = (CL == 2 ? tile_ * 2 + (int)crank : tile_) * 128;
  const int h_lo = part * n_acc / nsp, h_hi = (part + 1) * n_acc / nsp, n_loc = h_hi - h_lo;
  const int g_lo = ::min(h_lo, NH1), g_hi = ::min(h_hi, NH1), n_g1 = g_hi - g_lo;       // GEMM1 halves of this part
  const int v_lo = ::max(h_lo, NH1) - NH1, v_hi = ::max(h_hi, NH1) - NH1, n_v = v_hi - v_lo;   // V-job halves of this part
  const bool x_writer = p.vjob ? n_v > 0 : part == 0;
  // ring positions of the slot sequence (identical arithmetic in every role)
  // single CTA: per 64-wide K chunk of GEMM0 two A0 tiles + four W0 tiles (plane x 128-row half), per 128-column half of
  // GEMM1 / the V job eight W tiles (K chunk x plane).  CTA pair: every CTA fetches HALF of the weight rows -- two A0 tiles +
  // two W0 tiles (plane; its 128 of the 256 rows) per K chunk, four W tiles per accumulator half (K chunk; both planes of its 64 rows)
  constexpr int G0S = CL == 2 ? 4 : 6, G1S = CL == 2 ? 4 : 8;
  const int seqEA = G0S * kc0;                    // 8 x chunks (loaded, or only reserved as store buffers if !film_mode)
  const int seqTab = seqEA + 8;                   // 8 RoPE-table chunks (if rope)
  const int seqG1 = seqTab + (p.rope ? 8 : 0);    // n_g1 * G1S W1 tiles
  const int seqVx = seqG1 + G1S * n_g1;           // 8 x chunks again (V job)
  const int seqV = seqVx + 8;                     // n_v * G1S W2 tiles

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmA0); umma::prefetch_tmap(&tmW0); umma::prefetch_tmap(&tmW1); umma::prefetch_tmap(&tmXin);
    umma::prefetch_tmap(&tmXout);
    if (p.eb_tma) { umma::prefetch_tmap(&tmC); if (p.vjob) umma::prefetch_tmap(&tmVt); }
    if (p.rope) umma::prefetch_tmap(&tmTab);
    if (p.vjob) umma::prefetch_tmap(&tmW2);
  }
  if (warp == 1 && lane == 0) {
    // a_ready / a2_ready / acc1_empty are waited on by the MMA warp (leader CTA in pair mode): one arrival per epilogue WARP
    // (4 * CH_NWG warps write the planes, 2 * CH_NWG drain an accumulator half) of each of the CL CTAs
    for (int i = 0; i < CH_NS; ++i) { umma::mbar_init(&s_full[i], 1); umma::mbar_init(&s_empty[i], 1); umma::mbar_init(&m_full[i], 1); }
    umma::mbar_init(acc0_full, 1); umma::mbar_init(a_ready, 4 * CH_NWG * CL);
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&acc1_full[i], 1); umma::mbar_init(&acc1_empty[i], 2 * CH_NWG * CL); }
    umma::mbar_init(a_reads_done, 1); umma::mbar_init(a2_ready, 4 * CH_NWG * CL); umma::mbar_init(x_stored, 1);
    for (int i = 0; i < 8; ++i) umma::mbar_init(&x_written[i], 128);
    umma::fence_barrier_init();
  }
  if (warp == 2) { if (CL == 2) tmem_alloc_pair<512>(tmem_slot); else umma::tmem_alloc<512>(tmem_slot); }
  umma::fence_before();
  __syncthreads();
  if (CL == 2) cluster_sync_all();     // the peer's barriers and tensor memory exist before anything is committed / arrived remotely
  umma::fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t slots_u32 = umma::smem_u32(sSlots);
  pdl_wait();
  CH_TRACE(0, threadIdx.x == 128);

  if (warp == 0) {
    // ================= TMA producer: one pass over the slot sequence =================
    // Rolled nested loops (the fully unrolled form of this role alone was 25 KB of cold code), but no per-load decode
    // arithmetic: the load -> consume -> release -> reload round trip of a slot is the critical path of the GEMMs.
    int sl_ = 0; uint32_t ph_ = 0;
    // kind 0: bf16 tile read by the MMA warp (3-D map), 1: fp32 tile read by this CTA's epilogue warps (2-D map), 2: reserve only
    auto issue = [&](const CUtensorMap* tm, int kind, int c0, int c1, int c2) {
      umma::mbar_wait(&s_empty[sl_], ph_ ^ 1);
      if (umma::elect_one()) {
        if (kind == 2) {
          umma::mbar_arrive(&s_full[sl_]);
        } else if (kind == 1) {
          umma::mbar_expect_tx(&s_full[sl_], CH_TILE);
          tma_load_2d(tm, &s_full[sl_], slots_u32 + sl_ * CH_TILE, c0, c1);
        } else if (CL == 2) {
          // MMA tiles of a pair are counted on the LEADER's m_full[slot] (the bytes of both CTAs), a barrier that only MMA
          // uses of the slot touch: the two CTAs are not in lockstep on their epilogue-consumed uses of a slot (x / table
          // chunks), so a peer's tile for the NEXT use must never land on a barrier whose current phase belongs to the
          // leader's own epilogue tile.  s_full[slot] only keeps its lap parity in step (nobody waits on it for MMA uses).
          umma::mbar_arrive(&s_full[sl_]);
          if (crank == 0) umma::mbar_expect_tx(&m_full[sl_], 2 * CH_TILE);
          tma_load_3d_pair(tm, mapa_rank(umma::smem_u32(&m_full[sl_]), 0), sSlots + sl_ * CH_TILE, c0, c1, c2);
        } else {
          umma::mbar_expect_tx(&s_full[sl_], CH_TILE);
          umma::tma_load_3d(tm, &s_full[sl_], sSlots + sl_ * CH_TILE, c0, c1, c2);
        }
      }
      __syncwarp();
      if (++sl_ == CH_NS) { sl_ = 0; ph_ ^= 1; }
    };
#pragma unroll 1
    for (int kc = 0; kc < kc0; ++kc) {
#pragma unroll 1
      for (int j = 0; j < G0S; ++j) {
        if (j < 2) issue(&tmA0, 0, kc * 64, m0, j);
        else if (CL == 2) issue(&tmW0, 0, kc * 64, (int)crank * 128, j - 2);            // my 128 of the 256 rows, plane j - 2
        else issue(&tmW0, 0, kc * 64, ((j - 2) & 1) * 128, (j - 2) >> 1);
      }
    }
    CH_TRACE(24, lane == 0);
#pragma unroll 1
    for (int j = 0; j < 8; ++j) issue(&tmXin, p.film_mode ? 1 : 2, ((j % CH_NWG) * (8 / CH_NWG) + j / CH_NWG) * 32, m0, 0);
    if (p.rope) {
      const int tab_row = m0 % p.T;
#pragma unroll 1
      for (int j = 0; j < 8; ++j) issue(&tmTab, 1, ((j % CH_NWG) * (8 / CH_NWG) + j / CH_NWG) * 32, tab_row, 0);
    }
    CH_TRACE(25, lane == 0);
#pragma unroll 1
    for (int h = g_lo; h < g_hi; ++h) {
#pragma unroll 1
      for (int j = 0; j < G1S; ++j) {
        if (CL == 2) issue(&tmW1, 0, (j & 3) * 64, h * 128 + (int)crank * 64, 0);       // box = both planes of my 64 rows
        else issue(&tmW1, 0, ((j >> 1) & 3) * 64, h * 128, j & 1);
      }
    }
    CH_TRACE(26, lane == 0);
    if (n_v > 0) {
      umma::mbar_wait(x_stored, 0);   // the x tile written by E_A is globally visible
#pragma unroll 1
      for (int j = 0; j < 8; ++j) issue(&tmXout, 1, ((j % CH_NWG) * (8 / CH_NWG) + j / CH_NWG) * 32, m0, 0);
#pragma unroll 1
      for (int hv = v_lo; hv < v_hi; ++hv) {
#pragma unroll 1
        for (int j = 0; j < G1S; ++j) {
          if (CL == 2) issue(&tmW2, 0, (j & 3) * 64, hv * 128 + (int)crank * 64, 0);
          else issue(&tmW2, 0, ((j >> 1) & 3) * 64, hv * 128, j & 1);
        }
      }
    }
    // every inbound tile of this CTA has been requested: what is left is the tail of GEMM1 / the last epilogue halves, so the
    // stream successor (launched with the programmatic-serialisation attribute when A2P_PDL is on) may be scheduled now: its
    // prologue overlaps this tail instead of idling through it; no-op without the attribute
    pdl_trigger();
  } else if (warp == 1 && (CL == 1 || crank == 0)) {
    // ================= MMA issuer (pair mode: the leader CTA issues for both) =================
    constexpr uint32_t idesc = CL == 2 ? umma::idesc_bf16_f32(256, 128) : umma::idesc_bf16_f32(128, 128);
    constexpr uint32_t idesc0 = CL == 2 ? umma::idesc_bf16_f32(256, 256) : idesc;     // GEMM0 of a pair: one N = 256 accumulator
    constexpr uint32_t TU = CH_TILE >> 4;
    const uint32_t lo0 = umma::desc_lo(slots_u32);
    uint32_t mpar = 0;     // pair mode: bit s = parity of the next MMA use of slot s on m_full[s]
    auto wait_tile = [&](int s_, int q_) {
      if (CL == 2) { umma::mbar_wait(&m_full[s_], (mpar >> s_) & 1u); mpar ^= 1u << s_; }
      else umma::mbar_wait(&s_full[s_], (q_ / CH_NS) & 1);
    };
    // ---- GEMM0 (A0 and W0 both from ring slots).  Single CTA: both 128-column halves of acc0 advance together.
    int q = 0;
    CH_TRACE(16, lane == 0);
#pragma unroll 1
    for (int kc = 0; kc < kc0; ++kc) {
      const int sa0 = q % CH_NS, sa1 = (q + 1) % CH_NS;
      wait_tile(sa0, q);
      wait_tile(sa1, q + 1);
      q += 2;
#pragma unroll 1
      for (int st4 = 0; st4 < G0S - 2; ++st4) {
        const int pw = CL == 2 ? st4 : st4 >> 1, nh = CL == 2 ? 1 : st4 & 1;   // weight plane; last tile of the plane
        {
          const int s = q % CH_NS;
          wait_tile(s, q);
          ++q;
          umma::fence_after();
          if (umma::elect_one()) {
            const uint32_t lob = lo0 + s * TU;
            const uint32_t d = tmem_base + (CL == 2 ? 0 : (st4 & 1) * 128);
            if (pw == 0) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                chain_mma_ss<CL>(d, umma::desc_make(lo0 + sa0 * TU + 2 * k), umma::desc_make(lob + 2 * k), idesc0, (kc | k) != 0 ? 1u : 0u);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                chain_mma_ss<CL>(d, umma::desc_make(lo0 + sa1 * TU + 2 * k), umma::desc_make(lob + 2 * k), idesc0, 1u);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                chain_mma_ss<CL>(d, umma::desc_make(lo0 + sa0 * TU + 2 * k), umma::desc_make(lob + 2 * k), idesc0, 1u);
            }
            chain_commit<CL>(&s_empty[s]);
            if (pw == 1 && nh == 1) {
              chain_commit<CL>(&s_empty[sa0]);
              chain_commit<CL>(&s_empty[sa1]);
              if (kc == kc0 - 1) chain_commit<CL>(acc0_full);
            }
          }
          __syncwarp();
        }
      }
    }
    // ---- GEMM1 and the V job: A = planes in tensor memory, B = weight tile from the ring
    CH_TRACE(17, lane == 0);
#pragma unroll 1
    for (int jh = 0; jh < n_loc; ++jh) {
      const int h = h_lo + jh;           // global accumulator half (columns / V half); buffers and parities follow the local index
      if (jh == 0 && n_g1 > 0) { umma::mbar_wait(a_ready, 0); umma::fence_after(); q = seqG1; CH_TRACE(18, lane == 0); }
      if (jh == n_g1 && n_v > 0) { umma::mbar_wait(a2_ready, 0); umma::fence_after(); q = seqV; }
      const int buf = jh & 1;
      if (jh >= 2) { umma::mbar_wait(&acc1_empty[buf], ((jh >> 1) - 1) & 1); umma::fence_after(); }
      CH_TRACE(44 + jh, lane == 0 && jh < 10);      // accumulator buffer free: the MMAs of local half jh start
      const uint32_t d = tmem_base + 256 + buf * 128;
#pragma unroll 1
      for (int st8 = 0; st8 < G1S; ++st8) {
        const int kc = CL == 2 ? st8 : st8 >> 1, pw = CL == 2 ? 1 : st8 & 1;
        {
          const int s = q % CH_NS;
          wait_tile(s, q);
          ++q;
          umma::fence_after();
          if (umma::elect_one()) {
            const uint32_t low = lo0 + s * TU;
            if (CL == 2) {
              // one slot = [weight plane 0: my 64 rows][weight plane 1: my 64 rows] of this K chunk:
              // plane pairs (act 0, w 0) (act 1, w 0) (act 0, w 1)
#pragma unroll
              for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int ks = kc * 4 + k;
                  chain_mma_ts<CL>(d, tmem_base + (ks >> 1) * 32 + (i == 1 ? 16 : 0) + (ks & 1) * 8,
                                   umma::desc_make(low + (i == 2 ? (8192 >> 4) : 0) + 2 * k), idesc, (kc | i | k) != 0 ? 1u : 0u);
                }
              }
            } else {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                if (pw == 1 && i == 1) break;          // plane pairs (act 0, w 0) (act 1, w 0) | (act 0, w 1)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const int ks = kc * 4 + k;             // 16-element k-step: chunk ks >> 1, half ks & 1
                  chain_mma_ts<CL>(d, tmem_base + (ks >> 1) * 32 + i * 16 + (ks & 1) * 8, umma::desc_make(low + 2 * k), idesc,
                                   (kc | pw | i | k) != 0 ? 1u : 0u);
                }
              }
            }
            chain_commit<CL>(&s_empty[s]);
            if (kc == 3 && pw == 1) {
              chain_commit<CL>(&acc1_full[buf]);
              if (h == g_hi - 1) chain_commit<CL>(a_reads_done);     // last GEMM1 half of this part (never true for a V half)
            }
          }
          __syncwarp();
        }
      }
      CH_TRACE(32 + h, lane == 0 && h < 12);
    }
  } else if (warp == 3) {
    // ================= store warp: x tile -> global by TMA, then hand the staging slots back =================
    if (lane == 0) {
      // one bulk group per chunk, in ring order; a round of CH_NWG chunks (one per warpgroup) is stored and its staging
      // slots handed back to the producer while the warpgroups are still working on the next round (the RoPE-table
      // chunks of pass 3 are waiting for these slots)
#pragma unroll 1
      for (int r = 0; r < 8 / CH_NWG; ++r) {
#pragma unroll 1
        for (int g = 0; g < CH_NWG; ++g) {
          const int j = r * CH_NWG + g;
          umma::mbar_wait(&x_written[j], 0);
          if (x_writer) {
            tma_store_2d(&tmXout, slots_u32 + ((seqEA + j) % CH_NS) * CH_TILE, (g * (8 / CH_NWG) + r) * 32, m0);
            bulk_commit();
          }
        }
        static_assert(CH_NWG == 4, "staggered waits below assume 4 stores per round");
        bulk_wait_read<3>(); umma::mbar_arrive(&s_empty[(seqEA + r * CH_NWG + 0) % CH_NS]);
        bulk_wait_read<2>(); umma::mbar_arrive(&s_empty[(seqEA + r * CH_NWG + 1) % CH_NS]);
        bulk_wait_read<1>(); umma::mbar_arrive(&s_empty[(seqEA + r * CH_NWG + 2) % CH_NS]);
        bulk_wait_read<0>(); umma::mbar_arrive(&s_empty[(seqEA + r * CH_NWG + 3) % CH_NS]);
      }
      bulk_wait_all();           // globally visible (the V job re-reads the tile; nothing may be in flight at exit)
      if (n_v > 0) umma::mbar_arrive(x_stored);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ================= epilogue warpgroups (CH_NWG x 4 warps; thread = TMEM lane = row) =================
    constexpr int NCH = 8 / CH_NWG;                        // 32-column chunks per warpgroup in E_A
    const int wg = (warp - 4) >> 2;
    const int wq = warp & 3;
    const int trow = wq * 32 + lane;                       // row inside the tile == TMEM lane
    const int et = threadIdx.x - 128;                      // 0 .. 128 * CH_NWG - 1
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const uint32_t stg_u32 = umma::smem_u32(sStg) + (warp - 4) * 2048;   // this warp's [32][16] fp32 staging tile
    const uint32_t row_off = trow * 128;                   // byte offset of this thread's row inside a slot
    const int rx = trow & 7;
    const uint32_t pb_u32 = umma::smem_u32(sPB);
    const int grow_own = m0 + trow;
    const int s_first = m0 / p.T;
    const int sl = grow_own / p.T - s_first;               // 0 or 1 (T >= 128)
    constexpr int NE = 128 * CH_NWG;

    // ---------------- parameter block -> shared memory (hidden behind GEMM0)
    {
      for (int i = et; i < 256; i += NE) {
        sPB[CH_PB_BIAS0 + i] = p.bias0 ? __ldg(p.bias0 + i) : 0.f;
        sPB[CH_PB_LNW + i] = p.ln_mode ? __ldg(p.ln_w + i) : 1.f;
        sPB[CH_PB_LNB + i] = p.ln_mode ? __ldg(p.ln_b + i) : 0.f;
        sPB[CH_PB_BIAS2 + i] = (p.vjob && p.bias2) ? __ldg(p.bias2 + i) : 0.f;
      }
      for (int i = et; i < 1024; i += NE) sPB[CH_PB_BIAS1 + i] = (p.bias1 && i < p.N1) ? __ldg(p.bias1 + i) : 0.f;
      if (p.film_mode) {
        const int n_samp = (p.M + p.T - 1) / p.T;
        for (int i = et; i < 1024; i += NE) {
          const int s = i >> 9, j = i & 511;                // sample-local index, [scale 256 | shift 256]
          const int smp = ::min(s_first + s, n_samp - 1);
          const float* fs = p.film + (long long)smp * p.film_ld;
          sPB[CH_PB_FILM + i] = __ldg(fs + (j < 256 ? p.film_scale_off + j : p.film_shift_off + (j - 256)));
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(NE) : "memory");
    }
    CH_TRACE(1, et == 0);

    // ---------------- E_A pass 1: x = x + film(acc0 + b0); x tile out by TMA store; x kept in TMEM; row sums
    umma::mbar_wait(acc0_full, 0);
    umma::fence_after();
    CH_TRACE(2, et == 0);
    float sum = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < NCH; ++cc) {
      const int c = wg * NCH + cc;
      const int qx = seqEA + cc * CH_NWG + wg;
      const int sx = qx % CH_NS;
      float v[32];
      umma::tmem_ld32(tmem_base + lane_addr + c * 32, v);
      umma::mbar_wait(&s_full[sx], (qx / CH_NS) & 1);
      umma::tmem_ld_wait();
      const uint32_t srow = slots_u32 + sx * CH_TILE + row_off;
      const uint32_t pbc = pb_u32 + (CH_PB_BIAS0 + c * 32) * 4;
      const uint32_t pfs = pb_u32 + (CH_PB_FILM + sl * 512 + c * 32) * 4;
      // 8 columns at a time: all shared-memory loads of the group are issued before the first dependent instruction
#pragma unroll
      for (int gq = 0; gq < 4; ++gq) {
        float4 bb[2], xo[2], sc[2], sh[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) bb[k] = lds128(pbc + (gq * 2 + k) * 16);
        if (p.film_mode) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            xo[k] = lds128(srow + (((gq * 2 + k) ^ rx) << 4));
            sc[k] = lds128(pfs + (gq * 2 + k) * 16);
            sh[k] = lds128(pfs + 1024 + (gq * 2 + k) * 16);
          }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int u = gq * 2 + k;
          f2 o01 = add2(f2{v[4 * u], v[4 * u + 1]}, f2{bb[k].x, bb[k].y}), o23 = add2(f2{v[4 * u + 2], v[4 * u + 3]}, f2{bb[k].z, bb[k].w});
          if (p.film_mode) {   // x + ((scale + 1) * (acc + b) + shift), same operation order as the unfused epilogue
            o01 = add2(f2{xo[k].x, xo[k].y}, fma2(add2(f2{sc[k].x, sc[k].y}, bc2(1.f)), o01, f2{sh[k].x, sh[k].y}));
            o23 = add2(f2{xo[k].z, xo[k].w}, fma2(add2(f2{sc[k].z, sc[k].w}, bc2(1.f)), o23, f2{sh[k].z, sh[k].w}));
          }
          v[4 * u] = o01.x; v[4 * u + 1] = o01.y; v[4 * u + 2] = o23.x; v[4 * u + 3] = o23.y;
          sts128(srow + ((u ^ rx) << 4), make_float4(o01.x, o01.y, o23.x, o23.y));
        }
      }
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 32; ++j) s4[j & 3] += v[j];
      sum += (s4[0] + s4[1]) + (s4[2] + s4[3]);
      tmem_st32(tmem_base + lane_addr + c * 32, v);
      // the chunk leaves by TMA store from the STORE WARP (warp 3): this thread only makes its writes visible to the async
      // proxy and signals; waiting for the stores to drain the staging slots is nobody's critical path
      umma::fence_proxy_async();
      umma::mbar_arrive(&x_written[cc * CH_NWG + wg]);
    }
    tmem_st_wait();
    CH_TRACE(3, et == 0);
    // ---------------- row statistics (two-pass LayerNorm; every warpgroup owns 256 / CH_NWG columns)
    float mean = 0.f, rstd = 1.f;
    if (p.ln_mode) {
      sRed[wg * 128 + trow] = sum;
      asm volatile("bar.sync 1, %0;" ::"n"(NE) : "memory");
      float tot = 0.f;
#pragma unroll
      for (int g = 0; g < CH_NWG; ++g) tot += sRed[g * 128 + trow];
      mean = tot / 256.f;
      asm volatile("bar.sync 1, %0;" ::"n"(NE) : "memory");
      float qs = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < NCH; ++cc) {
        float v[32];
        umma::tmem_ld32(tmem_base + lane_addr + (wg * NCH + cc) * 32, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d_ = v[j] - mean; qs += d_ * d_; }
      }
      sRed[wg * 128 + trow] = qs;
      asm volatile("bar.sync 1, %0;" ::"n"(NE) : "memory");
      tot = 0.f;
#pragma unroll
      for (int g = 0; g < CH_NWG; ++g) tot += sRed[g * 128 + trow];
      rstd = rsqrtf(tot / 256.f + 1e-5f);
    }
    CH_TRACE(4, et == 0);
    ChainEpiCtx ectx{tmem_base + lane_addr, slots_u32, row_off, pb_u32, s_full, s_empty, wg, trow, p.ln_mode, mean, rstd, peer};
    chain_emit_planes<CL>(ectx, p.rope, -1, seqTab);
    __syncwarp();
    if (lane == 0) arrive_at_leader<CL>(a_ready, crank);
    CH_TRACE(5, et == 0);

    // ---------------- E_B: accumulator half h is drained by the warpgroup PAIR (h & 1): warpgroup `sub` of the pair takes 64 of
    //                  its 128 columns (four 16-column chunks through this warp's staging tile)
    const int pair = wg >> 1, sub = wg & 1;
    const int r8 = lane >> 2, u4 = lane & 3;               // transposed phase: lane -> (row sub-index, 16-byte unit)
    bool vprep_done = false;
    // one extra (empty) iteration for a warpgroup pair that drains no V half of this part: its plane chunks are still needed
#pragma unroll 1
    for (int jh = pair; jh < n_loc || (n_v > 0 && !vprep_done); jh += 2) {
      const int h = h_lo + jh;
      const bool vj = h >= NH1;
      if (vj && !vprep_done) {
        if (n_g1 > 0) { umma::mbar_wait(a_reads_done, 0); umma::fence_after(); }   // every GEMM1 MMA of this part has read the rotated planes
        chain_emit_planes<CL>(ectx, 0, seqVx, 0);
        __syncwarp();
        if (lane == 0) arrive_at_leader<CL>(a2_ready, crank);
        vprep_done = true;
      }
      if (jh >= n_loc) break;
      const int buf = jh & 1;
      umma::mbar_wait(&acc1_full[buf], (jh >> 1) & 1);
      umma::fence_after();
      CH_TRACE(6 + h, trow == 0 && sub == 0 && h < 10);      // acc1_full(h) seen by the draining warpgroup pair
      const int remap_first = p.remap_rps > 0 ? (m0 / p.remap_rps + 1) * p.remap_pad : 0;   // pad rows in front of the tile's first sample
      const int remap_edge = p.remap_rps > 0 ? (m0 / p.remap_rps + 1) * p.remap_rps : 0x7fffffff;   // first row of the next sample (T >= 128)
      // the accumulator chunks are read one ahead: the tcgen05.ld of chunk k + 1 is in flight while chunk k is processed
      float v[16];
      const uint32_t acc_addr = tmem_base + lane_addr + 256 + buf * 128 + sub * 64;
      const bool tma_out = p.eb_tma && !vj;
      tmem_ld16(acc_addr, v);
#pragma unroll 1
      for (int k = 0; k < 4; ++k) {
        const int c16 = sub * 4 + k;                       // 16-column chunk of the half
        umma::tmem_ld_wait();
        if (k == 3) { umma::fence_before(); __syncwarp(); if (lane == 0) arrive_at_leader<CL>(&acc1_empty[buf], crank); }
        if (p.eb_tma && !tma_out) {      // the TMA stores that read this warp's staging tile last have finished reading it
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
        }
        if (tma_out) {
          // ---- thread = row: bias / scale / GELU and the plane split in registers, the two bf16 planes of the [32 rows x 16
          // columns] chunk staged row-major ([32][32 B] each) and written by ONE TMA store per plane.  The transposed path
          // below spends 8 + 8 shared-memory wavefronts and 8 global-store instructions of 8 partial lines each per chunk and
          // warp -- the E_B phase was bound by that LSU traffic, not by its arithmetic (profiles/r02_chain_eb_timeline.txt).
          const int col = h * 128 + c16 * 16;
          const float osc = (p.scale_ncols != 0 && col >= p.scale_ncols) ? 1.f : p.out_scale;
          const uint32_t pbb = pb_u32 + (CH_PB_BIAS1 + col) * 4;
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 bb = lds128(pbb + q * 16);      // same address in every lane: broadcast
            f2 o01 = add2(f2{v[4 * q], v[4 * q + 1]}, f2{bb.x, bb.y}), o23 = add2(f2{v[4 * q + 2], v[4 * q + 3]}, f2{bb.z, bb.w});
            if (p.gelu) {
              o01 = gelu_as2(o01); o23 = gelu_as2(o23);
            } else {
              o01 = mul2(o01, bc2(osc)); o23 = mul2(o23, bc2(osc));
            }
            split_act_pair(o01.x, o01.y, hi[2 * q], lo[2 * q]);
            split_act_pair(o23.x, o23.y, hi[2 * q + 1], lo[2 * q + 1]);
          }
          if (k < 3) tmem_ld16(acc_addr + (k + 1) * 16, v);
          // the previous chunk's TMA stores have finished READING the staging tile (they had the arithmetic above to do so)
          if (lane == 0) bulk_wait_read<0>();
          __syncwarp();
          const uint32_t srow = stg_u32 + lane * 32;
          sts128u(srow, hi[0], hi[1], hi[2], hi[3]);
          sts128u(srow + 16, hi[4], hi[5], hi[6], hi[7]);
          sts128u(srow + 1024, lo[0], lo[1], lo[2], lo[3]);
          sts128u(srow + 1040, lo[4], lo[5], lo[6], lo[7]);
          umma::fence_proxy_async();
          __syncwarp();
          if (lane == 0 && col < p.N1 && m0 + wq * 32 < p.M) {
            tma_store_3d(&tmC, stg_u32, col, m0 + wq * 32, 0);
            tma_store_3d(&tmC, stg_u32 + 1024, col, m0 + wq * 32, 1);
            bulk_commit();
          }
          continue;
        }
        {
          const uint32_t srow = stg_u32 + lane * 64;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            sts128(srow + ((q ^ ((lane >> 1) & 3)) << 4), make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
          if (k < 3) tmem_ld16(acc_addr + (k + 1) * 16, v);
        }
        __syncwarp();
        if (!vj) {
          // rows = tokens, columns = output features: 8 rows x 64 B per warp instruction
          const int col = h * 128 + c16 * 16 + u4 * 4;
          const bool col_ok = col < p.N1;
          const float4 bb = lds128(pb_u32 + (CH_PB_BIAS1 + (col_ok ? col : 0)) * 4);
          const float osc = (p.scale_ncols != 0 && col >= p.scale_ncols) ? 1.f : p.out_scale;
#pragma unroll 2
          for (int it = 0; it < 4; ++it) {
            const int r = it * 8 + r8;
            const float4 a = lds128(stg_u32 + r * 64 + ((u4 ^ ((r >> 1) & 3)) << 4));
            f2 o01 = add2(f2{a.x, a.y}, f2{bb.x, bb.y}), o23 = add2(f2{a.z, a.w}, f2{bb.z, bb.w});
            const int grow = m0 + wq * 32 + r;
            if (p.gelu) {
              o01 = gelu_as2(o01); o23 = gelu_as2(o23);
            } else {
              o01 = mul2(o01, bc2(osc)); o23 = mul2(o23, bc2(osc));
            }
            if (!(col_ok && grow < p.M)) continue;
            const long long orow = grow + remap_first + (grow >= remap_edge ? p.remap_pad : 0);
            uint32_t h0, l0, h1, l1;
            split_act_pair(o01.x, o01.y, h0, l0);
            split_act_pair(o23.x, o23.y, h1, l1);
            __nv_bfloat16* dst = p.Cp + orow * p.ldcp + col;
            *reinterpret_cast<uint2*>(dst) = make_uint2(h0, h1);
            *reinterpret_cast<uint2*>(dst + p.cp_plane_stride) = make_uint2(l0, l1);
          }
        } else {
          // V job: the accumulator is V[tokens, channels]; store V^T: lane = (channel, token half), 16 consecutive tokens = 32 B per plane
          const int chl = lane & 15, th = lane >> 4;
          const int ch = (h - NH1) * 128 + c16 * 16 + chl;
          const float b2 = lds32(pb_u32 + (CH_PB_BIAS2 + ch) * 4);
          float t[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int r = th * 16 + j;
            t[j] = lds32(stg_u32 + r * 64 + ((((chl >> 2) ^ ((r >> 1) & 3)) << 4) | ((chl & 3) << 2))) + b2;
          }
          const int tok0 = m0 + wq * 32 + th * 16;
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_act_pair(t[2 * e], t[2 * e + 1], hi[e], lo[e]);
          if (p.eb_tma) {
            // the V^T planes of the chunk ([16 channels][32 tokens] bf16 each) go back into the staging tile and leave by TMA
            __syncwarp();                                  // every lane has read its column of the fp32 tile
            const uint32_t srow = stg_u32 + chl * 64 + th * 32;
            sts128u(srow, hi[0], hi[1], hi[2], hi[3]);
            sts128u(srow + 16, hi[4], hi[5], hi[6], hi[7]);
            sts128u(srow + 1024, lo[0], lo[1], lo[2], lo[3]);
            sts128u(srow + 1040, lo[4], lo[5], lo[6], lo[7]);
            umma::fence_proxy_async();
            __syncwarp();
            if (lane == 0 && m0 + wq * 32 < p.M) {
              tma_store_3d(&tmVt, stg_u32, m0 + wq * 32, (h - NH1) * 128 + c16 * 16, 0);
              tma_store_3d(&tmVt, stg_u32 + 1024, m0 + wq * 32, (h - NH1) * 128 + c16 * 16, 1);
              bulk_commit();
            }
          } else {
            __nv_bfloat16* dst = p.Vt + (long long)ch * p.ldvt + tok0;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              if (tok0 + 8 * u < p.M) {   // M % 8 == 0 (checked by the launcher)
                *reinterpret_cast<uint4*>(dst + 8 * u) = make_uint4(hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
                *reinterpret_cast<uint4*>(dst + p.vt_plane_stride + 8 * u) = make_uint4(lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
              }
            }
          }
        }
        __syncwarp();
      }
      CH_TRACE(54 + h, trow == 0 && sub == 0 && h < 10);     // half h drained
    }
    if (p.eb_tma && lane == 0) bulk_wait_all();     // nothing may still read this CTA's shared memory at exit
    CH_TRACE(30, et == 0);
  }
  __syncthreads();
  if (CL == 2) cluster_sync_all();     // no CTA leaves while the leader may still issue MMAs on / commit into its peer
  if (warp == 2) {
    umma::fence_after();
    if (CL == 2) tmem_dealloc_pair<512>(tmem_base); else umma::tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
inline int make_tmap_f32_2d(CUtensorMap* tm, const void* base, long long cols, long long rows, long long ld, int box_cols, int box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) A2P_FAIL("cuTensorMapEncodeTiled entry point not available");
  if ((ld * 4) % 16 || (reinterpret_cast<uintptr_t>(base) % 16)) A2P_FAIL("TMA fp32 operand not 16-byte aligned (ld=%lld)", ld);
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) A2P_FAIL("cuTensorMapEncodeTiled(fp32) failed (%d) cols=%lld rows=%lld ld=%lld", (int)r, cols, rows, ld);
  return 0;
}

#ifndef A2P_CHAIN_PAIR_DEFAULT
#define A2P_CHAIN_PAIR_DEFAULT 2
#endif
// A2P_CHAIN_PAIR: 1 = CTA-pair mode for every chain launch (cta_group::2 MMAs, every CTA streams half of each weight tile),
// 0 = one CTA per tile, 2 = auto: pairs only for the launches whose GEMM0 streams a long weight matrix (K0 >= 1024: the FFN2
// chains, the only ones the pair mode speeds up -- profiles/r02_chain_pair_mode.txt)
inline int& chain_mode_override() { static int v = 0; return v; }   // tests: 1 / 2 forces the mode of the next launches
inline int& chain_nsplit_override() { static int v = 0; return v; }  // tests: > 0 forces the N split of the next launches
inline int chain_cluster_size(int K0) {
  if (chain_mode_override() > 0) return chain_mode_override();
  static int v = -1;
  if (v < 0) { const char* e = getenv("A2P_CHAIN_PAIR"); v = e ? atoi(e) : A2P_CHAIN_PAIR_DEFAULT; }
  if (v == 2) return K0 >= 1024 ? 2 : 1;
  return v == 1 ? 2 : 1;
}

// N split policy of the chain launches (used by engine.cu forward_core): a launch of `tiles` 128-row tiles whose GEMM1 / V job has n_acc 128-column
// accumulator halves is cut into s parts per tile when the step's `concurrent` forwards together leave SMs idle.  The parts
// repeat GEMM0 + E_A, so the split pays where that prefix is short against the GEMM1 / E_B tail: s = the largest divisor of
// n_acc with n_acc / s >= 2 halves per part and tiles * s * concurrent <= budget CTAs, budget = 160 (the machine) for launches
// with K0 <= 256 (the FFN1 + GELU launch: 45 -> 33 -> 27 us at s = 1 / 2 / 4) and 80 for the K0 = 1024 launches (FFN2 -> LN ->
// Q|K|V: 57 -> 49 -> 42 us at s = 1 / 2 / 3, but their redundant prefix is 60 % of the launch).  Measured on the loop
// (profiles/r02_chain_nsplit_pdl.txt): B = 4 + 7 %, B = 8 + 1..3 %, B >= 16 never splits.
// A2P_CHAIN_NSPLIT=0 off, 1 auto (default), n >= 2 force (capped by n_acc); A2P_CHAIN_SPLIT_BUDGET (CTAs, default 160),
// A2P_CHAIN_SPLIT_MINH (halves per part, default 2).
inline int chain_nsplit_for(int tiles, int n_acc, int concurrent, int K0) {
  static int mode = -1, budget = 160, minh = 2;
  if (mode < 0) {
    const char* e = getenv("A2P_CHAIN_NSPLIT"); mode = e ? atoi(e) : 1; if (mode < 0) mode = 0;
    if ((e = getenv("A2P_CHAIN_SPLIT_BUDGET"))) budget = atoi(e);
    if ((e = getenv("A2P_CHAIN_SPLIT_MINH"))) minh = atoi(e) > 0 ? atoi(e) : 1;
  }
  if (mode == 0 || n_acc < 2) return 1;
  if (mode >= 2) return mode < n_acc ? mode : n_acc;
  const long long cap = K0 <= 256 ? budget : budget / 2;
  int best = 1;
  for (int s_ = 2; s_ <= 4 && s_ <= n_acc; ++s_)
    if (n_acc % s_ == 0 && n_acc / s_ >= minh && (long long)tiles * s_ * (concurrent > 0 ? concurrent : 1) <= cap) best = s_;
  return best;
}

struct ChainOperands {
  const __nv_bfloat16* A0; long long a0_rows, a0_ld, a0_plane_stride;   // [2][a0_rows][a0_ld], K0 valid columns
  const __nv_bfloat16* W0; long long w0_plane_stride;                   // [2][256][K0]
  const __nv_bfloat16* W1; long long w1_plane_stride;                   // [2][N1][256]
  const __nv_bfloat16* W2; long long w2_plane_stride;                   // [2][256][256] (null without a V job)
  float* x;                                                             // [M][256] fp32 residual stream (read if film_mode, always written)
  float* x_out;                                                         // null: update x in place; else the new stream is written here (required by nsplit > 1 with film_mode)
  const float* rope_ext; long long rope_ext_rows;                       // [T + 128][256] fp32: (cos, sin) pairs of position (row % T)
};

inline int launch_umma_chain(const ChainOperands& o, const ChainParams& p, cudaStream_t st, int force_cl = 0) {
  if (p.K0 % 8 || p.N1 % 8 || p.N1 <= 0 || p.N1 > 1024) A2P_FAIL("chain: bad K0=%d / N1=%d", p.K0, p.N1);
  if (p.T < 128 || p.M % 8) A2P_FAIL("chain: needs T >= 128 and M %% 8 == 0 (T=%d M=%d)", p.T, p.M);
  if (p.vjob && (!o.W2 || !p.Vt)) A2P_FAIL("chain: V job needs W2 and Vt");
  if (p.rope && (!o.rope_ext || o.rope_ext_rows < p.T + 128)) A2P_FAIL("chain: RoPE needs the extended table (T + 128 rows)");
  CUtensorMap tA0, tW0, tW1, tW2, tXin, tXout, tTab, tC, tVt;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  const int cl = force_cl > 0 ? force_cl : chain_cluster_size(p.K0);
  const int n_acc = ceil_div(p.N1, 128) + (p.vjob ? 2 : 0);
  const int nsp = p.nsplit > 1 ? p.nsplit : 1;
  if (nsp > n_acc) A2P_FAIL("chain: nsplit=%d exceeds the %d accumulator halves of the launch", nsp, n_acc);
  if (nsp > 1 && p.film_mode && (!o.x_out || o.x_out == o.x)) A2P_FAIL("chain: nsplit > 1 needs a separate x_out buffer");
  // pair mode: a CTA loads its 128 of the 256 W0 rows (one plane per box) and, for a 128-column accumulator half of GEMM1 / the
  // V job, both planes of its 64 rows in ONE box [64 k][64 rows][2 planes] = one 16 KB slot
  const int w1rows = cl == 2 ? 64 : 128, w1planes = cl == 2 ? 2 : 1;
  A2P_TRY(make_tmap_bf16_3d(&tA0, o.A0, p.K0, o.a0_rows, 2, o.a0_ld, o.a0_plane_stride, 64, 128, sw));
  A2P_TRY(make_tmap_bf16_3d(&tW0, o.W0, p.K0, 256, 2, p.K0, o.w0_plane_stride, 64, 128, sw));
  A2P_TRY(make_tmap_bf16_3d(&tW1, o.W1, 256, p.N1, 2, 256, o.w1_plane_stride, 64, w1rows, sw, w1planes));
  if (p.vjob) A2P_TRY(make_tmap_bf16_3d(&tW2, o.W2, 256, 256, 2, 256, o.w2_plane_stride, 64, w1rows, sw, w1planes));
  else tW2 = tW1;
  A2P_TRY(make_tmap_f32_2d(&tXin, o.x, 256, p.M, 256, 32, 128));
  if (o.x_out && o.x_out != o.x) A2P_TRY(make_tmap_f32_2d(&tXout, o.x_out, 256, p.M, 256, 32, 128));
  else tXout = tXin;
  if (p.rope) A2P_TRY(make_tmap_f32_2d(&tTab, o.rope_ext, 256, o.rope_ext_rows, 256, 32, 128));
  else tTab = tXin;
  // E_B output planes by TMA store ([16 columns x 32 rows] boxes per plane) unless the rows are remapped (final_layer -> padded
  // TCN layout: a 32-row box may straddle a sample boundary) or the columns do not come in whole 16-column chunks
  static const bool eb_tma_env = !(getenv("A2P_CHAIN_EB_TMA") && atoi(getenv("A2P_CHAIN_EB_TMA")) == 0);
  ChainParams pp = p;
  pp.eb_tma = (eb_tma_env && p.remap_rps == 0 && p.N1 % 16 == 0 && p.ldcp % 8 == 0 && p.cp_plane_stride % 8 == 0 &&
               reinterpret_cast<uintptr_t>(p.Cp) % 16 == 0) ? 1 : 0;
  if (pp.eb_tma && p.vjob && (p.ldvt % 8 || p.vt_plane_stride % 8 || reinterpret_cast<uintptr_t>(p.Vt) % 16)) pp.eb_tma = 0;
  if (pp.eb_tma) A2P_TRY(make_tmap_bf16_3d(&tC, p.Cp, p.N1, p.M, 2, p.ldcp, p.cp_plane_stride, 16, 32, CU_TENSOR_MAP_SWIZZLE_NONE));
  else tC = tXin;
  if (pp.eb_tma && p.vjob) A2P_TRY(make_tmap_bf16_3d(&tVt, p.Vt, p.M, 256, 2, p.ldvt, p.vt_plane_stride, 32, 16, CU_TENSOR_MAP_SWIZZLE_NONE));
  else tVt = tXin;
  const int tiles = ceil_div(p.M, 128);
  if (cl == 2) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((tiles + 1) / 2 * 2 * nsp); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = CH_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    int na = 1;
    if (pdl_enabled()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    A2P_CUDA(cudaLaunchKernelEx(&cfg, umma_chain_kernel<2>, tA0, tW0, tW1, tW2, tXin, tXout, tTab, tC, tVt, pp));
  } else {
    A2P_CUDA(launch_pdl(umma_chain_kernel<1>, dim3(tiles * nsp), dim3(CH_THREADS), (size_t)CH_SMEM_BYTES, st, tA0, tW0, tW1, tW2, tXin, tXout,
                        tTab, tC, tVt, pp));
  }
  return 0;
}

inline int init_umma_chain() {
  A2P_CUDA(cudaFuncSetAttribute(umma_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES));
  A2P_CUDA(cudaFuncSetAttribute(umma_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES));
  return 0;
}

// ext[r][2i], ext[r][2i+1] = tab[r % T][i]  (cos, sin), r < T + 128: any 128 consecutive rows of the residual stream
// (position = row % T) map to 128 consecutive rows of the extended table, so one TMA box covers a tile
__global__ void rope_ext_kernel(const float2* __restrict__ tab, float2* __restrict__ ext, int T, int half) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (T + 128) * half) return;
  const int r = idx / half, i = idx - r * half;
  ext[idx] = tab[(long long)(r % T) * half + i];
}

}  // namespace a2p
