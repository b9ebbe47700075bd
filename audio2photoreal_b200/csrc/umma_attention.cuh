// K1 core on the 5th-gen tensor cores: softmax(Q K^T) V for one (sample-row, 64-column head group, 128-query
// tile) per CTA, split-bf16 operands (TERMS planes), fp32 accumulation in TMEM, fp32 online softmax in registers.
//
//   warp 0    TMA producer : Q tile once; per 64-key block the K tile [64 keys][64 cols] and the V^T tile
//                            [64 cols][64 keys] of every plane (SWIZZLE_128B), 2-stage ring
//   warp 1    MMA issuer   : S = Q K^T (M128 x N64, K = dh) into one of 2 TMEM S buffers, then O_blk = P V
//                            (M128 x N=dh, K = 64 keys) into one of 2 TMEM PV buffers; S(i+1) is issued before
//                            PV(i) so the tensor pipe works while the softmax warps process S(i)
//   warp 2    TMEM alloc   : 256 columns (2 x 64 S + 2 x 64 PV)
//   warps 4-7 softmax      : thread = query row; tcgen05.ld S -> running max / exp2 / sum -> P split into bf16
//                            planes written to smem in the UMMA K-major SWIZZLE_128B layout -> after PV(i):
//                            O = O * alpha + PV (O lives in registers, so no TMEM read-modify-write)
// A head group is 64 consecutive model columns (2 heads at dh = 32, 1 head at dh = 64): the heads of a group
// share the K / V^T tiles and are processed back to back.  Keys come from a cached "main" memory plus an
// optional "extra" source (the 2 per-step time tokens, model/diffusion.py:392-393).
// Q must be pre-scaled by log2(e)/sqrt(dh) (done by the Q-projection GEMM epilogue).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"

#ifndef A2P_POLY_OF_4
#define A2P_POLY_OF_4 0
#endif
#ifndef A2P_ATTN_TRACE
#define A2P_ATTN_TRACE 0   // 1: clock64 timeline of CTA (0,0,0) into TcAttnParams::trace (scripts/gpu_attn_trace.py)
#endif

namespace a2p {

struct TcAttnParams {
  int T, R, D, dh, rows_per_branch;
  int q_col0;                 // column of head-group 0 inside the Q tensor
  int k_col0;                 // column of head-group 0 inside the main K tensor (self-attn: D)
  int n_keys, n_extra;
  long long k_row_stride[2];  // rows between consecutive samples in the main K tensor per branch (0 = shared)
  long long v_col_stride[2];  // columns between consecutive samples in the main V^T tensor per branch
  int kx_col0, kx_row_stride; // extra K tensor: column of group 0 (= layer * D), rows per sample (2)
  int vx_row0, vx_col_stride; // extra V^T tensor: row of group 0 (= layer * D), columns per sample (2)
  __nv_bfloat16* Op; long long op_plane_stride; long long o_ld;   // output planes [TERMS][R*T][o_ld]
  float* O;                                                       // optional fp32 output [R*T][o_ld] (tests)
  // split-KV tail (umma_attention2.cuh): work items [0, split_full) are whole tiles; the remaining tiles are cut into
  // split_parts key ranges each (one work item per range) whose partial (max, sum, O) are merged by the last finisher
  float* split_scratch; int* split_counters; int split_full, split_parts, n_qt, n_groups;
  int skew_ns;                                                    // start delay of softmax warpgroup 1 (see kernel)
  long long* trace;                                               // optional [64 iters][16] clock64 timestamps of CTA (0,0,0) (diagnostics)
};

template <int TERMS>
struct TcAttnCfg {
  static constexpr int NPROD = TERMS == 1 ? 1 : (TERMS == 2 ? 3 : 6);
  static constexpr int NPBUF = TERMS == 3 ? 1 : 2;
  static constexpr int Q_BYTES = TERMS * 128 * 128;            // [128 rows][64 cols] bf16 per plane
  static constexpr int KV_STAGE_BYTES = TERMS * 2 * 64 * 128;  // K tile + V^T tile per plane (8 KB each)
  static constexpr int P_BYTES = TERMS * 128 * 128;            // [128 rows][64 keys] bf16 per plane
  static constexpr int SMEM_BYTES = Q_BYTES + 2 * KV_STAGE_BYTES + NPBUF * P_BYTES + 1024 + 512;
  static constexpr int THREADS = 384;                          // 4 role warps + 2 softmax warpgroups
};

__device__ __forceinline__ void st_shared_v4(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// Softmax work split: TWO warpgroups per CTA.  Warpgroup w owns key columns [32w, 32w+32) of every 64-key block and
// runs its own online softmax (running max / sum / O) over that key subset; the two partial results are merged
// once at the end (flash-decoding style).  With one warp per SM sub-partition the softmax was latency-bound
// (issue slots 36 % busy, ncu profiles/r01e); two warps per scheduler hide the TMEM / MUFU / mbarrier latencies
// and halve the per-thread register footprint.
template <int TERMS, int DH>
__global__ void __launch_bounds__(384, 1)
umma_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                 const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV0,
                 const __grid_constant__ CUtensorMap tmV1, const __grid_constant__ CUtensorMap tmKx,
                 const __grid_constant__ CUtensorMap tmVx, TcAttnParams p) {
  using Cfg = TcAttnCfg<TERMS>;
  constexpr int G = 64 / DH;          // heads per group
  constexpr int NPB = Cfg::NPBUF;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Cfg::Q_BYTES;
  uint8_t* sP = sKV + 2 * Cfg::KV_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + NPB * Cfg::P_BYTES);
  uint64_t* q_full = bars;            // [1]
  uint64_t* kv_full = bars + 1;       // [2]
  uint64_t* kv_empty = bars + 3;      // [2]
  uint64_t* s_full = bars + 5;        // [2]
  uint64_t* s_empty = bars + 7;       // [2]       256 arrivals
  uint64_t* p_full = bars + 9;        // [2 wg][2]  128 arrivals each
  uint64_t* p_empty = bars + 13;      // [2 wg][2]
  uint64_t* pv_full = bars + 17;      // [2 wg][2]
  uint64_t* pv_empty = bars + 21;     // [2 wg][2]  128 arrivals each
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 25);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, g = blockIdx.y, r = blockIdx.z;
  const int br = r >= p.rows_per_branch ? 1 : 0;
  const int rr = r - br * p.rows_per_branch;
  const int nb_main = ceil_div(p.n_keys, 64);
  const int n_blocks = nb_main + (p.n_extra > 0 ? 1 : 0);
  const int n_iter = n_blocks * G;

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmQ);
    umma::prefetch_tmap(br ? &tmK1 : &tmK0);
    umma::prefetch_tmap(br ? &tmV1 : &tmV0);
  }
  if (warp == 1 && lane == 0) {
    umma::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(&kv_full[i], 1); umma::mbar_init(&kv_empty[i], 1);
      umma::mbar_init(&s_full[i], 1); umma::mbar_init(&s_empty[i], 256);
    }
    for (int i = 0; i < 4; ++i) {
      umma::mbar_init(&p_full[i], 128); umma::mbar_init(&p_empty[i], 1);
      umma::mbar_init(&pv_full[i], 1); umma::mbar_init(&pv_empty[i], 128);
    }
    umma::fence_barrier_init();
  }
  if (warp == 2) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  const uint32_t tmS = tmem_base, tmPV = tmem_base + 128;   // S buffers: +0, +64 ; PV buffers [wg][b]: +128 + (wg*2+b)*64

  // Single-thread roles run warp-uniformly with only the issuing instructions under elect.sync (see umma_gemm.cuh).
  if (warp == 0) {
    // ================= TMA producer =================
    if (umma::elect_one()) {
      umma::mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int i = 0; i < TERMS; ++i)
        umma::tma_load_3d(&tmQ, q_full, sQ + i * 16384, p.q_col0 + g * 64, r * p.T + q0, i);
    }
    __syncwarp();
    const CUtensorMap* tK = br ? &tmK1 : &tmK0;
    const CUtensorMap* tV = br ? &tmV1 : &tmV0;
    const int k_row_base = (int)(rr * p.k_row_stride[br]);
    const int v_col_base = (int)(rr * p.v_col_stride[br]);
    for (int j = 0; j < n_blocks; ++j) {
      const int st = j & 1;
      umma::mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
      if (umma::elect_one()) {
        umma::mbar_expect_tx(&kv_full[st], Cfg::KV_STAGE_BYTES);
        uint8_t* sk = sKV + st * Cfg::KV_STAGE_BYTES;
        uint8_t* sv = sk + TERMS * 8192;
        if (j < nb_main) {
#pragma unroll
          for (int i = 0; i < TERMS; ++i) umma::tma_load_3d(tK, &kv_full[st], sk + i * 8192, p.k_col0 + g * 64, k_row_base + j * 64, i);
#pragma unroll
          for (int i = 0; i < TERMS; ++i) umma::tma_load_3d(tV, &kv_full[st], sv + i * 8192, v_col_base + j * 64, g * 64, i);
        } else {
#pragma unroll
          for (int i = 0; i < TERMS; ++i) umma::tma_load_3d(&tmKx, &kv_full[st], sk + i * 8192, p.kx_col0 + g * 64, r * p.kx_row_stride, i);
#pragma unroll
          for (int i = 0; i < TERMS; ++i) umma::tma_load_3d(&tmVx, &kv_full[st], sv + i * 8192, r * p.vx_col_stride, p.vx_row0 + g * 64, i);
        }
      }
      __syncwarp();
    }
    pdl_trigger();    // all key blocks requested: the successor's prologue may overlap this CTA's tail (no-op without A2P_PDL)
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idS = umma::idesc_bf16_f32(128, 64);
    constexpr uint32_t idPV = umma::idesc_bf16_f32(128, DH);
    const uint32_t loQ = umma::desc_lo(umma::smem_u32(sQ));
    const uint32_t loKV = umma::desc_lo(umma::smem_u32(sKV));
    const uint32_t loP = umma::desc_lo(umma::smem_u32(sP));
    auto issue_pv = [&](int i) {
      const int j = i / G, hh = i - j * G, b = i & 1, pb = i % NPB;
#pragma unroll
      for (int wg = 0; wg < 2; ++wg) {   // each warpgroup's 32 keys accumulate into their own PV buffer
        umma::mbar_wait(&p_full[wg * 2 + pb], (i / NPB) & 1);
        umma::mbar_wait(&pv_empty[wg * 2 + b], ((i >> 1) & 1) ^ 1);
        umma::fence_after();
        if (umma::elect_one()) {
          const uint32_t lov = loKV + (j & 1) * (Cfg::KV_STAGE_BYTES >> 4) + TERMS * (8192 >> 4) + hh * (DH * 128 >> 4) + 4 * wg;
          const uint32_t lop = loP + pb * (Cfg::P_BYTES >> 4) + 4 * wg;
#pragma unroll
          for (int pr = 0; pr < Cfg::NPROD; ++pr) {
#pragma unroll
            for (int k = 0; k < 2; ++k)
              umma::mma_bf16(tmPV + (wg * 2 + b) * 64, umma::desc_make(lop + prod_a(pr) * (16384 >> 4) + 2 * k),
                             umma::desc_make(lov + prod_b(pr) * (8192 >> 4) + 2 * k), idPV, (pr | k) != 0 ? 1u : 0u);
          }
          umma::mma_commit(&pv_full[wg * 2 + b]);
          umma::mma_commit(&p_empty[wg * 2 + pb]);
          if (wg == 1 && hh == G - 1) umma::mma_commit(&kv_empty[j & 1]);
        }
        __syncwarp();
      }
    };
    const bool tr = A2P_ATTN_TRACE && p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && lane == 0;
    umma::mbar_wait(q_full, 0);
    for (int i = 0; i < n_iter; ++i) {
      const int j = i / G, hh = i - j * G, b = i & 1;
      if (tr && i < 64) p.trace[i * 16 + 8] = clock64();
      if (hh == 0) umma::mbar_wait(&kv_full[j & 1], (j >> 1) & 1);
      if (tr && i < 64) p.trace[i * 16 + 9] = clock64();
      umma::mbar_wait(&s_empty[b], ((i >> 1) & 1) ^ 1);
      umma::fence_after();
      if (umma::elect_one()) {
        const uint32_t lok = loKV + (j & 1) * (Cfg::KV_STAGE_BYTES >> 4) + hh * (DH / 8);
        const uint32_t loq = loQ + hh * (DH / 8);
#pragma unroll
        for (int pr = 0; pr < Cfg::NPROD; ++pr) {
#pragma unroll
          for (int k = 0; k < DH / 16; ++k)
            umma::mma_bf16(tmS + b * 64, umma::desc_make(loq + prod_a(pr) * (16384 >> 4) + 2 * k),
                           umma::desc_make(lok + prod_b(pr) * (8192 >> 4) + 2 * k), idS, (pr | k) != 0 ? 1u : 0u);
        }
        umma::mma_commit(&s_full[b]);
      }
      __syncwarp();
      if (tr && i < 64) p.trace[i * 16 + 10] = clock64();
      if (i > 0) issue_pv(i - 1);
      if (tr && i < 64) p.trace[i * 16 + 11] = clock64();
    }
    issue_pv(n_iter - 1);
  } else if (warp >= 4) {
    // ================= softmax / output (2 warpgroups, key-column split) =================
    const int wg = (warp - 4) >> 2;
    const int wq = warp & 3;
    const int trow = wq * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    float m[G], l[G], o[G][DH];
#pragma unroll
    for (int h = 0; h < G; ++h) {
      m[h] = -INFINITY; l[h] = 0.f;
#pragma unroll
      for (int c = 0; c < DH; ++c) o[h][c] = 0.f;
    }
    float alpha_pend = 1.f;
    const uint32_t sP_u32 = umma::smem_u32(sP);
    // optional start skew of warpgroup 1 (experiment knob; measured no effect -- the softmax is issue/latency bound,
    // not MUFU bound: profiles/r01f_attention_experiments.txt)
    if (wg == 1 && p.skew_ns > 0) __nanosleep(p.skew_ns);
    auto consume_pv = [&](int i, float alpha) {
      const int b = i & 1, hh = i % G;
      umma::mbar_wait(&pv_full[wg * 2 + b], (i >> 1) & 1);
      umma::fence_after();
      float v[DH];
      umma::tmem_ld32(tmPV + lane_addr + (wg * 2 + b) * 64, v);
      if (DH == 64) umma::tmem_ld32(tmPV + lane_addr + (wg * 2 + b) * 64 + 32, v + (DH == 64 ? 32 : 0));
      umma::tmem_ld_wait();
      umma::fence_before();
      umma::mbar_arrive(&pv_empty[wg * 2 + b]);
#pragma unroll
      for (int h = 0; h < G; ++h)
        if (h == hh) {
#pragma unroll
          for (int c = 0; c < DH; ++c) o[h][c] = o[h][c] * alpha + v[c];
        }
    };
    const bool tr = A2P_ATTN_TRACE && p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 128;
    for (int i = 0; i < n_iter; ++i) {
      const int j = i / G, hh = i - j * G, b = i & 1, pb = i % NPB;
      if (tr && i < 64) p.trace[i * 16 + 0] = clock64();
      umma::mbar_wait(&s_full[b], (i >> 1) & 1);
      umma::fence_after();
      if (tr && i < 64) p.trace[i * 16 + 1] = clock64();
      float s[32];
      umma::tmem_ld32(tmS + lane_addr + b * 64 + wg * 32, s);
      umma::tmem_ld_wait();
      if (tr && i < 64) p.trace[i * 16 + 2] = clock64();
      umma::fence_before();
      umma::mbar_arrive(&s_empty[b]);
      const int nv_blk = (j < nb_main) ? ::min(64, p.n_keys - j * 64) : p.n_extra;
      const int nvalid = ::max(0, ::min(32, nv_blk - 32 * wg));   // valid keys among this warpgroup's 32 columns
      float alpha = 1.f, mnew = 0.f;
      if (nvalid > 0) {   // warp-uniform
        if (nvalid < 32) {
#pragma unroll
          for (int c = 0; c < 32; ++c) s[c] = c < nvalid ? s[c] : -INFINITY;
        }
        float mx = s[0];
#pragma unroll
        for (int c = 1; c < 32; ++c) mx = fmaxf(mx, s[c]);
        float mold = 0.f;
#pragma unroll
        for (int h = 0; h < G; ++h)
          if (h == hh) { mold = m[h]; mnew = fmaxf(mold, mx); m[h] = mnew; }
        alpha = umma::ex2_approx(mold - mnew);
      }
      // P planes -> smem (K-major SWIZZLE_128B: 16-byte chunk index XOR (row & 7)); exp2 + split fused per chunk
      if (tr && i < 64) p.trace[i * 16 + 3] = clock64();
      umma::mbar_wait(&p_empty[wg * 2 + pb], ((i / NPB) & 1) ^ 1);
      if (tr && i < 64) p.trace[i * 16 + 4] = clock64();
      const uint32_t prow = sP_u32 + pb * Cfg::P_BYTES + trow * 128;
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t pk[TERMS][4];
        if (ch * 8 < nvalid) {   // warp-uniform: fully masked 8-key chunks cost no exp2
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            // A2P_POLY_OF_4 of every 4 pairs take the FMA-pipe polynomial, the rest MUFU.EX2
            const bool poly = e < A2P_POLY_OF_4;
            const float a = poly ? umma::ex2_poly(s[ch * 8 + 2 * e] - mnew) : umma::ex2_approx(s[ch * 8 + 2 * e] - mnew);
            const float bb = poly ? umma::ex2_poly(s[ch * 8 + 2 * e + 1] - mnew) : umma::ex2_approx(s[ch * 8 + 2 * e + 1] - mnew);
            rs0 += a; rs1 += bb;
            uint32_t sp[TERMS];
            if (TERMS == 3) umma::split_bf16_pair_trunc3(a, bb, sp);   // probabilities in [0,1]: 21+ bits are plenty
            else umma::split_bf16_pair<TERMS>(a, bb, sp);
#pragma unroll
            for (int t = 0; t < TERMS; ++t) pk[t][e] = sp[t];
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e)
#pragma unroll
            for (int t = 0; t < TERMS; ++t) pk[t][e] = 0u;
        }
#pragma unroll
        for (int t = 0; t < TERMS; ++t)
          st_shared_v4(prow + t * 16384 + (((4 * wg + ch) ^ (trow & 7)) << 4), pk[t][0], pk[t][1], pk[t][2], pk[t][3]);
      }
#pragma unroll
      for (int h = 0; h < G; ++h)
        if (h == hh) l[h] = l[h] * alpha + (rs0 + rs1);
      if (tr && i < 64) p.trace[i * 16 + 5] = clock64();
      umma::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      umma::mbar_arrive(&p_full[wg * 2 + pb]);
      if (tr && i < 64) p.trace[i * 16 + 6] = clock64();
      if (i > 0) consume_pv(i - 1, alpha_pend);
      if (tr && i < 64) p.trace[i * 16 + 7] = clock64();
      alpha_pend = alpha;
    }
    consume_pv(n_iter - 1, alpha_pend);
    // ---- merge the two key-subset partials (warpgroup 1 -> smem -> warpgroup 0), normalise, store
    constexpr int XS = G * (DH + 2) + 2;             // floats per row in the exchange buffer (stride 68 or 70: 8-byte aligned)
    float* xch = reinterpret_cast<float*>(sKV) + trow * XS;   // K/V stages are dead: every MMA has completed
    if (wg == 1) {
#pragma unroll
      for (int h = 0; h < G; ++h) {
        xch[h * (DH + 2)] = m[h];
        xch[h * (DH + 2) + 1] = l[h];
#pragma unroll
        for (int c = 0; c < DH; ++c) xch[h * (DH + 2) + 2 + c] = o[h][c];
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the 8 softmax warps only
    const int row = q0 + trow;
    if (wg == 0 && row < p.T) {
      const long long grow = (long long)r * p.T + row;
#pragma unroll
      for (int h = 0; h < G; ++h) {
        const float m1 = xch[h * (DH + 2)], l1 = xch[h * (DH + 2) + 1];
        const float mm = fmaxf(m[h], m1);
        const float w0 = umma::ex2_approx(m[h] - mm), w1 = umma::ex2_approx(m1 - mm);   // exp2(-inf) = 0 for an empty subset
        const float inv = 1.f / (l[h] * w0 + l1 * w1);
        const float f0 = w0 * inv, f1 = w1 * inv;
#pragma unroll
        for (int c = 0; c < DH; ++c) o[h][c] = o[h][c] * f0 + xch[h * (DH + 2) + 2 + c] * f1;
        const int col = g * 64 + h * DH;
        if (p.O) {
          float* dst = p.O + grow * p.o_ld + col;
#pragma unroll
          for (int c = 0; c < DH; c += 4)
            *reinterpret_cast<float4*>(dst + c) = make_float4(o[h][c], o[h][c + 1], o[h][c + 2], o[h][c + 3]);
        }
        if (p.Op) {
#pragma unroll
          for (int c = 0; c < DH; c += 8) {
            uint32_t pk[TERMS][4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              uint32_t sp[TERMS];
              umma::split_bf16_pair<TERMS>(o[h][c + 2 * e], o[h][c + 2 * e + 1], sp);
#pragma unroll
              for (int t = 0; t < TERMS; ++t) pk[t][e] = sp[t];
            }
#pragma unroll
            for (int t = 0; t < TERMS; ++t)
              *reinterpret_cast<uint4*>(p.Op + t * p.op_plane_stride + grow * p.o_ld + col + c) = make_uint4(pk[t][0], pk[t][1], pk[t][2], pk[t][3]);
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 2) {
    umma::fence_after();
    umma::tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
struct TcAttnOperands {
  const __nv_bfloat16* Q; long long q_rows, q_ld, q_plane_stride;                 // [P][q_rows][q_ld]
  const __nv_bfloat16* K[2]; long long k_rows[2], k_ld[2], k_plane_stride[2];     // main K per branch
  const __nv_bfloat16* Vt[2]; long long vt_cols[2], vt_ld[2], vt_plane_stride[2]; // main V^T per branch: [P][vt_rows][vt_ld]
  long long vt_rows;
  const __nv_bfloat16* Kx; long long kx_rows, kx_ld, kx_plane_stride;             // extra K  [P][kx_rows][kx_ld]
  const __nv_bfloat16* Vx; long long vx_rows, vx_cols, vx_ld, vx_plane_stride;    // extra V^T [P][vx_rows][vx_ld]
};

template <int TERMS, int DH>
int launch_umma_attn_t(const TcAttnOperands& o, const TcAttnParams& p, cudaStream_t st) {
  using Cfg = TcAttnCfg<TERMS>;
  CUtensorMap tq, tk[2], tv[2], tkx, tvx;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  A2P_TRY(make_tmap_bf16_3d(&tq, o.Q, o.q_ld, o.q_rows, TERMS, o.q_ld, o.q_plane_stride, 64, 128, sw));
  for (int b = 0; b < 2; ++b) {
    const int s = o.K[b] ? b : 0;
    A2P_TRY(make_tmap_bf16_3d(&tk[b], o.K[s], o.k_ld[s], o.k_rows[s], TERMS, o.k_ld[s], o.k_plane_stride[s], 64, 64, sw));
    A2P_TRY(make_tmap_bf16_3d(&tv[b], o.Vt[s], o.vt_cols[s], o.vt_rows, TERMS, o.vt_ld[s], o.vt_plane_stride[s], 64, 64, sw));
  }
  if (o.Kx) {
    A2P_TRY(make_tmap_bf16_3d(&tkx, o.Kx, o.kx_ld, o.kx_rows, TERMS, o.kx_ld, o.kx_plane_stride, 64, 64, sw));
    A2P_TRY(make_tmap_bf16_3d(&tvx, o.Vx, o.vx_cols, o.vx_rows, TERMS, o.vx_ld, o.vx_plane_stride, 64, 64, sw));
  } else {
    tkx = tk[0]; tvx = tv[0];
  }
  dim3 grid(ceil_div(p.T, 128), p.D / 64, p.R);
  A2P_CUDA(launch_pdl(umma_attn_kernel<TERMS, DH>, grid, dim3(Cfg::THREADS), (size_t)Cfg::SMEM_BYTES, st, tq, tk[0], tk[1], tv[0],
                      tv[1], tkx, tvx, p));
  return 0;
}

inline int launch_umma_attn(int terms, const TcAttnOperands& o, const TcAttnParams& p, cudaStream_t st) {
  if (p.dh == 32) {
    if (terms == 2) return launch_umma_attn_t<2, 32>(o, p, st);
    if (terms == 3) return launch_umma_attn_t<3, 32>(o, p, st);
    if (terms == 1) return launch_umma_attn_t<1, 32>(o, p, st);
  } else if (p.dh == 64) {
    if (terms == 2) return launch_umma_attn_t<2, 64>(o, p, st);
    if (terms == 3) return launch_umma_attn_t<3, 64>(o, p, st);
    if (terms == 1) return launch_umma_attn_t<1, 64>(o, p, st);
  }
  A2P_FAIL("umma_attn: unsupported terms=%d dh=%d", terms, p.dh);
}

inline int init_umma_attn() {
#define A2P_SET(T, H) A2P_CUDA(cudaFuncSetAttribute(umma_attn_kernel<T, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcAttnCfg<T>::SMEM_BYTES));
  A2P_SET(1, 32) A2P_SET(2, 32) A2P_SET(3, 32) A2P_SET(1, 64) A2P_SET(2, 64) A2P_SET(3, 64)
#undef A2P_SET
  return 0;
}

// fp32 [rows][cols] (row stride ld)  ->  bf16 planes of the TRANSPOSE: dst[t][c][col_base(r)] with
// column index = (r / rows_per_sample) * sample_col_stride + r % rows_per_sample; untouched pad columns must be
// zeroed by the caller once (P = 0 times non-finite garbage would poison the PV product).
template <int TERMS>
__global__ void transpose_split_kernel(const float* __restrict__ src, long long ld, __nv_bfloat16* __restrict__ dst,
                                       long long plane_stride, long long dst_ld, int rows, int cols, int rows_per_sample,
                                       long long sample_col_stride, float scale) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[(long long)r * ld + c] * scale : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) {
      __nv_bfloat16 sp[TERMS];
      umma::split_bf16<TERMS>(tile[threadIdx.x][i], sp);
      const long long col = (long long)(r / rows_per_sample) * sample_col_stride + (r % rows_per_sample);
#pragma unroll
      for (int t = 0; t < TERMS; ++t) dst[t * plane_stride + (long long)c * dst_ld + col] = sp[t];
    }
  }
}

inline int launch_transpose_split(int terms, const float* src, long long ld, __nv_bfloat16* dst, long long plane_stride,
                                  long long dst_ld, int rows, int cols, int rows_per_sample, long long sample_col_stride,
                                  float scale, cudaStream_t st) {
  dim3 grid(ceil_div(rows, 32), ceil_div(cols, 32)), block(32, 8);
  if (terms == 1) transpose_split_kernel<1><<<grid, block, 0, st>>>(src, ld, dst, plane_stride, dst_ld, rows, cols, rows_per_sample, sample_col_stride, scale);
  else if (terms == 2) transpose_split_kernel<2><<<grid, block, 0, st>>>(src, ld, dst, plane_stride, dst_ld, rows, cols, rows_per_sample, sample_col_stride, scale);
  else if (terms == 3) transpose_split_kernel<3><<<grid, block, 0, st>>>(src, ld, dst, plane_stride, dst_ld, rows, cols, rows_per_sample, sample_col_stride, scale);
  else A2P_FAIL("transpose_split: terms must be 1..3");
  A2P_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace a2p
