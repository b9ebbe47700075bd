// K1 core for SHORT key sets (<= 64 keys, no extra keys; head dim 32, split-bf16 x2): the keyframe cross-attention of the
// pose model (20 keys).  With a single key block per (row tile, head pair) the second-generation kernel
// (umma_attention2.cuh) spends its time in per-CTA fixed costs: barrier / tensor-memory set-up, the first TMA round trip and
// the output store, three rounds of them for 320 tiles on 148 SMs.  Here ONE CTA walks over ALL head pairs of its
// (sample-row, 128-query tile): the set-up is paid once, the Q / K / V^T tiles of the next head pairs are in flight while the
// current one is in the softmax, and a launch is a single partial round (R * ceil(T / 128) CTAs).
//
//   warp 0     TMA producer : per head pair g one stage = Q tile [128][64] + K tile [64 keys][64] + V^T tile [64][64 keys],
//                             two planes each (SWIZZLE_128B), 3-stage ring
//   warp 1     MMA issuer   : S_g = Q_g K_g^T per head (M128 x N64, K = 32, 3 plane products), one head pair AHEAD of
//                             O_g = P_g V_g (M128 x N32, K = 64 keys, A operand = P planes in tensor memory)
//   warp 2     TMEM alloc   : 512 columns, same map as umma_attention2.cuh (per head 2 x 64 S/P + 2 x 32 PV)
//   warps 4-11 softmax      : warpgroup w = head w of the pair; thread = query row.  One block per head: max, exp2, sum, P
//                             planes by tcgen05.st over the S buffer; the PV result of pair g is normalised and stored while
//                             pair g + 1 is already in the tensor pipe.
// Keys beyond n_keys inside the 64-key box (the next sample's keys or the TMA zero fill) are masked to -inf before the max.
// Operands / params as umma_attention.cuh; Q pre-scaled by log2(e)/sqrt(dh).
#pragma once
#include "umma_attention2.cuh"

namespace a2p {

struct AttnShortCfg {
  static constexpr int NST = 3;
  static constexpr int STAGE_BYTES = 2 * 16384 + 2 * 8192 + 2 * 8192;   // Q planes | K planes | V^T planes
  static constexpr int SMEM_BYTES = NST * STAGE_BYTES + 1024 + 512;
  static constexpr int THREADS = 384;
};

__global__ void __launch_bounds__(384, 1)
umma_attn_short_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                       const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV0,
                       const __grid_constant__ CUtensorMap tmV1, TcAttnParams p) {
  using Cfg = AttnShortCfg;
  constexpr int NST = Cfg::NST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sStage = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sStage + NST * Cfg::STAGE_BYTES);
  uint64_t* st_full = bars;            // [3]
  uint64_t* st_empty = bars + 3;       // [3]
  uint64_t* s_full = bars + 6;         // [head][2]
  uint64_t* p_ready = bars + 10;       // [head][2]  128 arrivals
  uint64_t* pv_full = bars + 14;       // [head][2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = ((int)blockIdx.x % p.n_qt) * 128, r = (int)blockIdx.x / p.n_qt;
  const int br = r >= p.rows_per_branch ? 1 : 0;
  const int rr = r - br * p.rows_per_branch;
  const int n_it = p.n_groups;         // head pairs, one iteration each

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmQ);
    umma::prefetch_tmap(br ? &tmK1 : &tmK0);
    umma::prefetch_tmap(br ? &tmV1 : &tmV0);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NST; ++i) { umma::mbar_init(&st_full[i], 1); umma::mbar_init(&st_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { umma::mbar_init(&s_full[i], 1); umma::mbar_init(&p_ready[i], 128); umma::mbar_init(&pv_full[i], 1); }
    umma::fence_barrier_init();
  }
  if (warp == 2) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  if (warp == 0) {
    // ================= TMA producer =================
    const CUtensorMap* tK = br ? &tmK1 : &tmK0;
    const CUtensorMap* tV = br ? &tmV1 : &tmV0;
    const int k_row_base = (int)(rr * p.k_row_stride[br]);
    const int v_col_base = (int)(rr * p.v_col_stride[br]);
    int st = 0; uint32_t ph = 0;
    for (int g = 0; g < n_it; ++g) {
      umma::mbar_wait(&st_empty[st], ph ^ 1);
      if (umma::elect_one()) {
        umma::mbar_expect_tx(&st_full[st], Cfg::STAGE_BYTES);
        uint8_t* sq = sStage + st * Cfg::STAGE_BYTES;
        uint8_t* sk = sq + 2 * 16384;
        uint8_t* sv = sk + 2 * 8192;
#pragma unroll
        for (int i = 0; i < 2; ++i) umma::tma_load_3d(&tmQ, &st_full[st], sq + i * 16384, p.q_col0 + g * 64, r * p.T + q0, i);
#pragma unroll
        for (int i = 0; i < 2; ++i) umma::tma_load_3d(tK, &st_full[st], sk + i * 8192, p.k_col0 + g * 64, k_row_base, i);
#pragma unroll
        for (int i = 0; i < 2; ++i) umma::tma_load_3d(tV, &st_full[st], sv + i * 8192, v_col_base, g * 64, i);
      }
      __syncwarp();
      if (++st == NST) { st = 0; ph ^= 1; }
    }
    pdl_trigger();    // all inbound tiles requested: the stream successor's prologue may overlap this CTA's tail (no-op without A2P_PDL)
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idS = umma::idesc_bf16_f32(128, 64);
    constexpr uint32_t idPV = umma::idesc_bf16_f32(128, 32);
    const uint32_t loStage = umma::desc_lo(umma::smem_u32(sStage));
    int st = 0; uint32_t ph = 0;        // stage / phase of pair i (S side)
    int stj = 0;                        // stage of pair i - 1 (PV side)
    for (int i = 0; i <= n_it; ++i) {
      if (i < n_it) {
        umma::mbar_wait(&st_full[st], ph);
        umma::fence_after();
        if (umma::elect_one()) {
          const uint32_t loq0 = loStage + st * (Cfg::STAGE_BYTES >> 4);
          const uint32_t lok0 = loq0 + (2 * 16384 >> 4);
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            const uint32_t loq = loq0 + w * 4, lok = lok0 + w * 4;     // head w: columns [32w, 32w+32) = +64 B
            const uint32_t d = tmem_base + w * 128 + (i & 1) * 64;
#pragma unroll
            for (int pr = 0; pr < 3; ++pr)
#pragma unroll
              for (int k = 0; k < 2; ++k)
                umma::mma_bf16(d, umma::desc_make(loq + prod_a(pr) * (16384 >> 4) + 2 * k),
                               umma::desc_make(lok + prod_b(pr) * (8192 >> 4) + 2 * k), idS, (pr | k) != 0 ? 1u : 0u);
            umma::mma_commit(&s_full[w * 2 + (i & 1)]);
          }
        }
        __syncwarp();
      }
      if (i > 0) {
        const int j = i - 1, b = j & 1;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          umma::mbar_wait(&p_ready[w * 2 + b], (j >> 1) & 1);
          umma::fence_after();
          if (umma::elect_one()) {
            const uint32_t lov = loStage + stj * (Cfg::STAGE_BYTES >> 4) + ((2 * 16384 + 2 * 8192) >> 4) + w * (32 * 128 >> 4);   // V^T rows [32w, 32w+32)
            const uint32_t d = tmem_base + 256 + w * 64 + b * 32;
            const uint32_t tp = tmem_base + w * 128 + b * 64;            // P planes: +0 (hi), +32 (lo); 8 columns per 16 keys
#pragma unroll
            for (int pr = 0; pr < 3; ++pr)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                mma_bf16_ts(d, tp + prod_a(pr) * 32 + 8 * k, umma::desc_make(lov + prod_b(pr) * (8192 >> 4) + 2 * k), idPV,
                            (pr | k) != 0 ? 1u : 0u);
            umma::mma_commit(&pv_full[w * 2 + b]);
            if (w == 1) umma::mma_commit(&st_empty[stj]);
          }
          __syncwarp();
        }
        if (++stj == NST) stj = 0;
      }
      if (i < n_it) { if (++st == NST) { st = 0; ph ^= 1; } }
    }
  } else if (warp >= 4) {
    // ================= softmax / output: warpgroup w owns head w of every pair =================
    const int w = (warp - 4) >> 2;
    const int wq = warp & 3;
    const int trow = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const uint32_t tmS = tmem_base + lane_addr + w * 128;
    const uint32_t tmO = tmem_base + lane_addr + 256 + w * 64;
    const int row = q0 + trow;
    const long long grow = (long long)r * p.T + row;
    const int nvalid = p.n_keys;        // <= 64
    float l_prev = 1.f;
    // PV result of pair j -> normalise -> store head w of this row
    auto finish = [&](int j, float lsum) {
      const int b = j & 1;
      umma::mbar_wait(&pv_full[w * 2 + b], (j >> 1) & 1);
      umma::fence_after();
      float o[32];
      umma::tmem_ld32(tmO + b * 32, o);
      umma::tmem_ld_wait();
      if (row < p.T) {
        const float inv = 1.f / lsum;
#pragma unroll
        for (int c = 0; c < 32; ++c) o[c] *= inv;
        const int col = j * 64 + w * 32;
        if (p.O) {
          float* dst = p.O + grow * p.o_ld + col;
#pragma unroll
          for (int c = 0; c < 32; c += 4) *reinterpret_cast<float4*>(dst + c) = make_float4(o[c], o[c + 1], o[c + 2], o[c + 3]);
        }
        if (p.Op) {
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            uint32_t pk[2][4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              uint32_t sp[2];
              umma::split_bf16_pair<2>(o[c + 2 * e], o[c + 2 * e + 1], sp);
              pk[0][e] = sp[0]; pk[1][e] = sp[1];
            }
#pragma unroll
            for (int t = 0; t < 2; ++t)
              *reinterpret_cast<uint4*>(p.Op + t * p.op_plane_stride + grow * p.o_ld + col + c) = make_uint4(pk[t][0], pk[t][1], pk[t][2], pk[t][3]);
          }
        }
      }
    };
#pragma unroll 1
    for (int i = 0; i < n_it; ++i) {
      const int b = i & 1;
      umma::mbar_wait(&s_full[w * 2 + b], (i >> 1) & 1);
      umma::fence_after();
      float s[64];
      umma::tmem_ld32(tmS + b * 64, s);
      umma::tmem_ld32(tmS + b * 64 + 32, s + 32);
      umma::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 64; ++c) s[c] = c < nvalid ? s[c] : -INFINITY;
      float mx0 = fmax3(s[0], s[1], s[2]), mx1 = fmax3(s[3], s[4], s[5]);
#pragma unroll
      for (int c = 6; c < 62; c += 4) { mx0 = fmax3(mx0, s[c], s[c + 1]); mx1 = fmax3(mx1, s[c + 2], s[c + 3]); }
      const float nm = -fmax3(fmaxf(mx0, mx1), s[62], s[63]);
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          float x0, x1;
          fadd2(x0, x1, s[hf * 32 + 2 * e], s[hf * 32 + 2 * e + 1], nm, nm);
          const float pa = umma::ex2_approx(x0), pb = umma::ex2_approx(x1);
          fadd2(rs0, rs1, rs0, rs1, pa, pb);
          split_prob_pair2(pa, pb, hi[e], lo[e]);
        }
        tmem_st16(tmS + b * 64 + hf * 16, hi);
        tmem_st16(tmS + b * 64 + 32 + hf * 16, lo);
      }
      tmem_st_wait2();
      umma::fence_before();
      umma::mbar_arrive(&p_ready[w * 2 + b]);
      // the previous pair's output leaves while this pair's PV product runs
      if (i > 0) finish(i - 1, l_prev);
      l_prev = rs0 + rs1;
    }
    finish(n_it - 1, l_prev);
  }
  __syncthreads();
  if (warp == 2) {
    umma::fence_after();
    umma::tmem_dealloc<512>(tmem_base);
  }
}

inline bool attn_short_disabled() {
  static int v = -1;
  if (v < 0) v = getenv("A2P_NO_ATTN_SHORT") ? 1 : 0;
  return v == 1;
}
inline bool attn_short_ok(const TcAttnParams& p) { return p.dh == 32 && p.n_keys >= 1 && p.n_keys <= 64 && p.n_extra == 0 && p.D % 64 == 0; }

inline int launch_umma_attn_short(const TcAttnOperands& o, const TcAttnParams& p, cudaStream_t st) {
  using Cfg = AttnShortCfg;
  if (!attn_short_ok(p)) A2P_FAIL("umma_attn_short: needs head dim 32, 1..64 keys, no extra keys");
  CUtensorMap tq, tk[2], tv[2];
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  A2P_TRY(make_tmap_bf16_3d(&tq, o.Q, o.q_ld, o.q_rows, 2, o.q_ld, o.q_plane_stride, 64, 128, sw));
  for (int b = 0; b < 2; ++b) {
    const int s = o.K[b] ? b : 0;
    A2P_TRY(make_tmap_bf16_3d(&tk[b], o.K[s], o.k_ld[s], o.k_rows[s], 2, o.k_ld[s], o.k_plane_stride[s], 64, 64, sw));
    A2P_TRY(make_tmap_bf16_3d(&tv[b], o.Vt[s], o.vt_cols[s], o.vt_rows, 2, o.vt_ld[s], o.vt_plane_stride[s], 64, 64, sw));
  }
  TcAttnParams q = p;
  q.n_qt = ceil_div(p.T, 128); q.n_groups = p.D / 64;
  dim3 grid(q.n_qt * p.R);
  A2P_CUDA(launch_pdl(umma_attn_short_kernel, grid, dim3(Cfg::THREADS), (size_t)Cfg::SMEM_BYTES, st, tq, tk[0], tk[1], tv[0], tv[1], q));
  return 0;
}

inline int init_umma_attn_short() {
  A2P_CUDA(cudaFuncSetAttribute(umma_attn_short_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnShortCfg::SMEM_BYTES));
  return 0;
}

}  // namespace a2p
