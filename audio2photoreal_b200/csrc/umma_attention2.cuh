// K1 core, second generation (head dim 32, split-bf16 x2): softmax(Q K^T) V for one (sample-row, 64-column head
// pair, 128-query tile) per CTA.  The two heads of the pair share the K / V^T tiles in shared memory and are
// processed CONCURRENTLY by two softmax warpgroups (thread = query row, 64 keys per iteration), so the fixed
// per-iteration latencies (mbarrier wake-up, tcgen05.ld, fences) are amortised over 64 scores per thread instead of
// 32 and no cross-warpgroup merge is needed.
//
//   warp 0     TMA producer : Q tile once; per 64-key block the K tile [64 keys][64 cols] and the V^T tile
//                             [64 cols][64 keys] of both planes (SWIZZLE_128B), 3-stage ring
//   warp 1     MMA issuer   : per block and head  S = Q K^T (M128 x N64, K = 32, 3 plane products) into one of two
//                             S/P TMEM buffers, one block AHEAD of  O_blk = P V (M128 x N32, K = 64 keys)
//   warp 2     TMEM alloc   : 512 columns: per head 2 x 64 (S / P) + 2 x 32 (PV)
//   warps 4-11 softmax      : warpgroup w = head w.  tcgen05.ld S -> running max / exp2 / sum -> P split into two
//                             bf16 planes.  PT = 1: the planes are written back with tcgen05.st over the S buffer and
//                             the PV product takes its A operand from TENSOR MEMORY (full-rate MMA, no shared-memory
//                             store / proxy fence);  PT = 0: planes go to shared memory (UMMA K-major SWIZZLE_128B).
//                             O accumulates in registers: O = O * alpha + PV.
// Synchronisation is by data flow only: S(i+2) and PV(i+2) are issued after p_ready(i+1), which the softmax warps
// signal after they have consumed S(i) / PV(i), so no "buffer free" barriers are needed for TMEM.
// Q must be pre-scaled by log2(e)/sqrt(dh) (done by the Q-projection epilogue).  Operands/params as umma_attention.cuh.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"
#include "umma.cuh"
#include "umma_attention.cuh"

namespace a2p {

template <int PT>
struct Attn2Cfg {
  static constexpr int NST = 3;
  static constexpr int Q_BYTES = 2 * 16384;            // [2 planes][128 rows][64 cols] bf16
  static constexpr int KV_STAGE_BYTES = 2 * 2 * 8192;  // K planes then V^T planes, 8 KB each
  static constexpr int P_BYTES = PT ? 0 : 2 * 2 * 16384;   // [head][plane][128 rows][64 keys] bf16
  static constexpr int SMEM_BYTES = 2 * Q_BYTES + NST * KV_STAGE_BYTES + P_BYTES + 1024 + 512;   // two Q buffers (persistent CTAs)
  static constexpr int THREADS = 384;
};

__device__ __forceinline__ void tmem_ld32b(uint32_t taddr, float* v) { umma::tmem_ld32(taddr, v); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait2() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// probabilities p in [0, 2^k]: plane 0 = upper 16 bits (truncation, exact residual), plane 1 = residual rounded half-up
// with an integer add (ALU pipe; the XU pipe is busy with MUFU.EX2).  |p - (p0 + p1)| <= 2^-17 p, unbiased.
__device__ __forceinline__ void split_prob_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  hi = __byte_perm(ua, ub, 0x7632);
  const float ra = a - __uint_as_float(ua & 0xFFFF0000u), rb = b - __uint_as_float(ub & 0xFFFF0000u);
  lo = __byte_perm(__float_as_uint(ra) + 0x8000u, __float_as_uint(rb) + 0x8000u, 0x7632);
}

// packed fp32 pair add (FADD2 on sm_100): (x0, x1) = (a0, a1) + (b0, b1) -- one issue slot for two elements
__device__ __forceinline__ void fadd2(float& x0, float& x1, float a0, float a1, float b0, float b1) {
  asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
      "mov.b64 ra, {%2, %3};\n\t"
      "mov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rc, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rc;\n\t}"
      : "=f"(x0), "=f"(x1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

// probabilities p >= 0: plane 0 = upper 16 bits (truncation); plane 1 = the exact residual p - plane0, rounded half-up with
// an integer add.  The residual is one packed add against the NEGATED plane-0 values (sign bit OR-ed in: p >= 0).
__device__ __forceinline__ void split_prob_pair2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  hi = __byte_perm(ua, ub, 0x7632);
  float ra, rb;
  fadd2(ra, rb, a, b, __uint_as_float((ua & 0xFFFF0000u) | 0x80000000u), __uint_as_float((ub & 0xFFFF0000u) | 0x80000000u));
  lo = __byte_perm(__float_as_uint(ra) + 0x8000u, __float_as_uint(rb) + 0x8000u, 0x7632);
}

// LO-plane variants of the split (template parameter LO of the kernel):
//   1: plane 1 = the residual TRUNCATED to bf16 (no rounding adds: 1 instruction per score less).  The truncation loses on average
//      2^-17 of every probability, one-sided; the row sum uses the unsplit probabilities, so the output is scaled by (1 + 2^-17)
//      at normalisation to stay unbiased in expectation.
//   2: plane 1 = cvt.rn.bf16x2.f32 of the residual pair (exact round-to-nearest in ONE instruction per pair, on the XU pipe)
__device__ __forceinline__ void split_prob_pair_trunc(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  hi = __byte_perm(ua, ub, 0x7632);
  float ra, rb;
  fadd2(ra, rb, a, b, __uint_as_float((ua & 0xFFFF0000u) | 0x80000000u), __uint_as_float((ub & 0xFFFF0000u) | 0x80000000u));
  lo = __byte_perm(__float_as_uint(ra), __float_as_uint(rb), 0x7632);
}
__device__ __forceinline__ void split_prob_pair_cvt(float a, float b, uint32_t& hi, uint32_t& lo) {
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  hi = __byte_perm(ua, ub, 0x7632);
  float ra, rb;
  fadd2(ra, rb, a, b, __uint_as_float((ua & 0xFFFF0000u) | 0x80000000u), __uint_as_float((ub & 0xFFFF0000u) | 0x80000000u));
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));     // upper half <- rb, lower half <- ra
}

#ifndef A2P_ATTN2_TRACE
#define A2P_ATTN2_TRACE 0   // 1: clock64 timeline of CTA (0,0,0) into TcAttnParams::trace (scripts/gpu_attn_trace.py 21)
#endif
#if A2P_ATTN2_TRACE
#define A2_TRACE(slot, cond) do { if (p.trace && blockIdx.x == 0 && (cond) && i < 64) p.trace[i * 32 + (slot)] = clock64(); } while (0)
#define A2_TRACE_SM(slot) do { A2_TRACE(slot, threadIdx.x == 128); A2_TRACE(16 + (slot), threadIdx.x == 256); } while (0)
#else
#define A2_TRACE(slot, cond) do { } while (0)
#define A2_TRACE_SM(slot) do { } while (0)
#endif

// POLY of every 4 exponentials are evaluated with the FMA-pipe polynomial (umma::ex2_poly) instead of MUFU.EX2
template <int PT, int POLY, int LO = 0, int QT = 0>
__global__ void __launch_bounds__(384, 1)
umma_attn2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                  const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV0,
                  const __grid_constant__ CUtensorMap tmV1, const __grid_constant__ CUtensorMap tmKx,
                  const __grid_constant__ CUtensorMap tmVx, TcAttnParams p) {
  using Cfg = Attn2Cfg<PT>;
  constexpr int NST = Cfg::NST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                  // [2 buffers][2 planes][128 rows][64 cols]
  uint8_t* sKV = sQ + 2 * Cfg::Q_BYTES;
  uint8_t* sP = sKV + NST * Cfg::KV_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + Cfg::P_BYTES);
  uint64_t* q_full = bars;             // [2] at bars[0], bars[26]  (see qf())
  uint64_t* kv_full = bars + 1;        // [3]
  uint64_t* kv_empty = bars + 4;       // [3]
  uint64_t* s_full = bars + 7;         // [head][2]
  uint64_t* p_ready = bars + 11;       // [head][2]  128 arrivals
  uint64_t* pv_full = bars + 15;       // [head][2]
  uint64_t* p_free = bars + 19;        // [head]     (PT = 0: the shared-memory P buffer has been read by PV)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 21);
  uint64_t* q_empty = bars + 28;       // [2] the S products of the work item that used Q buffer b have completed
  auto qf = [&](int b) -> uint64_t* { return b ? bars + 26 : q_full; };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Work items.  Item -> (tile, key range): tiles beyond split_full are cut into split_parts key ranges (a launch whose tile count
  // is not a multiple of the SM count would otherwise leave most SMs idle during its last round, and a small batch would not
  // fill the machine at all).  A CTA processes items blockIdx.x, blockIdx.x + gridDim.x, ...: with one CTA per item this is the
  // plain one-tile kernel; with a grid of (at most) one CTA per SM the CTAs are PERSISTENT -- tensor memory, barriers and the
  // tensor maps are set up once, the producer warp prefetches the next item's Q tile and first key blocks while the softmax
  // warps are still normalising / storing the current item, and the per-tile launch / drain gaps disappear.  All pipeline
  // indices (S / P / PV buffer, barrier parity, K/V ring stage) run on over the items: block i of an item has the global
  // index base + i, base = number of key blocks of this CTA's earlier items, identical in every role.
  const int n_items = p.split_full + (p.n_qt * p.n_groups * p.R - p.split_full) * p.split_parts;
  const int nb_main = ceil_div(p.n_keys, 64);
  const int n_blocks_all = nb_main + (p.n_extra > 0 ? 1 : 0);
  struct Item { int tile, part, nparts, q0, g, r, br, rr, kb0, n_blocks; };
  auto decode = [&](int item) -> Item {
    Item it;
    it.tile = item; it.part = 0; it.nparts = 1;
    if (item >= p.split_full) {
      const int t_ = item - p.split_full;
      it.tile = p.split_full + t_ / p.split_parts; it.part = t_ % p.split_parts; it.nparts = p.split_parts;
    }
    it.q0 = (it.tile % p.n_qt) * 128; it.g = (it.tile / p.n_qt) % p.n_groups; it.r = it.tile / (p.n_qt * p.n_groups);
    it.br = it.r >= p.rows_per_branch ? 1 : 0;
    it.rr = it.r - it.br * p.rows_per_branch;
    it.kb0 = it.part * n_blocks_all / it.nparts;
    it.n_blocks = (it.part + 1) * n_blocks_all / it.nparts - it.kb0;     // key blocks [kb0, kb0 + n_blocks): local i <-> global kb0 + i
    return it;
  };

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmQ);
    umma::prefetch_tmap(&tmK0); umma::prefetch_tmap(&tmV0);
    if (p.R > p.rows_per_branch) { umma::prefetch_tmap(&tmK1); umma::prefetch_tmap(&tmV1); }
  }
  if (warp == 1 && lane == 0) {
    // QT = 1: the softmax warps copy the Q planes into tensor memory (A operand of S from TMEM); they release the Q buffer
    for (int i = 0; i < 2; ++i) { umma::mbar_init(qf(i), 1); umma::mbar_init(&q_empty[i], QT ? 256 : 1); }
    umma::mbar_init(bars + 30, 256);   // q_tm: the item's Q planes are in tensor memory (QT)
    for (int i = 0; i < NST; ++i) { umma::mbar_init(&kv_full[i], 1); umma::mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { umma::mbar_init(&s_full[i], 1); umma::mbar_init(&p_ready[i], 128); umma::mbar_init(&pv_full[i], 1); }
    for (int i = 0; i < 2; ++i) umma::mbar_init(&p_free[i], 1);
    umma::fence_barrier_init();
  }
  if (warp == 2) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  // TMEM map: head w: S/P buffers at w*128 + b*64, PV buffers at 256 + w*64 + b*32
  if (warp == 0) {
    // ================= TMA producer =================
    int st = 0; uint32_t ph = 0;
    int itn = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++itn) {
      const Item it = decode(item);
      const int qb = itn & 1;
      umma::mbar_wait(&q_empty[qb], ((itn >> 1) & 1) ^ 1);      // the item that used this Q buffer two items ago is done with it
      if (umma::elect_one()) {
        umma::mbar_expect_tx(qf(qb), Cfg::Q_BYTES);
#pragma unroll
        for (int i = 0; i < 2; ++i)
          umma::tma_load_3d(&tmQ, qf(qb), sQ + qb * Cfg::Q_BYTES + i * 16384, p.q_col0 + it.g * 64, it.r * p.T + it.q0, i);
      }
      __syncwarp();
      const CUtensorMap* tK = it.br ? &tmK1 : &tmK0;
      const CUtensorMap* tV = it.br ? &tmV1 : &tmV0;
      const int k_row_base = (int)(it.rr * p.k_row_stride[it.br]);
      const int v_col_base = (int)(it.rr * p.v_col_stride[it.br]);
      for (int j = 0; j < it.n_blocks; ++j) {
        umma::mbar_wait(&kv_empty[st], ph ^ 1);
        if (umma::elect_one()) {
          umma::mbar_expect_tx(&kv_full[st], Cfg::KV_STAGE_BYTES);
          uint8_t* sk = sKV + st * Cfg::KV_STAGE_BYTES;
          uint8_t* sv = sk + 2 * 8192;
          const int gb = it.kb0 + j;
          if (gb < nb_main) {
#pragma unroll
            for (int i = 0; i < 2; ++i) umma::tma_load_3d(tK, &kv_full[st], sk + i * 8192, p.k_col0 + it.g * 64, k_row_base + gb * 64, i);
#pragma unroll
            for (int i = 0; i < 2; ++i) umma::tma_load_3d(tV, &kv_full[st], sv + i * 8192, v_col_base + gb * 64, it.g * 64, i);
          } else {
#pragma unroll
            for (int i = 0; i < 2; ++i) umma::tma_load_3d(&tmKx, &kv_full[st], sk + i * 8192, p.kx_col0 + it.g * 64, it.r * p.kx_row_stride, i);
#pragma unroll
            for (int i = 0; i < 2; ++i) umma::tma_load_3d(&tmVx, &kv_full[st], sv + i * 8192, it.r * p.vx_col_stride, p.vx_row0 + it.g * 64, i);
          }
        }
        __syncwarp();
        if (++st == NST) { st = 0; ph ^= 1; }
      }
    }
    pdl_trigger();    // every key block of this CTA's last item is requested: the successor's prologue may overlap the tail (no-op without A2P_PDL)
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idS = umma::idesc_bf16_f32(128, 64);
    constexpr uint32_t idPV = umma::idesc_bf16_f32(128, 32);
    const uint32_t loKV = umma::desc_lo(umma::smem_u32(sKV));
    const uint32_t loP = umma::desc_lo(umma::smem_u32(sP));
    int st = 0; uint32_t ph = 0;        // stage / phase of the next S block (runs on over the items)
    int stj = 0;                        // stage of the next PV block
    int base = 0, itn = 0;              // global index of this item's block 0; item counter
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++itn) {
      const int n_blocks = decode(item).n_blocks;
      const int qb = itn & 1;
      const uint32_t loQ = umma::desc_lo(umma::smem_u32(sQ + qb * Cfg::Q_BYTES));
      if (QT) { umma::mbar_wait(bars + 30, itn & 1); umma::fence_after(); }
      else umma::mbar_wait(qf(qb), (itn >> 1) & 1);
      for (int i = 0; i <= n_blocks; ++i) {
        const int gi = base + i;          // buffers / parities follow the global block index
        A2_TRACE(8, lane == 0);
        if (i < n_blocks) {
          umma::mbar_wait(&kv_full[st], ph);
          umma::fence_after();
          A2_TRACE(9, lane == 0);
          if (umma::elect_one()) {
#pragma unroll
            for (int w = 0; w < 2; ++w) {
              const uint32_t lok = loKV + st * (Cfg::KV_STAGE_BYTES >> 4) + w * 4;   // head w: columns [32w, 32w+32) = +64 B
              const uint32_t loq = loQ + w * 4;
              const uint32_t d = tmem_base + w * 128 + (gi & 1) * 64;
#pragma unroll
              for (int pr = 0; pr < 3; ++pr)
#pragma unroll
                for (int k = 0; k < 2; ++k)
                  if (QT) mma_bf16_ts(d, tmem_base + 384 + w * 32 + prod_a(pr) * 16 + 8 * k,
                                      umma::desc_make(lok + prod_b(pr) * (8192 >> 4) + 2 * k), idS, (pr | k) != 0 ? 1u : 0u);
                  else umma::mma_bf16(d, umma::desc_make(loq + prod_a(pr) * (16384 >> 4) + 2 * k),
                                      umma::desc_make(lok + prod_b(pr) * (8192 >> 4) + 2 * k), idS, (pr | k) != 0 ? 1u : 0u);
              umma::mma_commit(&s_full[w * 2 + (gi & 1)]);
            }
            if (!QT && i == n_blocks - 1) umma::mma_commit(&q_empty[qb]);      // last read of this item's Q tile
          }
          __syncwarp();
        }
        if (i > 0) {
          const int gj = gi - 1, b = gj & 1;
#pragma unroll
          for (int w = 0; w < 2; ++w) {
            A2_TRACE(10 + 2 * w, lane == 0);
            umma::mbar_wait(&p_ready[w * 2 + b], (gj >> 1) & 1);
            umma::fence_after();
            A2_TRACE(11 + 2 * w, lane == 0);
            if (umma::elect_one()) {
              const uint32_t lov = loKV + stj * (Cfg::KV_STAGE_BYTES >> 4) + 2 * (8192 >> 4) + w * (32 * 128 >> 4);   // V^T rows [32w, 32w+32)
              const uint32_t d = tmem_base + 256 + w * 64 + b * 32;
              const uint32_t tp = tmem_base + w * 128 + b * 64;                    // P planes: +0 (hi), +32 (lo); 8 columns per 16 keys
              const uint32_t lop = loP + w * (2 * 16384 >> 4);
#pragma unroll
              for (int pr = 0; pr < 3; ++pr)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const uint64_t bd = umma::desc_make(lov + prod_b(pr) * (8192 >> 4) + 2 * k);
                  if (PT) mma_bf16_ts(d, tp + prod_a(pr) * 32 + 8 * k, bd, idPV, (pr | k) != 0 ? 1u : 0u);
                  else umma::mma_bf16(d, umma::desc_make(lop + prod_a(pr) * (16384 >> 4) + 2 * k), bd, idPV, (pr | k) != 0 ? 1u : 0u);
                }
              umma::mma_commit(&pv_full[w * 2 + b]);
              if (!PT) umma::mma_commit(&p_free[w]);
              if (w == 1) umma::mma_commit(&kv_empty[stj]);
            }
            __syncwarp();
          }
          if (++stj == NST) stj = 0;
        }
        if (i < n_blocks) { if (++st == NST) { st = 0; ph ^= 1; } }
      }
      base += n_blocks;
    }
  } else if (warp >= 4) {
    // ================= softmax / output: warpgroup w owns head w =================
    const int w = (warp - 4) >> 2;
    const int wq = warp & 3;
    const int trow = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const uint32_t tmS = tmem_base + lane_addr + w * 128;
    const uint32_t tmO = tmem_base + lane_addr + 256 + w * 64;
    const uint32_t sP_u32 = umma::smem_u32(sP) + w * (2 * 16384) + trow * 128;
    float m = -INFINITY, l = 0.f, o[32];
    float alpha_pend = 1.f;
    int base = 0, kb0 = 0;               // global index of the current item's block 0; its first key block
    // The two warpgroups would run in lockstep (both S tiles arrive together) and hit the MUFU-bound exponential phase at
    // the same time; starting head 1 half an iteration late lets one warpgroup's exponentials overlap the other's
    // load / max / barrier phase on every scheduler.
    if (w == 1 && p.skew_ns > 0) __nanosleep(p.skew_ns);
    auto consume_pv = [&](int j, float alpha) {
      const int b = j & 1;
      umma::mbar_wait(&pv_full[w * 2 + b], (j >> 1) & 1);
      umma::fence_after();
      float v[32];
      umma::tmem_ld32(tmO + b * 32, v);
      umma::tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 32; ++c) o[c] = o[c] * alpha + v[c];
    };
    // one 64-key block; MASKED blocks (a ragged tail of the main keys, the extra keys) live in a second copy of the body so
    // that the hot loop carries no per-score select
    auto block = [&](int i, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const int gi = base + i;           // S / P / PV buffers and barrier parities follow the global block index
      const int b = gi & 1;
      A2_TRACE_SM(0);
      umma::mbar_wait(&s_full[w * 2 + b], (gi >> 1) & 1);
      umma::fence_after();
      A2_TRACE_SM(1);
      float s[64];
      umma::tmem_ld32(tmS + b * 64, s);
      umma::tmem_ld32(tmS + b * 64 + 32, s + 32);
      // PV(i-1) landed long ago: fold it into O now, so that its tcgen05.ld shares the wait with the S loads and its 32
      // FMAs fill the issue slots between the exponentials below instead of trailing the iteration
      if (i > 0) consume_pv(gi - 1, alpha_pend);
      else umma::tmem_ld_wait();
      A2_TRACE_SM(2);
      if (MASKED) {
        const int gb = kb0 + i;
        const int nvalid = (gb < nb_main) ? ::min(64, p.n_keys - gb * 64) : p.n_extra;   // warp-uniform
#pragma unroll
        for (int c = 0; c < 64; ++c) s[c] = c < nvalid ? s[c] : -INFINITY;
      }
      // four independent 3-input max chains (8 dependent instructions each instead of 15)
      float mx0 = fmax3(s[0], s[1], s[2]), mx1 = fmax3(s[3], s[4], s[5]), mx2 = fmax3(s[6], s[7], s[8]), mx3 = fmax3(s[9], s[10], s[11]);
#pragma unroll
      for (int c = 12; c < 60; c += 8) {
        mx0 = fmax3(mx0, s[c], s[c + 1]); mx1 = fmax3(mx1, s[c + 2], s[c + 3]);
        mx2 = fmax3(mx2, s[c + 4], s[c + 5]); mx3 = fmax3(mx3, s[c + 6], s[c + 7]);
      }
      const float mx = fmax3(fmax3(mx0, mx1, mx2), fmax3(mx3, s[60], s[61]), fmaxf(s[62], s[63]));
      const float mnew = fmaxf(m, mx);
      const float alpha = umma::ex2_approx(m - mnew);
      m = mnew;
      const float nm = -mnew;
      A2_TRACE_SM(3);
      if (!PT && gi > 0) umma::mbar_wait(&p_free[w], (gi - 1) & 1);   // PV(gi-1) has finished reading the shared-memory planes
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        uint32_t hi[16], lo[16];
        // software pipeline: the exponentials of pair e + DEPTH are issued before the split of pair e, so that the MUFU pipe
        // (one warp instruction per 8 cycles) is fed every ~6 issue slots instead of in one burst of 64
        constexpr int DEPTH = 3;
        float pa[16], pb[16];
#pragma unroll
        for (int e = 0; e < DEPTH; ++e) {
          float x0, x1;
          fadd2(x0, x1, s[hf * 32 + 2 * e], s[hf * 32 + 2 * e + 1], nm, nm);
          pa[e] = ((2 * e) & 3) < POLY ? umma::ex2_poly(x0) : umma::ex2_approx(x0);
          pb[e] = ((2 * e + 1) & 3) < POLY ? umma::ex2_poly(x1) : umma::ex2_approx(x1);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          if (e + DEPTH < 16) {
            const int f = e + DEPTH;
            float x0, x1;
            fadd2(x0, x1, s[hf * 32 + 2 * f], s[hf * 32 + 2 * f + 1], nm, nm);
            pa[f] = ((2 * f) & 3) < POLY ? umma::ex2_poly(x0) : umma::ex2_approx(x0);
            pb[f] = ((2 * f + 1) & 3) < POLY ? umma::ex2_poly(x1) : umma::ex2_approx(x1);
          }
          fadd2(rs0, rs1, rs0, rs1, pa[e], pb[e]);
          if (LO == 1) split_prob_pair_trunc(pa[e], pb[e], hi[e], lo[e]);
          else if (LO == 2) split_prob_pair_cvt(pa[e], pb[e], hi[e], lo[e]);
          else split_prob_pair2(pa[e], pb[e], hi[e], lo[e]);
        }
        if (PT) {
          tmem_st16(tmS + b * 64 + hf * 16, hi);
          tmem_st16(tmS + b * 64 + 32 + hf * 16, lo);
        } else {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            st_shared_v4(sP_u32 + (((hf * 4 + u) ^ (trow & 7)) << 4), hi[4 * u], hi[4 * u + 1], hi[4 * u + 2], hi[4 * u + 3]);
            st_shared_v4(sP_u32 + 16384 + (((hf * 4 + u) ^ (trow & 7)) << 4), lo[4 * u], lo[4 * u + 1], lo[4 * u + 2], lo[4 * u + 3]);
          }
        }
      }
      A2_TRACE_SM(4);
      if (PT) { tmem_st_wait2(); umma::fence_before(); }
      else umma::fence_proxy_async();
      umma::mbar_arrive(&p_ready[w * 2 + b]);
      A2_TRACE_SM(5);
      l = l * alpha + (rs0 + rs1);
      A2_TRACE_SM(6);
      alpha_pend = alpha;
    };
    int itn = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++itn) {
    const Item it = decode(item);
    const int n_blocks = it.n_blocks, tile = it.tile, part = it.part, nparts = it.nparts, q0 = it.q0, g = it.g, r = it.r;
    kb0 = it.kb0;
    if (QT) {
      // Q planes of head w, row trow: shared memory (SWIZZLE_128B rows of 128 B: 2 heads x 32 dims) -> tensor memory columns
      // [384 + 32 w + 16 plane, +16) in the A-operand layout (8 columns per 16 dims), so that S = Q K^T reads A from TMEM like PV
      // does: no 4 KB shared-memory fetch of Q per MMA (48 KB per block).  The previous item's S products are complete: this
      // warpgroup has consumed the last S tile of that item.
      const int qb = itn & 1;
      umma::mbar_wait(qf(qb), (itn >> 1) & 1);
      const uint32_t qrow = umma::smem_u32(sQ + qb * Cfg::Q_BYTES) + trow * 128;
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
        uint32_t qv[16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t a_ = qrow + pl * 16384 + (((w * 4 + u) ^ (trow & 7)) << 4);
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(qv[4 * u]), "=r"(qv[4 * u + 1]), "=r"(qv[4 * u + 2]), "=r"(qv[4 * u + 3]) : "r"(a_));
        }
        tmem_st16(tmem_base + lane_addr + 384 + w * 32 + pl * 16, qv);
      }
      tmem_st_wait2();
      umma::fence_before();
      umma::mbar_arrive(bars + 30);            // q_tm
      umma::mbar_arrive(&q_empty[qb]);         // the shared-memory Q buffer may be refilled
    }
    m = -INFINITY; l = 0.f; alpha_pend = 1.f;
#pragma unroll
    for (int c = 0; c < 32; ++c) o[c] = 0.f;
    const int n_full = ::max(0, ::min(p.n_keys / 64 - kb0, n_blocks));    // leading blocks whose 64 keys are all valid
    int i = 0;
#pragma unroll 1
    for (; i < n_full; ++i) block(i, std::false_type{});
#pragma unroll 1
    for (; i < n_blocks; ++i) block(i, std::true_type{});
    consume_pv(base + n_blocks - 1, alpha_pend);
    bool writer = true;
    if (nparts > 1) {
      // ---- split tile: publish this key range's (max, sum, O); the LAST work item of the tile to finish merges them all
      const int slot = tile - p.split_full;
      float* my = p.split_scratch + ((size_t)(slot * nparts + part) * 2 + w) * 34 * 128 + trow;
      my[0] = m; my[128] = l;
#pragma unroll
      for (int c = 0; c < 32; ++c) my[(2 + c) * 128] = o[c];
      __threadfence();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      int* flag = reinterpret_cast<int*>(bars + 24);
      if (threadIdx.x == 128) *flag = atomicAdd(p.split_counters + slot, 1);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      writer = (*flag == nparts - 1);
      if (writer) {
        __threadfence();
        // Fold the key ranges in INDEX order starting from range 0 (this CTA's own partial is re-read like the others):
        // the fp32 result must not depend on which work item happened to finish last (bit-identical reruns).
        {
          const float* o0 = p.split_scratch + ((size_t)(slot * nparts + 0) * 2 + w) * 34 * 128 + trow;
          m = __ldcg(o0); l = __ldcg(o0 + 128);
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = __ldcg(o0 + (2 + c) * 128);
        }
#pragma unroll 1
        for (int pp = 1; pp < nparts; ++pp) {
          const float* ot = p.split_scratch + ((size_t)(slot * nparts + pp) * 2 + w) * 34 * 128 + trow;
          const float m1 = __ldcg(ot), l1 = __ldcg(ot + 128);
          const float mm = fmaxf(m, m1);
          const float w0 = umma::ex2_approx(m - mm), w1 = umma::ex2_approx(m1 - mm);   // exp2(-inf) = 0 for an empty range
          l = l * w0 + l1 * w1;
#pragma unroll
          for (int c = 0; c < 32; ++c) o[c] = o[c] * w0 + __ldcg(ot + (2 + c) * 128) * w1;
          m = mm;
        }
        if (threadIdx.x == 128) p.split_counters[slot] = 0;   // ready for the next launch
      }
    }
    // ---- normalise, store head w of this row
    const int row = q0 + trow;
    if (writer && row < p.T) {
      const long long grow = (long long)r * p.T + row;
      const float inv = (LO == 1 ? 1.00000762939453125f : 1.f) / l;      // LO = 1: (1 + 2^-17), see split_prob_pair_trunc
#pragma unroll
      for (int c = 0; c < 32; ++c) o[c] *= inv;
      const int col = g * 64 + w * 32;
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
    base += n_blocks;
    }   // items
  }
  __syncthreads();
  if (warp == 2) {
    umma::fence_after();
    umma::tmem_dealloc<512>(tmem_base);
  }
}

inline int attn2_num_sms() {
  static int v = -1;
  if (v < 0) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); if (v <= 0) v = 148; }
  return v;
}
#ifndef A2P_ATTN_PERSIST_DEFAULT
#define A2P_ATTN_PERSIST_DEFAULT 1
#endif
inline int& attn2_persist_override() { static int v = -1; return v; }     // tests: 0 / 1 forces the mode of the next launches
inline bool attn2_persistent() {
  if (attn2_persist_override() >= 0) return attn2_persist_override() != 0;
  static int v = -1;
  if (v < 0) { const char* e = getenv("A2P_ATTN_PERSIST"); v = e ? (atoi(e) != 0) : A2P_ATTN_PERSIST_DEFAULT; }
  return v == 1;
}
inline bool attn2_split_disabled() {
  static int v = -1;
  if (v < 0) v = getenv("A2P_NO_SPLIT_KV") ? 1 : 0;
  return v == 1;
}
// scratch for the split tiles of one launch: at most num_sms work items x 2 heads x (max, sum, 32 accumulators) x 128 rows
inline size_t attn2_split_scratch_floats() { return (size_t)attn2_num_sms() * 2 * 34 * 128; }
inline size_t attn2_split_counter_ints() { return (size_t)attn2_num_sms(); }

template <int PT, int POLY, int LO = 0, int QT = 0>
int launch_umma_attn2_t(const TcAttnOperands& o, const TcAttnParams& p, cudaStream_t st) {
  using Cfg = Attn2Cfg<PT>;
  CUtensorMap tq, tk[2], tv[2], tkx, tvx;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  A2P_TRY(make_tmap_bf16_3d(&tq, o.Q, o.q_ld, o.q_rows, 2, o.q_ld, o.q_plane_stride, 64, 128, sw));
  for (int b = 0; b < 2; ++b) {
    const int s = o.K[b] ? b : 0;
    A2P_TRY(make_tmap_bf16_3d(&tk[b], o.K[s], o.k_ld[s], o.k_rows[s], 2, o.k_ld[s], o.k_plane_stride[s], 64, 64, sw));
    A2P_TRY(make_tmap_bf16_3d(&tv[b], o.Vt[s], o.vt_cols[s], o.vt_rows, 2, o.vt_ld[s], o.vt_plane_stride[s], 64, 64, sw));
  }
  if (o.Kx) {
    A2P_TRY(make_tmap_bf16_3d(&tkx, o.Kx, o.kx_ld, o.kx_rows, 2, o.kx_ld, o.kx_plane_stride, 64, 64, sw));
    A2P_TRY(make_tmap_bf16_3d(&tvx, o.Vx, o.vx_cols, o.vx_rows, 2, o.vx_ld, o.vx_plane_stride, 64, 64, sw));
  } else {
    tkx = tk[0]; tvx = tv[0];
  }
  TcAttnParams q = p;
  q.n_qt = ceil_div(p.T, 128); q.n_groups = p.D / 64;
  const int n_tiles = q.n_qt * q.n_groups * p.R;
  const int nb = ceil_div(p.n_keys, 64) + (p.n_extra > 0 ? 1 : 0);
  const int sms = attn2_num_sms();
  q.split_full = n_tiles; q.split_parts = 1;
  const int rem = n_tiles % sms;
  if (p.split_scratch && p.split_counters && rem > 0 && !attn2_split_disabled()) {
    int parts = sms / rem;                         // work items that fit beside each other in the last round
    if (parts > nb / 2) parts = nb / 2;            // at least two key blocks per range
    if (parts > 8) parts = 8;
    if (parts >= 2) { q.split_full = n_tiles - rem; q.split_parts = parts; }
  }
  const int n_items = q.split_full + (n_tiles - q.split_full) * q.split_parts;
  // persistent CTAs (A2P_ATTN_PERSIST=1): one CTA per SM walks items blockIdx.x, + gridDim.x, ...; default: one CTA per item
  dim3 grid(attn2_persistent() && n_items > sms ? sms : n_items);
  A2P_CUDA(launch_pdl(umma_attn2_kernel<PT, POLY, LO, QT>, grid, dim3(Cfg::THREADS), (size_t)Cfg::SMEM_BYTES, st, tq, tk[0], tk[1], tv[0], tv[1],
                      tkx, tvx, q));
  return 0;
}

// variant: 1 = P planes in shared memory; 2 (default) = P planes and Q planes in tensor memory; 8 = as 2 with Q in shared memory;
// 3 / 4 = as 8 with 1 / 2 of every 4 exponentials on the FMA pipe; 6 / 7 = as 8 with the lo-plane split variants
inline int launch_umma_attn2(int variant, const TcAttnOperands& o, const TcAttnParams& p, cudaStream_t st) {
  if (p.dh != 32) A2P_FAIL("umma_attn2: head dim must be 32");
  switch (variant) {
    case 1: return launch_umma_attn2_t<0, 0>(o, p, st);
    case 2: return launch_umma_attn2_t<1, 0, 0, 1>(o, p, st);  // default: P planes AND Q planes in tensor memory
    case 3: return launch_umma_attn2_t<1, 1>(o, p, st);
    case 4: return launch_umma_attn2_t<1, 2>(o, p, st);
    case 6: return launch_umma_attn2_t<1, 0, 1>(o, p, st);     // lo plane truncated (+ expectation compensation)
    case 7: return launch_umma_attn2_t<1, 0, 2>(o, p, st);     // lo plane by cvt.rn.bf16x2.f32
    case 8: return launch_umma_attn2_t<1, 0, 0, 0>(o, p, st);  // as 2 with the Q planes read from shared memory by every S product (A/B)
  }
  A2P_FAIL("umma_attn2: unknown variant %d", variant);
}

inline int init_umma_attn2() {
#define A2P_SET(PT_, PL_, LO_) A2P_CUDA(cudaFuncSetAttribute(umma_attn2_kernel<PT_, PL_, LO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn2Cfg<PT_>::SMEM_BYTES));
  A2P_SET(0, 0, 0) A2P_SET(1, 0, 0) A2P_SET(1, 1, 0) A2P_SET(1, 2, 0) A2P_SET(1, 0, 1) A2P_SET(1, 0, 2)
#undef A2P_SET
  A2P_CUDA(cudaFuncSetAttribute(umma_attn2_kernel<1, 0, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn2Cfg<1>::SMEM_BYTES));
  return 0;
}

}  // namespace a2p
