// K1 core, third generation (head dim 32, split-bf16 x2): as umma_attention2.cuh (one CTA = one 128-query tile of one
// sample row and one PAIR of heads sharing the K / V^T tiles; S = Q K^T on tcgen05 one key block ahead of O_blk = P V with
// the P planes as the TENSOR-MEMORY A operand), but with SIXTEEN softmax warps instead of eight:
//
//   warpgroup (w, kh), w = head, kh = key half: thread = query row, 32 of the 64 keys of every block.  Each warpgroup
//   runs its own online softmax (running max / sum / O in registers) over its key subset and its own PV accumulator;
//   the two partial results of a head are merged once at the end (flash-decoding style).
//
// Why: the 8-warp kernel is latency-bound, not pipe-bound (ncu profiles/r01j: issue slots 52 % busy, MUFU 42 %, tensor
// 31 %; 2 warps per scheduler cannot cover the MUFU / tcgen05.ld / mbarrier latencies of a 64-score dependent chain).
// Four warps per scheduler do.  To fit 640 threads x 96 registers the O accumulator lives in TENSOR MEMORY (the PV products
// accumulate there) and is rescaled lazily: probabilities are taken against a stale running maximum that is only advanced
// when the true maximum has grown by more than 2^8 (values stay exact to 16 bits in the bf16 planes; the final O / l is
// unchanged), so the read-modify-write of O is rare instead of once per block.
//
//   warp 0     TMA producer | warp 1 MMA issuer | warp 2 TMEM allocator | warps 4-19 softmax
//   TMEM: head w: S / P buffers at w*128 + b*64 (key half kh uses columns [32 kh, 32 kh + 32): 16 hi + 16 lo after the
//   in-place split); O accumulators at 256 + (w*2 + kh)*32.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include <type_traits>

#include "common.cuh"
#include "umma.cuh"
#include "umma_attention.cuh"
#include "umma_attention2.cuh"

namespace a2p {

struct Attn3Cfg {
  static constexpr int NST = 3;
  static constexpr int Q_BYTES = 2 * 16384;
  static constexpr int KV_STAGE_BYTES = 2 * 2 * 8192;
  static constexpr int XCH_FLOATS = 34;                       // per row: m, l, o[32]
  static constexpr int XCH_BYTES = 2 * 128 * XCH_FLOATS * 4;  // [head][row]
  static constexpr int SMEM_BYTES = Q_BYTES + NST * KV_STAGE_BYTES + XCH_BYTES + 1024 + 512;
  static constexpr int THREADS = 640;
};

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ void tmem_st32f(uint32_t taddr, const float* v) {
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

template <int POLY>
__global__ void __launch_bounds__(640, 1)
umma_attn3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                  const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV0,
                  const __grid_constant__ CUtensorMap tmV1, const __grid_constant__ CUtensorMap tmKx,
                  const __grid_constant__ CUtensorMap tmVx, TcAttnParams p) {
  using Cfg = Attn3Cfg;
  constexpr int NST = Cfg::NST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sKV = sQ + Cfg::Q_BYTES;
  float* sX = reinterpret_cast<float*>(sKV + NST * Cfg::KV_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sX) + Cfg::XCH_BYTES);
  uint64_t* q_full = bars;             // [1]
  uint64_t* kv_full = bars + 1;        // [3]
  uint64_t* kv_empty = bars + 4;       // [3]
  uint64_t* s_full = bars + 7;         // [head][2]
  uint64_t* p_ready = bars + 11;       // [head][kh][2]  128 arrivals
  uint64_t* pv_full = bars + 19;       // [head][kh]: PV(j) of that warpgroup has completed (phase j)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 27);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, g = blockIdx.y, r = blockIdx.z;
  const int br = r >= p.rows_per_branch ? 1 : 0;
  const int rr = r - br * p.rows_per_branch;
  const int nb_main = ceil_div(p.n_keys, 64);
  const int n_blocks = nb_main + (p.n_extra > 0 ? 1 : 0);

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmQ);
    umma::prefetch_tmap(br ? &tmK1 : &tmK0);
    umma::prefetch_tmap(br ? &tmV1 : &tmV0);
  }
  if (warp == 1 && lane == 0) {
    umma::mbar_init(q_full, 1);
    for (int i = 0; i < NST; ++i) { umma::mbar_init(&kv_full[i], 1); umma::mbar_init(&kv_empty[i], 1); }
    for (int i = 0; i < 4; ++i) umma::mbar_init(&s_full[i], 1);
    for (int i = 0; i < 8; ++i) umma::mbar_init(&p_ready[i], 128);
    for (int i = 0; i < 4; ++i) umma::mbar_init(&pv_full[i], 1);
    umma::fence_barrier_init();
  }
  if (warp == 2) umma::tmem_alloc<512>(tmem_slot);
  pdl_trigger();
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (umma::elect_one()) {
      umma::mbar_expect_tx(q_full, Cfg::Q_BYTES);
#pragma unroll
      for (int i = 0; i < 2; ++i) umma::tma_load_3d(&tmQ, q_full, sQ + i * 16384, p.q_col0 + g * 64, r * p.T + q0, i);
    }
    __syncwarp();
    const CUtensorMap* tK = br ? &tmK1 : &tmK0;
    const CUtensorMap* tV = br ? &tmV1 : &tmV0;
    const int k_row_base = (int)(rr * p.k_row_stride[br]);
    const int v_col_base = (int)(rr * p.v_col_stride[br]);
    int st = 0; uint32_t ph = 0;
#pragma unroll 1
    for (int j = 0; j < n_blocks; ++j) {
      umma::mbar_wait_nc(&kv_empty[st], ph ^ 1);
      if (umma::elect_one()) {
        umma::mbar_expect_tx(&kv_full[st], Cfg::KV_STAGE_BYTES);
        uint8_t* sk = sKV + st * Cfg::KV_STAGE_BYTES;
        uint8_t* sv = sk + 2 * 8192;
        const bool mainb = j < nb_main;
        const CUtensorMap* mk = mainb ? tK : &tmKx;
        const CUtensorMap* mv = mainb ? tV : &tmVx;
        const int kc0 = mainb ? p.k_col0 + g * 64 : p.kx_col0 + g * 64, kc1 = mainb ? k_row_base + j * 64 : r * p.kx_row_stride;
        const int vc0 = mainb ? v_col_base + j * 64 : r * p.vx_col_stride, vc1 = mainb ? g * 64 : p.vx_row0 + g * 64;
#pragma unroll
        for (int i = 0; i < 2; ++i) umma::tma_load_3d(mk, &kv_full[st], sk + i * 8192, kc0, kc1, i);
#pragma unroll
        for (int i = 0; i < 2; ++i) umma::tma_load_3d(mv, &kv_full[st], sv + i * 8192, vc0, vc1, i);
      }
      __syncwarp();
      if (++st == NST) { st = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idS = umma::idesc_bf16_f32(128, 64);
    constexpr uint32_t idPV = umma::idesc_bf16_f32(128, 32);
    const uint32_t loQ = umma::desc_lo(umma::smem_u32(sQ));
    const uint32_t loKV = umma::desc_lo(umma::smem_u32(sKV));
    umma::mbar_wait_nc(q_full, 0);
    int st = 0; uint32_t ph = 0;
    int stj = 0;
#pragma unroll 1
    for (int i = 0; i <= n_blocks; ++i) {
      if (i < n_blocks) {
        umma::mbar_wait_nc(&kv_full[st], ph);
        umma::fence_after();
        if (umma::elect_one()) {
#pragma unroll 1
          for (int w = 0; w < 2; ++w) {
            const uint32_t lok = loKV + st * (Cfg::KV_STAGE_BYTES >> 4) + w * 4;
            const uint32_t loq = loQ + w * 4;
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
#pragma unroll 1
        for (int u = 0; u < 4; ++u) {            // (head w, key half kh)
          const int w = u >> 1, kh = u & 1;
          umma::mbar_wait_nc(&p_ready[u * 2 + b], (j >> 1) & 1);
          umma::fence_after();
          if (umma::elect_one()) {
            // V^T rows [32w, 32w+32) (channels), K offset = keys kh*32 + 16k
            const uint32_t lov = loKV + stj * (Cfg::KV_STAGE_BYTES >> 4) + 2 * (8192 >> 4) + w * (32 * 128 >> 4) + kh * 4;
            const uint32_t d = tmem_base + 256 + u * 32;                 // O accumulates over all key blocks
            const uint32_t tp = tmem_base + w * 128 + b * 64 + kh * 32;   // hi at +0, lo at +16; 8 columns per 16 keys
#pragma unroll
            for (int pr = 0; pr < 3; ++pr)
#pragma unroll
              for (int k = 0; k < 2; ++k)
                mma_bf16_ts(d, tp + prod_a(pr) * 16 + 8 * k, umma::desc_make(lov + prod_b(pr) * (8192 >> 4) + 2 * k), idPV,
                            (j | pr | k) != 0 ? 1u : 0u);
            umma::mma_commit(&pv_full[u]);
            if (u == 3) umma::mma_commit(&kv_empty[stj]);
          }
          __syncwarp();
        }
        if (++stj == NST) stj = 0;
      }
      if (i < n_blocks) { if (++st == NST) { st = 0; ph ^= 1; } }
    }
  } else if (warp >= 4) {
    // ================= softmax / output =================
    const int u = (warp - 4) >> 2;             // warpgroup: head w = u >> 1, key half kh = u & 1
    const int w = u >> 1, kh = u & 1;
    const int wq = warp & 3;
    const int trow = wq * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    const uint32_t tmS = tmem_base + lane_addr + w * 128 + kh * 32;
    const uint32_t tmO = tmem_base + lane_addr + 256 + u * 32;
    float m = -INFINITY, l = 0.f;           // m: the (stale) maximum the probabilities are taken against
    auto block = [&](int i, auto masked_tag) {
      constexpr bool MASKED = decltype(masked_tag)::value;
      const int b = i & 1;
      umma::mbar_wait_nc(&s_full[w * 2 + b], (i >> 1) & 1);
      umma::fence_after();
      float s[32];
      umma::tmem_ld32(tmS + b * 64, s);
      umma::tmem_ld_wait();
      if (MASKED) {
        const int nv_blk = (i < nb_main) ? ::min(64, p.n_keys - i * 64) : p.n_extra;
        const int nvalid = nv_blk - 32 * kh;                   // valid keys among this warpgroup's 32 (may be <= 0)
#pragma unroll
        for (int c = 0; c < 32; ++c) s[c] = c < nvalid ? s[c] : -INFINITY;
      }
      float mx0 = fmax3(s[0], s[1], s[2]), mx1 = fmax3(s[3], s[4], s[5]);
#pragma unroll
      for (int c = 6; c < 30; c += 4) { mx0 = fmax3(mx0, s[c], s[c + 1]); mx1 = fmax3(mx1, s[c + 2], s[c + 3]); }
      const float mx = fmax3(fmaxf(mx0, mx1), s[30], s[31]);
      // advance the reference maximum only when it is stale by more than 2^8 (or not set yet)
      const bool need = mx > m + 8.f;                          // false for mx = -inf; true for m = -inf and finite mx
      if (i == 0) {
        if (need) m = mx;                                      // O is still empty: nothing to rescale
      } else if (__any_sync(0xffffffffu, need)) {
        umma::mbar_wait_nc(&pv_full[u], (i - 1) & 1);          // every PV product issued so far has landed in O
        umma::fence_after();
        float ov[32];
        umma::tmem_ld32(tmO, ov);
        umma::tmem_ld_wait();
        const float alpha = need ? umma::ex2_approx(m - mx) : 1.f;   // m = -inf: O and l are zero, alpha = 0 is fine
#pragma unroll
        for (int c = 0; c < 32; ++c) ov[c] *= alpha;
        l *= alpha;
        if (need) m = mx;
        tmem_st32f(tmO, ov);
      }
      const float nm = (MASKED && m == -INFINITY) ? 0.f : -m;  // a warpgroup may own no valid key at all
      float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {          // 16 keys at a time: 8 + 8 packed registers live
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float x0, x1;
          fadd2(x0, x1, s[hf * 16 + 2 * e], s[hf * 16 + 2 * e + 1], nm, nm);
          const float a = ((2 * e) & 3) < POLY ? umma::ex2_poly(x0) : umma::ex2_approx(x0);
          const float bb = ((2 * e + 1) & 3) < POLY ? umma::ex2_poly(x1) : umma::ex2_approx(x1);
          fadd2(rs0, rs1, rs0, rs1, a, bb);
          split_prob_pair2(a, bb, hi[e], lo[e]);
        }
        tmem_st8(tmS + b * 64 + hf * 8, hi);
        tmem_st8(tmS + b * 64 + 16 + hf * 8, lo);
      }
      tmem_st_wait2();
      umma::fence_before();
      umma::mbar_arrive(&p_ready[u * 2 + b]);
      l += rs0 + rs1;
    };
    const int n_full = ::min(p.n_keys / 64, n_blocks);
    int i = 0;
#pragma unroll 1
    for (; i < n_full; ++i) block(i, std::false_type{});
#pragma unroll 1
    for (; i < n_blocks; ++i) block(i, std::true_type{});
    umma::mbar_wait_nc(&pv_full[u], (n_blocks - 1) & 1);
    umma::fence_after();
    float o[32];
    umma::tmem_ld32(tmO, o);
    umma::tmem_ld_wait();
    // ---- merge the two key-half partials of head w (kh = 1 -> smem -> kh = 0), normalise, store
    float* xr = sX + (w * 128 + trow) * Cfg::XCH_FLOATS;
    if (kh == 1) {
      xr[0] = m; xr[1] = l;
#pragma unroll
      for (int c = 0; c < 32; ++c) xr[2 + c] = o[c];
    }
    asm volatile("bar.sync %0, 256;" ::"r"(1 + w) : "memory");   // the two warpgroups of head w
    const int row = q0 + trow;
    if (kh == 0 && row < p.T) {
      const float m1 = xr[0], l1 = xr[1];
      const float mm = fmaxf(m, m1);
      const float w0 = umma::ex2_approx(m - mm), w1 = umma::ex2_approx(m1 - mm);   // exp2(-inf) = 0 for an empty key subset
      const float inv = 1.f / (l * w0 + l1 * w1);
      const float f0 = w0 * inv, f1 = w1 * inv;
#pragma unroll
      for (int c = 0; c < 32; ++c) o[c] = o[c] * f0 + xr[2 + c] * f1;
      const long long grow = (long long)r * p.T + row;
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
  }
  __syncthreads();
  if (warp == 2) {
    umma::fence_after();
    umma::tmem_dealloc<512>(tmem_base);
  }
}

template <int POLY>
int launch_umma_attn3_t(const TcAttnOperands& o, const TcAttnParams& p, cudaStream_t st) {
  using Cfg = Attn3Cfg;
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
  dim3 grid(ceil_div(p.T, 128), p.D / 64, p.R);
  A2P_CUDA(launch_pdl(umma_attn3_kernel<POLY>, grid, dim3(Cfg::THREADS), (size_t)Cfg::SMEM_BYTES, st, tq, tk[0], tk[1], tv[0], tv[1],
                      tkx, tvx, p));
  return 0;
}

// variant 5: all exponentials on MUFU; 6: 1 of every 4 on the FMA pipe
inline int launch_umma_attn3(int variant, const TcAttnOperands& o, const TcAttnParams& p, cudaStream_t st) {
  if (p.dh != 32) A2P_FAIL("umma_attn3: head dim must be 32");
  return variant == 6 ? launch_umma_attn3_t<1>(o, p, st) : launch_umma_attn3_t<0>(o, p, st);
}

inline int init_umma_attn3() {
  A2P_CUDA(cudaFuncSetAttribute(umma_attn3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Cfg::SMEM_BYTES));
  A2P_CUDA(cudaFuncSetAttribute(umma_attn3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Attn3Cfg::SMEM_BYTES));
  return 0;
}

}  // namespace a2p
