// Fused "row chain" kernel for the D = 256 (pose) denoiser on sm_100a, split-bf16 x2 operands (3 tensor-core
// products per MAC, fp32 accumulation in TMEM).  One CTA owns 128 consecutive rows of the residual stream and runs,
// without leaving the SM:
//
//   GEMM0    acc0[128,256]  = A0[128,K0] * W0[256,K0]^T          A0 planes streamed from global by TMA
//   E_A      x = x + (film_scale + 1) * (acc0 + b0) + film_shift (or x = acc0 + b0), written back as fp32;
//            h = LayerNorm(x) (two-pass, fp32), optional full-width RoPE, split into bf16 planes written straight
//            into shared memory in the UMMA K-major SWIZZLE_128B layout (the A operand of GEMM1)
//   GEMM1    acc1[128,N1]   = h[128,256] * W1[N1,256]^T          N1 in {104, 256, 512, 1024}, 128 columns at a time
//   E_B      bias, optional scale / exact GELU, split into planes -> global (Q|K planes, FFN hidden planes, ...)
//   V job    acc[256,128]   = W2[256,256] * h'[128,256]^T        (h' = un-rotated LayerNorm output): V^T planes for
//            the PV product of the attention kernel come out already transposed
//
// This replaces, per decoder layer, 4 LayerNorm(+RoPE) launches and 9 GEMM launches of the unfused arm by 4 launches
// (transformer_modules.py:190-217: out_proj+FiLM+residual -> norm -> rotate -> in_proj of the NEXT block), keeps the
// 128x256 activation tile on chip between the two GEMMs and reads every weight once per 128 rows.
//
// Roles (384 threads):  warp 0 TMA producer | warp 1 MMA issuer | warp 2 TMEM allocator | warps 4-11 two epilogue
// warpgroups (thread = TMEM lane = row).  In E_A the warpgroups split the 256 columns (row statistics are exchanged
// through shared memory); in E_B they alternate 128-column accumulator halves, so the epilogue of half i overlaps the
// MMAs of half i+1.  TMEM: columns [0,256) acc0 / x, [256,384) and [384,512) the two GEMM1 accumulators.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"
#include "umma_gemm.cuh"

namespace a2p {

struct ChainParams {
  int M, T;                  // rows (samples * T); RoPE position = row % T, FiLM sample = row / T
  int K0;                    // GEMM0 reduction length (any multiple of 8; TMA zero-fills the tail of the last 64-chunk)
  const float* bias0;        // [256]
  int film_mode;             // 1: x += (scale + 1) * (acc0 + b0) + shift    0: x = acc0 + b0
  const float* film; long long film_ld; int film_scale_off, film_shift_off;
  float* x;                  // [M][256] fp32 residual stream
  int ln_mode;               // 1: h = LayerNorm(x) * ln_w + ln_b            0: h = x
  const float* ln_w; const float* ln_b;
  int rope;                  // rotate h before GEMM1 (the V job always uses the un-rotated h)
  const float2* rope_tab;    // [max_pos][128] (cos, sin)
  int N1;                    // GEMM1 output columns
  const float* bias1; float out_scale; int scale_ncols;   // out_scale applies to columns < scale_ncols (0 = all)
  int gelu;
  __nv_bfloat16* Cp; long long cp_plane_stride, ldcp; int remap_rps, remap_pad;
  int vjob; const float* bias2; __nv_bfloat16* Vt; long long vt_plane_stride, ldvt;
};

constexpr int CH_THREADS = 384;
constexpr int CH_RING = 4;
constexpr int CH_TILE = 16384;                    // one [128 rows][64 k] bf16 tile
constexpr int CH_A_BYTES = 8 * CH_TILE;           // [2 planes][4 k-chunks]
constexpr int CH_STG_BYTES = 8 * 4096;            // per epilogue warp: [32][32] fp32, 16-byte units XOR-swizzled by (row & 7)
constexpr int CH_RED_BYTES = 1024;                // [2 warpgroups][128 rows] fp32
constexpr int CH_SMEM_BYTES = CH_A_BYTES + CH_RING * CH_TILE + CH_STG_BYTES + CH_RED_BYTES + 256 + 1024;

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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void st_shared_v4u(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__global__ void __launch_bounds__(CH_THREADS, 1)
umma_chain_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmW0,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;                                   // [plane][k-chunk] tiles
  uint8_t* sRing = sA + CH_A_BYTES;
  float* sStg = reinterpret_cast<float*>(sRing + CH_RING * CH_TILE);
  float* sRed = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(sStg) + CH_STG_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sRed) + CH_RED_BYTES);
  uint64_t* r_full = bars;             // [4]
  uint64_t* r_empty = bars + 4;        // [4]
  uint64_t* a0_full = bars + 8;        // [4]
  uint64_t* a0_empty = bars + 12;      // [4]
  uint64_t* acc0_full = bars + 16;
  uint64_t* a_ready = bars + 17;       // 256 arrivals
  uint64_t* acc1_full = bars + 18;     // [2]
  uint64_t* acc1_empty = bars + 20;    // [2] 128 arrivals
  uint64_t* a_reads_done = bars + 22;
  uint64_t* a2_ready = bars + 23;      // 256 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int kc0 = ceil_div(p.K0, 64);
  const int NH1 = ceil_div(p.N1, 128);
  const int n_acc = NH1 + (p.vjob ? 2 : 0);

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmA0); umma::prefetch_tmap(&tmW0); umma::prefetch_tmap(&tmW1);
    if (p.vjob) umma::prefetch_tmap(&tmW2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 4; ++i) {
      umma::mbar_init(&r_full[i], 1); umma::mbar_init(&r_empty[i], 1);
      umma::mbar_init(&a0_full[i], 1); umma::mbar_init(&a0_empty[i], 1);
    }
    umma::mbar_init(acc0_full, 1); umma::mbar_init(a_ready, 256);
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&acc1_full[i], 1); umma::mbar_init(&acc1_empty[i], 128); }
    umma::mbar_init(a_reads_done, 1); umma::mbar_init(a2_ready, 256);
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
    int rs = 0; uint32_t rph = 0;
    auto ring_load = [&](const CUtensorMap* tm, int k0, int row0, int plane) {
      umma::mbar_wait(&r_empty[rs], rph ^ 1);
      if (umma::elect_one()) {
        umma::mbar_expect_tx(&r_full[rs], CH_TILE);
        umma::tma_load_3d(tm, &r_full[rs], sRing + rs * CH_TILE, k0, row0, plane);
      }
      __syncwarp();
      if (++rs == CH_RING) { rs = 0; rph ^= 1; }
    };
    for (int kc = 0; kc < kc0; ++kc) {
      const int slot = kc & 3;
      umma::mbar_wait(&a0_empty[slot], ((kc >> 2) & 1) ^ 1);
      if (umma::elect_one()) {
        umma::mbar_expect_tx(&a0_full[slot], 2 * CH_TILE);
        umma::tma_load_3d(&tmA0, &a0_full[slot], sA + slot * CH_TILE, kc * 64, m0, 0);
        umma::tma_load_3d(&tmA0, &a0_full[slot], sA + (4 + slot) * CH_TILE, kc * 64, m0, 1);
      }
      __syncwarp();
      for (int pw = 0; pw < 2; ++pw)
        for (int nh = 0; nh < 2; ++nh) ring_load(&tmW0, kc * 64, nh * 128, pw);
    }
    for (int h = 0; h < NH1; ++h)
      for (int kc = 0; kc < 4; ++kc)
        for (int pw = 0; pw < 2; ++pw) ring_load(&tmW1, kc * 64, h * 128, pw);
    if (p.vjob)
      for (int mh = 0; mh < 2; ++mh)
        for (int kc = 0; kc < 4; ++kc)
          for (int pw = 0; pw < 2; ++pw) ring_load(&tmW2, kc * 64, mh * 128, pw);
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma::idesc_bf16_f32(128, 128);
    constexpr uint32_t TU = CH_TILE >> 4;   // descriptor units per tile
    const uint32_t loA = umma::desc_lo(umma::smem_u32(sA));
    const uint32_t loR = umma::desc_lo(umma::smem_u32(sRing));
    int rs = 0; uint32_t rph = 0;
    // ---- GEMM0: both 128-column halves of acc0 advance together (A0 chunk loaded once)
    for (int kc = 0; kc < kc0; ++kc) {
      const int slot = kc & 3;
      umma::mbar_wait(&a0_full[slot], (kc >> 2) & 1);
      for (int pw = 0; pw < 2; ++pw)
        for (int nh = 0; nh < 2; ++nh) {
          umma::mbar_wait(&r_full[rs], rph);
          umma::fence_after();
          if (umma::elect_one()) {
            const uint32_t lob = loR + rs * TU;
            const uint32_t d = tmem_base + nh * 128;
            if (pw == 0) {
#pragma unroll
              for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma::mma_bf16(d, umma::desc_make(loA + (i * 4 + slot) * TU + 2 * k), umma::desc_make(lob + 2 * k), idesc,
                                 (kc | i | k) != 0 ? 1u : 0u);
            } else {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma::mma_bf16(d, umma::desc_make(loA + slot * TU + 2 * k), umma::desc_make(lob + 2 * k), idesc, 1u);
            }
            umma::mma_commit(&r_empty[rs]);
            if (pw == 1 && nh == 1) {
              umma::mma_commit(&a0_empty[slot]);
              if (kc == kc0 - 1) umma::mma_commit(acc0_full);
            }
          }
          __syncwarp();
          if (++rs == CH_RING) { rs = 0; rph ^= 1; }
        }
    }
    // ---- GEMM1 (A = planes written by E_A) and the V job (A = W2 tile from the ring, B = un-rotated planes)
    for (int h = 0; h < n_acc; ++h) {
      const bool vj = h >= NH1;
      if (h == 0) { umma::mbar_wait(a_ready, 0); umma::fence_after(); }
      if (h == NH1 && vj) { umma::mbar_wait(a2_ready, 0); umma::fence_after(); }
      const int buf = h & 1;
      if (h >= 2) { umma::mbar_wait(&acc1_empty[buf], ((h >> 1) - 1) & 1); umma::fence_after(); }
      const uint32_t d = tmem_base + 256 + buf * 128;
      for (int kc = 0; kc < 4; ++kc)
        for (int pw = 0; pw < 2; ++pw) {
          umma::mbar_wait(&r_full[rs], rph);
          umma::fence_after();
          if (umma::elect_one()) {
            const uint32_t low = loR + rs * TU;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              if (pw == 1 && i == 1) break;          // plane pairs (0,0) (1,0) | (0,1)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t dact = umma::desc_make(loA + (i * 4 + kc) * TU + 2 * k);   // activation plane i
                const uint64_t dw = umma::desc_make(low + 2 * k);                         // weight plane pw
                const uint32_t accf = (kc | pw | i | k) != 0 ? 1u : 0u;
                if (vj) umma::mma_bf16(d, dw, dact, idesc, accf);
                else umma::mma_bf16(d, dact, dw, idesc, accf);
              }
            }
            umma::mma_commit(&r_empty[rs]);
            if (kc == 3 && pw == 1) {
              umma::mma_commit(&acc1_full[buf]);
              if (h == NH1 - 1) umma::mma_commit(a_reads_done);
            }
          }
          __syncwarp();
          if (++rs == CH_RING) { rs = 0; rph ^= 1; }
        }
    }
  } else if (warp >= 4) {
    // ================= epilogue warpgroups =================
    const int wg = (warp - 4) >> 2;
    const int wq = warp & 3;
    const int trow = wq * 32 + lane;                       // row inside the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(wq * 32) << 16;
    float* stg = sStg + (warp - 4) * 1024;                 // [32][32] fp32, unit (q) of row r stored at unit q ^ (r & 7)
    const int rsub = lane >> 3, uq = lane & 7;             // transposed phase: lane -> (row sub-index, 16-byte unit)
    const uint32_t sA_u32 = umma::smem_u32(sA);

    // ---------------- E_A pass 1: x = x + film(acc0 + b0); keep x in TMEM; row sums
    umma::mbar_wait(acc0_full, 0);
    umma::fence_after();
    float sum = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const int c = wg * 4 + cc;
      float v[32];
      umma::tmem_ld32(tmem_base + lane_addr + c * 32, v);
      umma::tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      __syncwarp();
      const int col = c * 32 + uq * 4;
      const float4 bb = __ldg(reinterpret_cast<const float4*>(p.bias0 + col));
#pragma unroll
      for (int grp = 0; grp < 2; ++grp) {
        float4 av[4], xv[4], scv[4], shv[4];
        bool ok[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = (grp * 4 + it) * 4 + rsub;
          const int grow = m0 + wq * 32 + r;
          ok[it] = grow < p.M;
          av[it] = *reinterpret_cast<const float4*>(stg + r * 32 + ((uq ^ (r & 7)) << 2));
          xv[it] = scv[it] = shv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok[it] && p.film_mode) {
            const float* fs = p.film + (long long)(grow / p.T) * p.film_ld;
            scv[it] = __ldg(reinterpret_cast<const float4*>(fs + p.film_scale_off + col));
            shv[it] = __ldg(reinterpret_cast<const float4*>(fs + p.film_shift_off + col));
            xv[it] = *reinterpret_cast<const float4*>(p.x + (long long)grow * 256 + col);
          }
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = (grp * 4 + it) * 4 + rsub;
          const int grow = m0 + wq * 32 + r;
          const float4 a = av[it];
          float4 o = make_float4(a.x + bb.x, a.y + bb.y, a.z + bb.z, a.w + bb.w);
          if (p.film_mode) {
            const float4 sc = scv[it], sh = shv[it], x = xv[it];
            o = make_float4(x.x + ((sc.x + 1.f) * o.x + sh.x), x.y + ((sc.y + 1.f) * o.y + sh.y),
                            x.z + ((sc.z + 1.f) * o.z + sh.z), x.w + ((sc.w + 1.f) * o.w + sh.w));
          }
          if (ok[it]) *reinterpret_cast<float4*>(p.x + (long long)grow * 256 + col) = o;
          *reinterpret_cast<float4*>(stg + r * 32 + ((uq ^ (r & 7)) << 2)) = o;
        }
      }
      __syncwarp();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 t = *reinterpret_cast<const float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2));
        v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) sum += v[j];
      tmem_st32(tmem_base + lane_addr + c * 32, v);
      __syncwarp();
    }
    tmem_st_wait();
    // ---------------- row statistics (two-pass LayerNorm; the two warpgroups own 128 columns each)
    float mean = 0.f, rstd = 1.f;
    if (p.ln_mode) {
      sRed[wg * 128 + trow] = sum;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mean = (sRed[trow] + sRed[128 + trow]) / 256.f;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      float qs = 0.f;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        float v[32];
        umma::tmem_ld32(tmem_base + lane_addr + (wg * 4 + cc) * 32, v);
        umma::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float d_ = v[j] - mean; qs += d_ * d_; }
      }
      sRed[wg * 128 + trow] = qs;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      rstd = rsqrtf((sRed[trow] + sRed[128 + trow]) / 256.f + 1e-5f);
    }
    // ---------------- planes of (rotated) LayerNorm(x) -> shared memory in the UMMA A-operand layout
    auto emit_planes = [&](bool rot) {
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        const int c = wg * 4 + cc;
        float v[32];
        umma::tmem_ld32(tmem_base + lane_addr + c * 32, v);
        if (rot) {   // stage this warp's 32 table rows (16 (cos, sin) pairs = 128 B each) with coalesced loads
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub;
            const int pos = (m0 + wq * 32 + r) % p.T;
            const float4 t = __ldg(reinterpret_cast<const float4*>(p.rope_tab + (long long)pos * 128) + c * 8 + uq);
            *reinterpret_cast<float4*>(stg + r * 32 + ((uq ^ (r & 7)) << 2)) = t;
          }
          __syncwarp();
        }
        umma::tmem_ld_wait();
        if (p.ln_mode) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 ww = __ldg(reinterpret_cast<const float4*>(p.ln_w + c * 32 + 4 * q));
            const float4 bb = __ldg(reinterpret_cast<const float4*>(p.ln_b + c * 32 + 4 * q));
            v[4 * q + 0] = (v[4 * q + 0] - mean) * rstd * ww.x + bb.x;
            v[4 * q + 1] = (v[4 * q + 1] - mean) * rstd * ww.y + bb.y;
            v[4 * q + 2] = (v[4 * q + 2] - mean) * rstd * ww.z + bb.z;
            v[4 * q + 3] = (v[4 * q + 3] - mean) * rstd * ww.w + bb.w;
          }
        }
        if (rot) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 cs = *reinterpret_cast<const float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2));   // (cos0, sin0, cos1, sin1)
            const float h0 = v[4 * q], h1 = v[4 * q + 1], h2 = v[4 * q + 2], h3 = v[4 * q + 3];
            v[4 * q + 0] = h0 * cs.x - h1 * cs.y; v[4 * q + 1] = h1 * cs.x + h0 * cs.y;
            v[4 * q + 2] = h2 * cs.z - h3 * cs.w; v[4 * q + 3] = h3 * cs.z + h2 * cs.w;
          }
          __syncwarp();   // staging tile is reused by the next chunk
        }
        const int kc = c >> 1, ub = (c & 1) * 4;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t pk[2][4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t sp[2];
            umma::split_bf16_pair<2>(v[8 * u + 2 * e], v[8 * u + 2 * e + 1], sp);
            pk[0][e] = sp[0]; pk[1][e] = sp[1];
          }
#pragma unroll
          for (int t = 0; t < 2; ++t)
            st_shared_v4u(sA_u32 + (t * 4 + kc) * CH_TILE + trow * 128 + (((ub + u) ^ (trow & 7)) << 4), pk[t][0], pk[t][1], pk[t][2], pk[t][3]);
        }
      }
      umma::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
    };
    emit_planes(p.rope != 0);
    umma::mbar_arrive(a_ready);

    // ---------------- E_B: this warpgroup drains accumulator halves h = wg, wg + 2, ...
    bool vprep_done = false;
#pragma unroll 1
    for (int h = wg; h < n_acc; h += 2) {
      const bool vj = h >= NH1;
      if (vj && !vprep_done) {
        // every GEMM1 MMA has read the rotated planes: overwrite them with the un-rotated ones for the V job
        umma::mbar_wait(a_reads_done, 0);
        emit_planes(false);
        umma::mbar_arrive(a2_ready);
        vprep_done = true;
      }
      const int buf = h & 1;
      umma::mbar_wait(&acc1_full[buf], (h >> 1) & 1);
      umma::fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        {
          float v[32];
          umma::tmem_ld32(tmem_base + lane_addr + 256 + buf * 128 + c * 32, v);
          umma::tmem_ld_wait();
          if (c == 3) { umma::fence_before(); umma::mbar_arrive(&acc1_empty[buf]); }
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        __syncwarp();
        // normal: rows = tokens, columns = output features.  V job: rows = output channels, columns = tokens.
        const int col = (vj ? m0 : h * 128) + c * 32 + uq * 4;
        const bool col_ok = vj ? (col < p.M) : (col < p.N1);
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!vj && col_ok && p.bias1) bb = __ldg(reinterpret_cast<const float4*>(p.bias1 + col));
        const float osc = (vj || (p.scale_ncols != 0 && col >= p.scale_ncols)) ? 1.f : p.out_scale;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int r = it * 4 + rsub;
          const float4 a = *reinterpret_cast<const float4*>(stg + r * 32 + ((uq ^ (r & 7)) << 2));
          float o[4] = {a.x, a.y, a.z, a.w};
          long long orow;
          bool ok = col_ok;
          if (vj) {
            const int ch = (h - NH1) * 128 + wq * 32 + r;
            const float b2 = p.bias2 ? __ldg(p.bias2 + ch) : 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] += b2;
            orow = ch;
          } else {
            const int grow = m0 + wq * 32 + r;
            ok = ok && grow < p.M;
            o[0] += bb.x; o[1] += bb.y; o[2] += bb.z; o[3] += bb.w;
            if (p.gelu) {
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] = gelu_erf(o[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) o[j] *= osc;
            }
            orow = grow;
            if (p.remap_rps > 0) orow += (long long)(grow / p.remap_rps + 1) * p.remap_pad;
          }
          if (!ok) continue;
          uint32_t pk[2][2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            uint32_t sp[2];
            umma::split_bf16_pair<2>(o[2 * e], o[2 * e + 1], sp);
            pk[0][e] = sp[0]; pk[1][e] = sp[1];
          }
          __nv_bfloat16* dst = vj ? p.Vt + orow * p.ldvt + col : p.Cp + orow * p.ldcp + col;
          const long long ps = vj ? p.vt_plane_stride : p.cp_plane_stride;
#pragma unroll
          for (int t = 0; t < 2; ++t) *reinterpret_cast<uint2*>(dst + t * ps) = make_uint2(pk[t][0], pk[t][1]);
        }
        __syncwarp();
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
struct ChainOperands {
  const __nv_bfloat16* A0; long long a0_rows, a0_ld, a0_plane_stride;   // [2][a0_rows][a0_ld], K0 valid columns
  const __nv_bfloat16* W0; long long w0_plane_stride;                   // [2][256][K0]
  const __nv_bfloat16* W1; long long w1_plane_stride;                   // [2][N1][256]
  const __nv_bfloat16* W2; long long w2_plane_stride;                   // [2][256][256] (null without a V job)
};

inline int launch_umma_chain(const ChainOperands& o, const ChainParams& p, cudaStream_t st) {
  if (p.K0 % 8 || p.N1 % 8 || p.N1 <= 0 || p.N1 > 1024) A2P_FAIL("chain: bad K0=%d / N1=%d", p.K0, p.N1);
  if (p.vjob && (!o.W2 || !p.Vt)) A2P_FAIL("chain: V job needs W2 and Vt");
  CUtensorMap tA0, tW0, tW1, tW2;
  const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
  A2P_TRY(make_tmap_bf16_3d(&tA0, o.A0, p.K0, o.a0_rows, 2, o.a0_ld, o.a0_plane_stride, 64, 128, sw));
  A2P_TRY(make_tmap_bf16_3d(&tW0, o.W0, p.K0, 256, 2, p.K0, o.w0_plane_stride, 64, 128, sw));
  A2P_TRY(make_tmap_bf16_3d(&tW1, o.W1, 256, p.N1, 2, 256, o.w1_plane_stride, 64, 128, sw));
  if (p.vjob) A2P_TRY(make_tmap_bf16_3d(&tW2, o.W2, 256, 256, 2, 256, o.w2_plane_stride, 64, 128, sw));
  else tW2 = tW1;
  const int grid = ceil_div(p.M, 128);
  A2P_CUDA(launch_pdl(umma_chain_kernel, dim3(grid), dim3(CH_THREADS), (size_t)CH_SMEM_BYTES, st, tA0, tW0, tW1, tW2, p));
  return 0;
}

inline int init_umma_chain() {
  A2P_CUDA(cudaFuncSetAttribute(umma_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES));
  return 0;
}

}  // namespace a2p
