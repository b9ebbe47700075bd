// Split-bf16 tensor-core GEMM for sm_100a:   C[M,N] = epi( sum_{(i,j) in products} A_i[M,K] * W_j[N,K]^T + bias )
//
// fp32 operands are pre-split into TERMS bf16 planes (x = p0 + p1 (+ p2), umma.cuh); the kernel multiplies
// the plane pairs with i + j < TERMS (3 products for TERMS = 2: ~2^-17 relative error; 6 for TERMS = 3:
// ~fp32) on the 5th-gen tensor cores and accumulates everything in ONE fp32 TMEM accumulator.
//
// Structure (persistent, warp-specialised, one CTA per SM):
//   warp 0   TMA producer   : per k-block loads TERMS A tiles + TERMS W tiles (128x64 bf16, SWIZZLE_128B)
//   warp 1   MMA issuer     : one elected thread issues tcgen05.mma (M=128, N=128, K=16), commits to mbarriers
//   warp 2   TMEM allocator : 256 columns = two 128x128 fp32 accumulators (epilogue of tile i overlaps MMA of i+1)
//   warps 4-7 epilogue      : tcgen05.ld 32x32b -> registers -> fused epilogue -> global
// "taps" > 1 turns it into the causal dilated Conv1d of the pose TCN (tap j reads A rows shifted by
// (taps-1-j)*dil; TMA zero-fills the negative rows).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "umma.cuh"

namespace a2p {

enum TcEpi : int {
  TC_F32 = 0,          // C = (acc + bias) * out_scale                    (fp32 out)
  TC_FILM = 1,         // C += (film_scale + 1) * (acc + bias) + film_shift  (fp32 residual stream)
  TC_GELU_PLANES = 2,  // planes( gelu(acc + bias) )                       (bf16 planes out)
  TC_PLANES = 3,       // planes( (acc + bias) * out_scale )
  TC_LRELU_PLANES = 4, // planes( lrelu(acc+bias) ) ; with skip: planes( (skip + lrelu)/2 ), fp32 copy kept in C
};

struct TcGemmParams {
  int M, N, K, taps, dil;
  const float* bias;
  float* C; long long ldc;
  __nv_bfloat16* Cp; long long cp_plane_stride; long long ldcp;
  const float* film; long long film_ld; int film_scale_off, film_shift_off, rows_per_sample;
  const float* skip; long long ldskip;
  float out_scale, slope;
  int scale_ncols;   // out_scale applies to columns < scale_ncols (0 = all columns)
  int bias_per_row;  // bias indexed by output row (swapped-operand GEMMs producing a transposed result)
  int remap_rps, remap_pad;  // if remap_rps > 0: output row = row + (row / remap_rps + 1) * remap_pad (write into a per-sample left-padded layout)
};

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 2;  // 16 KB (A and W tiles have the same size)

template <int TERMS>
struct TcCfg {
  static constexpr int NPROD = TERMS == 1 ? 1 : (TERMS == 2 ? 3 : 6);
  static constexpr int STAGES = TERMS == 1 ? 6 : (TERMS == 2 ? 3 : 2);
  static constexpr int STAGE_BYTES = 2 * TERMS * TC_TILE_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/ + 4 * 32 * 36 * 4 /*epilogue staging*/;
};

// plane pairs (i, j) with i + j < TERMS, most significant first (compile-time tables: indices fold into immediates)
__host__ __device__ constexpr int prod_a(int pr) { return pr == 0 ? 0 : pr == 1 ? 0 : pr == 2 ? 1 : pr == 3 ? 0 : pr == 4 ? 2 : 1; }
__host__ __device__ constexpr int prod_b(int pr) { return pr == 0 ? 0 : pr == 1 ? 1 : pr == 2 ? 0 : pr == 3 ? 2 : pr == 4 ? 0 : 1; }

template <int TERMS, int EPI>
__global__ void __launch_bounds__(256, 1) umma_gemm_kernel(const __grid_constant__ CUtensorMap tmA,
                                                           const __grid_constant__ CUtensorMap tmW, TcGemmParams p) {
  using Cfg = TcCfg<TERMS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]
  uint64_t* tfull = bars + 2 * Cfg::STAGES;    // [2]
  uint64_t* tempty = tfull + 2;                // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* stage_tile = reinterpret_cast<float*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES + 256);   // 4 warps x [32][36] fp32

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_n = ceil_div(p.N, TC_BN), tiles_m = ceil_div(p.M, TC_BM);
  const int n_tiles = tiles_m * tiles_n;
  const int kb_per_tap = ceil_div(p.K, TC_BK);
  const int n_kb = kb_per_tap * p.taps;

  if (warp == 0 && lane == 0) {
    umma::prefetch_tmap(&tmA);
    umma::prefetch_tmap(&tmW);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::STAGES; ++i) { umma::mbar_init(&full[i], 1); umma::mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&tfull[i], 1); umma::mbar_init(&tempty[i], 128); }
    umma::fence_barrier_init();
  }
  if (warp == 2) umma::tmem_alloc<2 * TC_BN>(tmem_slot);
  umma::fence_before();
  __syncthreads();
  umma::fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();      // everything above overlapped the predecessor's tail; its outputs are visible from here on

  // NOTE on the two single-thread roles: the whole warp runs the loops (all values stay warp-uniform, so the
  // compiler keeps descriptors / TMEM addresses in uniform registers) and only the issuing instructions sit under
  // elect.sync -- an `if (lane == 0)` region makes ptxas wrap every UTCHMMA / UTMALDG in an ELECT/R2UR waterfall loop.
  if (warp == 0) {
    // ================= TMA producer =================
    int stage = 0; uint32_t phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * TC_BM, n0 = (t % tiles_n) * TC_BN;
      for (int kb = 0; kb < n_kb; ++kb) {
        const int tap = kb / kb_per_tap, k0 = (kb - tap * kb_per_tap) * TC_BK;
        const int shift = (p.taps - 1 - tap) * p.dil;
        umma::mbar_wait(&empty[stage], phase ^ 1);
        if (umma::elect_one()) {
          umma::mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
#pragma unroll
          for (int i = 0; i < TERMS; ++i) umma::tma_load_3d(&tmA, &full[stage], sa + i * TC_TILE_BYTES, k0, m0 - shift, i);
#pragma unroll
          for (int i = 0; i < TERMS; ++i)
            umma::tma_load_3d(&tmW, &full[stage], sa + (TERMS + i) * TC_TILE_BYTES, k0, n0, i * p.taps + tap);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
    }
    pdl_trigger();   // all operand tiles requested: the successor's prologue may overlap this CTA's last tiles (no-op without A2P_PDL)
  } else if (warp == 1) {
    // ================= MMA issuer =================
    constexpr uint32_t idesc = umma::idesc_bf16_f32(TC_BM, TC_BN);
    const uint32_t lo0 = umma::desc_lo(umma::smem_u32(smem));
    int stage = 0; uint32_t phase = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      umma::mbar_wait(&tempty[acc], acc_phase ^ 1);
      umma::fence_after();
      const uint32_t d = tmem_base + acc * TC_BN;
      for (int kb = 0; kb < n_kb; ++kb) {
        umma::mbar_wait(&full[stage], phase);
        umma::fence_after();
        if (umma::elect_one()) {
          const uint32_t lo = lo0 + stage * (Cfg::STAGE_BYTES >> 4);
#pragma unroll
          for (int pr = 0; pr < Cfg::NPROD; ++pr) {
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma::mma_bf16(d, umma::desc_make(lo + prod_a(pr) * (TC_TILE_BYTES >> 4) + 2 * k),
                             umma::desc_make(lo + (TERMS + prod_b(pr)) * (TC_TILE_BYTES >> 4) + 2 * k), idesc,
                             (kb | pr | k) != 0 ? 1u : 0u);
          }
          umma::mma_commit(&empty[stage]);
          if (kb == n_kb - 1) umma::mma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4) {
    // ================= epilogue =================
    const int wq = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
      const int m0 = (t / tiles_n) * TC_BM, n0 = (t % tiles_n) * TC_BN;
      umma::mbar_wait(&tfull[acc], acc_phase);
      umma::fence_after();
      // TMEM gives each thread one ROW (32 consecutive columns).  Storing that way touches 32 different 128-B
      // lines per instruction, so the 32x32 chunk is transposed through a per-warp smem staging tile
      // ([32][36] floats, 128-bit accesses conflict-free both ways) and written 4 rows x 128 B per instruction.
      float* tw = stage_tile + wq * (32 * 36);
      const int rsub = lane >> 3, c4 = (lane & 7) * 4;
#pragma unroll 1
      for (int c = 0; c < TC_BN / 32; ++c) {
        const int col = n0 + c * 32 + c4;
        if (n0 + c * 32 >= p.N) break;   // warp-uniform
        {
          float v[32];
          umma::tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + acc * TC_BN + c * 32, v);
          umma::tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(tw + lane * 36 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
        __syncwarp();
        const bool col_ok = col < p.N;
        float4 bb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias && col_ok && !p.bias_per_row) bb = __ldg(reinterpret_cast<const float4*>(p.bias + col));
        const float oscale = (p.scale_ncols == 0 || col < p.scale_ncols) ? p.out_scale : 1.f;
        // two groups of 4 row-slices: all global loads of a group (residual x / skip / FiLM vectors) are issued
        // BEFORE any store, otherwise every load->store pair serialises on a ~1 us global-memory round trip
        // (the compiler cannot hoist loads over the stores that may alias them).
#pragma unroll
        for (int grp = 0; grp < 2; ++grp) {
          float4 av[4], xv[4], scv[4], shv[4];
          int rows[4];
          bool ok[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int r = (grp * 4 + q) * 4 + rsub;
            rows[q] = m0 + wq * 32 + r;
            ok[q] = rows[q] < p.M && col_ok;
            av[q] = *reinterpret_cast<const float4*>(tw + r * 36 + c4);
            xv[q] = scv[q] = shv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok[q]) {
              if (EPI == TC_FILM) {
                const float* fs = p.film + (long long)(rows[q] / p.rows_per_sample) * p.film_ld;
                scv[q] = __ldg(reinterpret_cast<const float4*>(fs + p.film_scale_off + col));
                shv[q] = __ldg(reinterpret_cast<const float4*>(fs + p.film_shift_off + col));
                xv[q] = *reinterpret_cast<const float4*>(p.C + (long long)rows[q] * p.ldc + col);
              } else if (EPI == TC_LRELU_PLANES) {
                if (p.skip) xv[q] = *reinterpret_cast<const float4*>(p.skip + (long long)rows[q] * p.ldskip + col);
              }
              if (p.bias_per_row && p.bias) { const float br_ = __ldg(p.bias + rows[q]); scv[q].x = br_; }
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (!ok[q]) continue;
            int row = rows[q];
            const float4 a = av[q];
            float4 bq = bb;
            if (p.bias_per_row && p.bias) bq = make_float4(scv[q].x, scv[q].x, scv[q].x, scv[q].x);
            float o[4] = {a.x + bq.x, a.y + bq.y, a.z + bq.z, a.w + bq.w};
            if (p.remap_rps > 0) row += (row / p.remap_rps + 1) * p.remap_pad;
            if (EPI == TC_F32) {
              *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) =
                  make_float4(o[0] * oscale, o[1] * oscale, o[2] * oscale, o[3] * oscale);
            } else if (EPI == TC_FILM) {
              const float4 sc = scv[q], sh = shv[q], x = xv[q];
              *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) =
                  make_float4(x.x + ((sc.x + 1.f) * o[0] + sh.x), x.y + ((sc.y + 1.f) * o[1] + sh.y),
                              x.z + ((sc.z + 1.f) * o[2] + sh.z), x.w + ((sc.w + 1.f) * o[3] + sh.w));
            } else {
              if (EPI == TC_GELU_PLANES) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = gelu_erf(o[j]);
              } else if (EPI == TC_LRELU_PLANES) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * p.slope;
                if (p.skip) {
                  const float4 sk = xv[q];
                  o[0] = (sk.x + o[0]) / 2.0f; o[1] = (sk.y + o[1]) / 2.0f; o[2] = (sk.z + o[2]) / 2.0f; o[3] = (sk.w + o[3]) / 2.0f;
                }
                if (p.C) *reinterpret_cast<float4*>(p.C + (long long)row * p.ldc + col) = make_float4(o[0], o[1], o[2], o[3]);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] *= oscale;
              }
              uint32_t pk[TERMS][2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                uint32_t sp[TERMS];
                umma::split_bf16_pair<TERMS>(o[2 * e], o[2 * e + 1], sp);
#pragma unroll
                for (int t = 0; t < TERMS; ++t) pk[t][e] = sp[t];
              }
#pragma unroll
              for (int t = 0; t < TERMS; ++t)
                *reinterpret_cast<uint2*>(p.Cp + t * p.cp_plane_stride + (long long)row * p.ldcp + col) = make_uint2(pk[t][0], pk[t][1]);
            }
          }
        }
        __syncwarp();
      }
      umma::fence_before();
      umma::mbar_arrive(&tempty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  __syncthreads();
  if (warp == 2) {
    umma::fence_after();
    umma::tmem_dealloc<2 * TC_BN>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// bf16 tensor [planes][rows][ld] (ld >= cols, contiguous cols); box = (box_cols, box_rows, 1)
inline int make_tmap_bf16_3d(CUtensorMap* tm, const void* base, long long cols, long long rows, long long planes, long long ld,
                             long long plane_stride, int box_cols, int box_rows, CUtensorMapSwizzle swz, int box_planes = 1) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) A2P_FAIL("cuTensorMapEncodeTiled entry point not available");
  if ((ld * 2) % 16 || (plane_stride * 2) % 16 || (reinterpret_cast<uintptr_t>(base) % 16))
    A2P_FAIL("TMA operand not 16-byte aligned (ld=%lld plane_stride=%lld)", ld, plane_stride);
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)planes};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)plane_stride * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, (cuuint32_t)box_planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) A2P_FAIL("cuTensorMapEncodeTiled failed (%d) cols=%lld rows=%lld planes=%lld ld=%lld", (int)r, cols, rows, planes, ld);
  return 0;
}

struct TcOperands {
  const __nv_bfloat16* A; long long lda, a_plane_stride;   // [TERMS][M][lda]
  const __nv_bfloat16* W; long long ldw, w_plane_stride;   // [TERMS*taps][N][ldw]
};

template <int TERMS, int EPI>
int launch_umma_gemm_t(const TcOperands& o, const TcGemmParams& p, int num_sms, cudaStream_t st) {
  using Cfg = TcCfg<TERMS>;
  CUtensorMap tmA, tmW;
  A2P_TRY(make_tmap_bf16_3d(&tmA, o.A, p.K, p.M, TERMS, o.lda, o.a_plane_stride, TC_BK, TC_BM, CU_TENSOR_MAP_SWIZZLE_128B));
  A2P_TRY(make_tmap_bf16_3d(&tmW, o.W, p.K, p.N, (long long)TERMS * p.taps, o.ldw, o.w_plane_stride, TC_BK, TC_BN,
                            CU_TENSOR_MAP_SWIZZLE_128B));
  const int n_tiles = ceil_div(p.M, TC_BM) * ceil_div(p.N, TC_BN);
  const int grid = n_tiles < num_sms ? n_tiles : num_sms;
  A2P_CUDA(launch_pdl(umma_gemm_kernel<TERMS, EPI>, dim3(grid), dim3(256), (size_t)Cfg::SMEM_BYTES, st, tmA, tmW, p));
  return 0;
}

template <int TERMS>
int launch_umma_gemm_e(const TcOperands& o, const TcGemmParams& p, int epi, int num_sms, cudaStream_t st) {
  switch (epi) {
    case TC_F32: return launch_umma_gemm_t<TERMS, TC_F32>(o, p, num_sms, st);
    case TC_FILM: return launch_umma_gemm_t<TERMS, TC_FILM>(o, p, num_sms, st);
    case TC_GELU_PLANES: return launch_umma_gemm_t<TERMS, TC_GELU_PLANES>(o, p, num_sms, st);
    case TC_PLANES: return launch_umma_gemm_t<TERMS, TC_PLANES>(o, p, num_sms, st);
    case TC_LRELU_PLANES: return launch_umma_gemm_t<TERMS, TC_LRELU_PLANES>(o, p, num_sms, st);
  }
  A2P_FAIL("umma_gemm: unknown epilogue %d", epi);
}

inline int launch_umma_gemm(int terms, const TcOperands& o, const TcGemmParams& p, int epi, int num_sms, cudaStream_t st) {
  if (p.N % 8 || p.K % 8) A2P_FAIL("umma_gemm: N and K must be multiples of 8 (N=%d K=%d)", p.N, p.K);
  if (terms == 1) return launch_umma_gemm_e<1>(o, p, epi, num_sms, st);
  if (terms == 2) return launch_umma_gemm_e<2>(o, p, epi, num_sms, st);
  if (terms == 3) return launch_umma_gemm_e<3>(o, p, epi, num_sms, st);
  A2P_FAIL("umma_gemm: terms must be 1, 2 or 3");
}

// pre-set the >48 KB dynamic smem attribute of every instantiation (must not happen inside a stream capture)
inline int init_umma_gemm() {
#define A2P_SET(T, E) A2P_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<T, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<T>::SMEM_BYTES));
#define A2P_SET_ALL(T) A2P_SET(T, TC_F32) A2P_SET(T, TC_FILM) A2P_SET(T, TC_GELU_PLANES) A2P_SET(T, TC_PLANES) A2P_SET(T, TC_LRELU_PLANES)
  A2P_SET_ALL(1) A2P_SET_ALL(2) A2P_SET_ALL(3)
#undef A2P_SET_ALL
#undef A2P_SET
  return 0;
}

// fp32 [rows, cols] (row stride ld) -> bf16 planes [TERMS][rows][cols]
template <int TERMS>
__global__ void split_planes_kernel(const float* __restrict__ src, long long ld, __nv_bfloat16* __restrict__ dst,
                                    long long plane_stride, long long rows, int cols, float scale) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int c4 = cols / 4;
  if (idx >= rows * c4) return;
  long long r = idx / c4;
  int c = (int)(idx - r * c4) * 4;
  float4 v = *reinterpret_cast<const float4*>(src + r * ld + c);
  const float in[4] = {v.x * scale, v.y * scale, v.z * scale, v.w * scale};
  __nv_bfloat16 pl[TERMS][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat16 s[TERMS];
    umma::split_bf16<TERMS>(in[j], s);
#pragma unroll
    for (int i = 0; i < TERMS; ++i) pl[i][j] = s[i];
  }
#pragma unroll
  for (int i = 0; i < TERMS; ++i)
    *reinterpret_cast<uint2*>(dst + i * plane_stride + r * cols + c) = *reinterpret_cast<const uint2*>(pl[i]);
}

inline int launch_split_planes(int terms, const float* src, long long ld, __nv_bfloat16* dst, long long plane_stride,
                               long long rows, int cols, float scale, cudaStream_t st) {
  if (cols % 4) A2P_FAIL("split_planes: cols must be a multiple of 4");
  long long total = rows * (cols / 4);
  unsigned blocks = (unsigned)((total + 255) / 256);
  if (terms == 1) split_planes_kernel<1><<<blocks, 256, 0, st>>>(src, ld, dst, plane_stride, rows, cols, scale);
  else if (terms == 2) split_planes_kernel<2><<<blocks, 256, 0, st>>>(src, ld, dst, plane_stride, rows, cols, scale);
  else if (terms == 3) split_planes_kernel<3><<<blocks, 256, 0, st>>>(src, ld, dst, plane_stride, rows, cols, scale);
  else A2P_FAIL("split_planes: terms must be 1..3");
  A2P_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace a2p
