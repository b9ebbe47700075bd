// Exact-fp32 FFMA GEMM with fused epilogues:  C[M,N] = epi(A[M,K] * W[N,K]^T + bias).
//
// This is the split_terms == 0 ("exact fp32") arm of every linear layer on the path and the on-device
// reference the tcgen05 kernels are unit-tested against.  A and W are both K-contiguous (activations
// row-major, nn.Linear weights [out,in]).  A "tap" mode turns it into the causal dilated Conv1d of
// the pose post-TCN (model/diffusion.py:201-224): K = taps*Kc and tap j reads A row (r - (taps-1-j)*dil).
#pragma once
#include "common.cuh"

namespace a2p {

enum Epi : int {
  EPI_BIAS = 0,          // acc + bias
  EPI_GELU = 1,          // exact-erf GELU (F.gelu, utils/model_util.py:69)
  EPI_MISH = 2,          // nn.Mish (time_mlp, model/diffusion.py:121-125)
  EPI_ADDROW_MISH = 3,   // mish(acc + bias + rowvec[row])  (t = to_time_cond(.) + cond_hidden; FiLM input Mish(t))
  EPI_FILM_RESID = 4,    // C += (film_scale + 1) * (acc + bias) + film_shift   (transformer_modules.py:122-124)
  EPI_LRELU = 5,         // leaky_relu(acc + bias, slope)
  EPI_LRELU_SKIPAVG = 6, // (skip + leaky_relu(acc + bias)) / 2           (model/diffusion.py:220-221)
  EPI_SILU = 7,          // nn.SiLU (non_attn_cond_projection, model/diffusion.py:175-180)
  EPI_RESID = 8          // C += acc + bias   (pre-LN encoder layer residuals, transformer_modules.py:73-76)
};

struct GemmParams {
  const float* A; long long lda;
  const float* W; long long ldw;
  const float* bias;
  float* C; long long ldc;
  int M, N, K;
  int taps, dil, Kc;
  int epi;
  BranchPtr rowvec;
  const float* film; long long film_ld; int film_scale_off; int film_shift_off; int rows_per_sample;
  const float* skip; long long ldskip;
  float slope;
};

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8;

__global__ void __launch_bounds__(256) sgemm_kernel(GemmParams p) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int bm = blockIdx.y * SG_BM, bn = blockIdx.x * SG_BN;
  const int lrow = tid >> 1, lk = (tid & 1) * 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = p.K / SG_BK;
  float4 ra, rb;
  auto gload = [&](int kt) {
    const int k0 = kt * SG_BK;
    const int tap = k0 / p.Kc;
    const int kc = k0 - tap * p.Kc + lk;
    const int arow = bm + lrow - (p.taps - 1 - tap) * p.dil;
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bm + lrow < p.M && arow >= 0) ra = *reinterpret_cast<const float4*>(p.A + (long long)arow * p.lda + kc);
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bn + lrow < p.N) rb = *reinterpret_cast<const float4*>(p.W + (long long)(bn + lrow) * p.ldw + k0 + lk);
  };
  auto sstore = [&](int buf) {
    As[buf][lk + 0][lrow] = ra.x; As[buf][lk + 1][lrow] = ra.y; As[buf][lk + 2][lrow] = ra.z; As[buf][lk + 3][lrow] = ra.w;
    Bs[buf][lk + 0][lrow] = rb.x; Bs[buf][lk + 1][lrow] = rb.y; Bs[buf][lk + 2][lrow] = rb.z; Bs[buf][lk + 3][lrow] = rb.w;
  };
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload(kt + 1);
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float a[8], b[8];
      *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

  // ---- epilogue
#pragma unroll
  for (int ih = 0; ih < 2; ++ih) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = bm + ih * 64 + ty * 4 + i;
      if (row >= p.M) continue;
      const float* rv = (p.epi == EPI_ADDROW_MISH) ? p.rowvec.at(row) : nullptr;
      const float* fs = nullptr;
      if (p.epi == EPI_FILM_RESID) fs = p.film + (long long)(row / p.rows_per_sample) * p.film_ld;
#pragma unroll
      for (int jh = 0; jh < 2; ++jh) {
        const int col = bn + jh * 64 + tx * 4;
        if (col >= p.N) continue;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = acc[ih * 4 + i][jh * 4 + j] + (p.bias ? p.bias[col + j] : 0.f);
        float* cptr = p.C + (long long)row * p.ldc + col;
        switch (p.epi) {
          case EPI_GELU:
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
            break;
          case EPI_MISH:
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = mishf(v[j]);
            break;
          case EPI_ADDROW_MISH:
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = mishf(v[j] + rv[col + j]);
            break;
          case EPI_FILM_RESID: {
            float4 x = *reinterpret_cast<const float4*>(cptr);
            const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float sc = fs[p.film_scale_off + col + j], sh = fs[p.film_shift_off + col + j];
              v[j] = xs[j] + ((sc + 1.f) * v[j] + sh);
            }
          } break;
          case EPI_LRELU:
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * p.slope;
            break;
          case EPI_SILU:
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = v[j] / (1.0f + expf(-v[j]));
            break;
          case EPI_RESID: {
            const float4 x = *reinterpret_cast<const float4*>(cptr);
            v[0] += x.x; v[1] += x.y; v[2] += x.z; v[3] += x.w;
          } break;
          case EPI_LRELU_SKIPAVG: {
            float4 s = *reinterpret_cast<const float4*>(p.skip + (long long)row * p.ldskip + col);
            const float ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float y = v[j] > 0.f ? v[j] : v[j] * p.slope;
              v[j] = (ss[j] + y) / 2.0f;
            }
          } break;
          default: break;
        }
        *reinterpret_cast<float4*>(cptr) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
}

inline int launch_sgemm(const GemmParams& p, cudaStream_t st) {
  if (p.K % SG_BK || p.Kc % SG_BK || p.N % 4 || p.lda % 4 || p.ldw % 4 || p.ldc % 4)
    A2P_FAIL("sgemm: unsupported shape M=%d N=%d K=%d Kc=%d lda=%lld ldc=%lld", p.M, p.N, p.K, p.Kc, p.lda, p.ldc);
  dim3 grid(ceil_div(p.N, SG_BN), ceil_div(p.M, SG_BM));
  sgemm_kernel<<<grid, 256, 0, st>>>(p);
  A2P_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace a2p

// ------------------------------------------------------------------ skinny GEMM (M <= 64 rows)
// The per-step conditioning linears (time MLP, FiLM table, time-token K/V rows) have M = 2B rows only: a
// 128x128-tile kernel leaves the chip idle and crawls through K.  Here one warp owns one output column n:
// it streams W[n,:] once (coalesced, read-only path) and keeps A (M x K fp32) in shared memory, so the
// kernel is bound by reading W once from L2/HBM.  Same fused epilogues as sgemm_kernel (no FiLM / skip).
namespace a2p {

template <int MB>
__global__ void __launch_bounds__(256) skinny_gemm_kernel(GemmParams p) {
  extern __shared__ __align__(16) float sA[];   // [rows of this block][K]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int row0 = blockIdx.y * MB;               // row block (the kernel choice never depends on M: batch-invariant results)
  const int mrows = ::min(MB, p.M - row0);
  for (int i = tid; i < mrows * (p.K / 4); i += 256) {
    const int r = i / (p.K / 4), c = i - r * (p.K / 4);
    reinterpret_cast<float4*>(sA)[i] = *reinterpret_cast<const float4*>(p.A + (long long)(row0 + r) * p.lda + c * 4);
  }
  __syncthreads();
  const int n = blockIdx.x * 8 + warp;
  if (n >= p.N) return;
  float acc[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = 0.f;
  const float* wr = p.W + (long long)n * p.ldw;
  for (int k = lane * 4; k < p.K; k += 128) {
    const float4 w = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      if (m < mrows) {
        const float4 a = *reinterpret_cast<const float4*>(sA + m * p.K + k);
        acc[m] = fmaf(a.x, w.x, acc[m]); acc[m] = fmaf(a.y, w.y, acc[m]);
        acc[m] = fmaf(a.z, w.z, acc[m]); acc[m] = fmaf(a.w, w.w, acc[m]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MB; ++m) {
#pragma unroll
    for (int o = 16; o; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
  }
  if (lane == 0) {
    const float b = p.bias ? p.bias[n] : 0.f;
#pragma unroll
    for (int m = 0; m < MB; ++m) {
      if (m >= mrows) break;
      float v = acc[m] + b;
      if (p.epi == EPI_GELU) v = gelu_erf(v);
      else if (p.epi == EPI_MISH) v = mishf(v);
      else if (p.epi == EPI_ADDROW_MISH) v = mishf(v + p.rowvec.at(row0 + m)[n]);
      p.C[(long long)(row0 + m) * p.ldc + n] = v;
    }
  }
}

inline int init_skinny_gemm() {
  A2P_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  A2P_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  A2P_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  return 0;
}

// true if the skinny kernel supports this GEMM (the CALL SITE decides whether to use it -- never the row count M,
// so that a batch row gets bit-identical results however the batch is sharded)
inline bool skinny_ok(const GemmParams& p) {
  return p.taps <= 1 && p.K % 4 == 0 && (size_t)16 * p.K * 4 <= 200 * 1024 &&
         (p.epi == EPI_BIAS || p.epi == EPI_GELU || p.epi == EPI_MISH || p.epi == EPI_ADDROW_MISH);
}

inline int launch_skinny_gemm(const GemmParams& p, cudaStream_t st) {
  dim3 grid(ceil_div(p.N, 8), ceil_div(p.M, 16));
  skinny_gemm_kernel<16><<<grid, 256, (size_t)16 * p.K * 4, st>>>(p);   // row blocks of 16: same per-row arithmetic for any M
  A2P_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace a2p
