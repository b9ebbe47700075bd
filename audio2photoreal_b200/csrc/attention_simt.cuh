// Exact-fp32 flash-style attention (FFMA), the split_terms == 0 arm of K1's softmax(QK^T/sqrt(dh))V core.
//
// One CTA = 64 queries of one (sample-row, head); keys stream through shared memory in tiles of 64 with
// an online softmax (running max / sum in registers).  Keys come from up to two sources so that the
// cached audio memory (rows 0..S_main-1, shared by every diffusion step) and the 2 per-step time-token
// rows (model/diffusion.py:392-393) never have to be concatenated in HBM.
#pragma once
#include <math.h>
#include "common.cuh"
#include "umma.cuh"

namespace a2p {

struct AttnParams {
  const float* Q; long long q_ld; long long q_sample_stride;
  BranchPtr K, V; long long kv_ld; int S_main;
  const float* Kx; const float* Vx; long long x_ld; long long x_sample_stride; int S_extra;
  float* O; long long o_ld; long long o_sample_stride;
  __nv_bfloat16* Op; long long op_plane_stride; int op_terms;   // optional split-bf16 planes output [terms][R*T][o_ld]
  int T, H, R;
  float scale_log2e;
};

template <int DH>
__global__ void __launch_bounds__(256) attn_simt_kernel(AttnParams p) {
  constexpr int BQ = 64, BK = 64, LDQ = DH + 4, LDP = BK + 4, NV = DH / 16;
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;
  float* Ks = Qs + BQ * LDQ;
  float* Vs = Ks + BK * LDQ;
  float* Ps = Vs + BK * DH;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int r = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int S_total = p.S_main + p.S_extra;

  const float* Qg = p.Q + (long long)r * p.q_sample_stride + h * DH;
  for (int idx = tid; idx < BQ * (DH / 4); idx += 256) {
    int row = idx / (DH / 4), c4 = idx - row * (DH / 4);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + row < p.T) v = *reinterpret_cast<const float4*>(Qg + (long long)(q0 + row) * p.q_ld + c4 * 4);
    *reinterpret_cast<float4*>(Qs + row * LDQ + c4 * 4) = v;
  }
  const float* Kg = p.K.at(r) + h * DH;
  const float* Vg = p.V.at(r) + h * DH;
  const float* Kxg = p.Kx ? p.Kx + (long long)r * p.x_sample_stride + h * DH : nullptr;
  const float* Vxg = p.Vx ? p.Vx + (long long)r * p.x_sample_stride + h * DH : nullptr;

  float m[4], l[4], o[4][NV];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) o[i][j] = 0.f;
  }

  for (int k0 = 0; k0 < S_total; k0 += BK) {
    __syncthreads();  // previous tile's Ks/Vs/Ps fully consumed (also covers the Q store on iter 0)
    for (int idx = tid; idx < BK * (DH / 4); idx += 256) {
      int row = idx / (DH / 4), c4 = idx - row * (DH / 4);
      int j = k0 + row;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (j < p.S_main) {
        kv = *reinterpret_cast<const float4*>(Kg + (long long)j * p.kv_ld + c4 * 4);
        vv = *reinterpret_cast<const float4*>(Vg + (long long)j * p.kv_ld + c4 * 4);
      } else if (j < S_total) {
        kv = *reinterpret_cast<const float4*>(Kxg + (long long)(j - p.S_main) * p.x_ld + c4 * 4);
        vv = *reinterpret_cast<const float4*>(Vxg + (long long)(j - p.S_main) * p.x_ld + c4 * 4);
      }
      *reinterpret_cast<float4*>(Ks + row * LDQ + c4 * 4) = kv;
      *reinterpret_cast<float4*>(Vs + row * DH + c4 * 4) = vv;
    }
    __syncthreads();
    // S = Q K^T : rows ty + 16 i, cols tx + 16 j
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll
    for (int d = 0; d < DH; d += 4) {
      float4 q[4], k[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) q[i] = *reinterpret_cast<const float4*>(Qs + (ty + 16 * i) * LDQ + d);
#pragma unroll
      for (int j = 0; j < 4; ++j) k[j] = *reinterpret_cast<const float4*>(Ks + (tx + 16 * j) * LDQ + d);
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          s[i][j] = fmaf(q[i].x, k[j].x, s[i][j]);
          s[i][j] = fmaf(q[i].y, k[j].y, s[i][j]);
          s[i][j] = fmaf(q[i].z, k[j].z, s[i][j]);
          s[i][j] = fmaf(q[i].w, k[j].w, s[i][j]);
        }
    }
    // online softmax
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + tx + 16 * j < S_total) ? s[i][j] * p.scale_log2e : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 8; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float mn = fmaxf(m[i], mx);
      const float alpha = exp2f(m[i] - mn);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float pv = exp2f(s[i][j] - mn);
        rs += pv;
        Ps[(ty + 16 * i) * LDP + tx + 16 * j] = pv;
      }
#pragma unroll
      for (int off = 8; off; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l[i] = l[i] * alpha + rs;
      m[i] = mn;
#pragma unroll
      for (int j = 0; j < NV; ++j) o[i][j] *= alpha;
    }
    __syncthreads();
    // O += P V : rows ty + 16 i, cols tx*NV .. +NV-1
#pragma unroll 4
    for (int k = 0; k < BK; k += 4) {
      float4 pr[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pr[i] = *reinterpret_cast<const float4*>(Ps + (ty + 16 * i) * LDP + k);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        float vv[NV];
        if constexpr (NV == 2) {
          float2 t = *reinterpret_cast<const float2*>(Vs + (k + kk) * DH + tx * 2);
          vv[0] = t.x; vv[1] = t.y;
        } else {
          float4 t = *reinterpret_cast<const float4*>(Vs + (k + kk) * DH + tx * 4);
          vv[0] = t.x; vv[1] = t.y; vv[2] = t.z; vv[3] = t.w;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float pk = kk == 0 ? pr[i].x : kk == 1 ? pr[i].y : kk == 2 ? pr[i].z : pr[i].w;
#pragma unroll
          for (int j = 0; j < NV; ++j) o[i][j] = fmaf(pk, vv[j], o[i][j]);
        }
      }
    }
  }
  float* Og = p.O + (long long)r * p.o_sample_stride + h * DH;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = q0 + ty + 16 * i;
    if (row >= p.T) continue;
    const float inv = 1.f / l[i];
    if (p.Op) {
      __nv_bfloat16* dst = p.Op + (long long)r * p.o_sample_stride + (long long)row * p.o_ld + h * DH + tx * NV;
      for (int t = 0; t < p.op_terms; ++t) {
        __nv_bfloat16 pl[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          float val = o[i][j] * inv;
          __nv_bfloat16 sp[3];
          umma::split_bf16<3>(val, sp);
          pl[j] = sp[t];
        }
        if constexpr (NV == 2) *reinterpret_cast<uint32_t*>(dst + t * p.op_plane_stride) = *reinterpret_cast<const uint32_t*>(pl);
        else *reinterpret_cast<uint2*>(dst + t * p.op_plane_stride) = *reinterpret_cast<const uint2*>(pl);
      }
      continue;
    }
    if constexpr (NV == 2) {
      *reinterpret_cast<float2*>(Og + (long long)row * p.o_ld + tx * 2) = make_float2(o[i][0] * inv, o[i][1] * inv);
    } else {
      *reinterpret_cast<float4*>(Og + (long long)row * p.o_ld + tx * 4) =
          make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
    }
  }
}

// opt in to >48 KB dynamic shared memory once per process (never inside a stream capture)
inline int init_attn_simt() {
  A2P_CUDA(cudaFuncSetAttribute(attn_simt_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(sizeof(float) * (64 * 36 * 2 + 64 * 32 + 64 * 68))));
  A2P_CUDA(cudaFuncSetAttribute(attn_simt_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(sizeof(float) * (64 * 68 * 2 + 64 * 64 + 64 * 68))));
  return 0;
}

inline int launch_attn_simt(const AttnParams& p, int dh, cudaStream_t st) {
  dim3 grid(ceil_div(p.T, 64), p.H, p.R);
  if (dh == 32) {
    constexpr int DH = 32;
    size_t sm = sizeof(float) * (64 * (DH + 4) * 2 + 64 * DH + 64 * 68);
    attn_simt_kernel<32><<<grid, 256, sm, st>>>(p);
  } else if (dh == 64) {
    constexpr int DH = 64;
    size_t sm = sizeof(float) * (64 * (DH + 4) * 2 + 64 * DH + 64 * 68);
    attn_simt_kernel<64><<<grid, 256, sm, st>>>(p);
  } else {
    A2P_FAIL("attention: head dim %d unsupported (32 or 64)", dh);
  }
  A2P_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace a2p
