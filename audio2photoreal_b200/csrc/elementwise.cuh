// HBM-bound kernels of the path: LayerNorm(+RoPE), timestep embedding, layout moves, and the K3
// sampler epilogue.  All are one-pass, 128-bit vectorised where the layout allows.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace a2p {

// ------------------------------------------------------------------ RoPE cos/sin table
// tab[pos][i] = (cos(pos * f_i), sin(pos * f_i)),  i < D/2.  Angles are fp32 products like the reference
// (rotary_embedding_torch.py:133: einsum of fp32 positions and fp32 freqs), accurate cosf/sinf.
__global__ void rope_table_kernel(const float* __restrict__ freqs, float2* __restrict__ tab, int max_pos, int half) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= max_pos * half) return;
  int pos = idx / half, i = idx - pos * half;
  float ang = __fmul_rn((float)pos, freqs[i]);
  tab[idx] = make_float2(cosf(ang), sinf(ang));
}

// ------------------------------------------------------------------ LayerNorm (+ optional RoPE)
// One warp per row of D (256 or 512) floats.  Writes h = LN(x)*w+b (if out_h) and rot(h) (if out_r),
// where rot is the full-width interleaved-pair rotation at position pos_base + (row % pos_mod)
// (rotary_embedding_torch.py:46-66).  eps = 1e-5, biased variance (nn.LayerNorm).
template <int D>
__global__ void __launch_bounds__(256) ln_rope_kernel(const float* __restrict__ x, long long ldx,
                                                      const float* __restrict__ w, const float* __restrict__ b,
                                                      float* __restrict__ out_h, float* __restrict__ out_r, long long ldo,
                                                      const float2* __restrict__ tab, int half, int pos_mod,
                                                      int pos_base, int rows) {
  constexpr int PER = D / 32;  // 8 or 16 floats per lane, as float4 chunks strided by 32 lanes
  constexpr int NV = PER / 4;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* xr = x + (long long)warp * ldx;
  float v[PER];
#pragma unroll
  for (int c = 0; c < NV; ++c) *reinterpret_cast<float4*>(v + 4 * c) = *reinterpret_cast<const float4*>(xr + (c * 32 + lane) * 4);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { float d = v[i] - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / D + 1e-5f);
  const int pos = pos_base + (pos_mod > 0 ? warp % pos_mod : 0);
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int col = (c * 32 + lane) * 4;
    float4 ww = *reinterpret_cast<const float4*>(w + col), bb = *reinterpret_cast<const float4*>(b + col);
    float h0 = (v[4 * c + 0] - mean) * rstd * ww.x + bb.x;
    float h1 = (v[4 * c + 1] - mean) * rstd * ww.y + bb.y;
    float h2 = (v[4 * c + 2] - mean) * rstd * ww.z + bb.z;
    float h3 = (v[4 * c + 3] - mean) * rstd * ww.w + bb.w;
    if (out_h) *reinterpret_cast<float4*>(out_h + (long long)warp * ldo + col) = make_float4(h0, h1, h2, h3);
    if (out_r) {
      float2 cs0 = tab[(long long)pos * half + col / 2], cs1 = tab[(long long)pos * half + col / 2 + 1];
      float r0 = h0 * cs0.x - h1 * cs0.y, r1 = h1 * cs0.x + h0 * cs0.y;
      float r2 = h2 * cs1.x - h3 * cs1.y, r3 = h3 * cs1.x + h2 * cs1.y;
      *reinterpret_cast<float4*>(out_r + (long long)warp * ldo + col) = make_float4(r0, r1, r2, r3);
    }
  }
}

// Same LayerNorm(+RoPE) but the outputs are written as TERMS split-bf16 planes [TERMS][rows][D] -- the A operand
// of the tensor-core GEMMs (LN / RoPE stay in fp32 registers; only the GEMM input is split).
template <int D, int TERMS>
__global__ void __launch_bounds__(256) ln_rope_planes_kernel(const float* __restrict__ x, long long ldx,
                                                             const float* __restrict__ w, const float* __restrict__ b,
                                                             __nv_bfloat16* __restrict__ out_h, __nv_bfloat16* __restrict__ out_r,
                                                             long long plane_stride, const float2* __restrict__ tab, int half,
                                                             int pos_mod, int pos_base, int rows) {
  constexpr int PER = D / 32;
  constexpr int NV = PER / 4;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  pdl_wait();
  if (warp >= rows) return;
  const float* xr = x + (long long)warp * ldx;
  float v[PER];
#pragma unroll
  for (int c = 0; c < NV; ++c) *reinterpret_cast<float4*>(v + 4 * c) = *reinterpret_cast<const float4*>(xr + (c * 32 + lane) * 4);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i];
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) { float d = v[i] - mean; q += d * d; }
#pragma unroll
  for (int o = 16; o; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / D + 1e-5f);
  const int pos = pos_base + (pos_mod > 0 ? warp % pos_mod : 0);
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int col = (c * 32 + lane) * 4;
    float4 ww = *reinterpret_cast<const float4*>(w + col), bb = *reinterpret_cast<const float4*>(b + col);
    float h[4];
    h[0] = (v[4 * c + 0] - mean) * rstd * ww.x + bb.x;
    h[1] = (v[4 * c + 1] - mean) * rstd * ww.y + bb.y;
    h[2] = (v[4 * c + 2] - mean) * rstd * ww.z + bb.z;
    h[3] = (v[4 * c + 3] - mean) * rstd * ww.w + bb.w;
    auto emit = [&](const float* val, __nv_bfloat16* dst) {
      __nv_bfloat16 pl[TERMS][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        __nv_bfloat16 sp[TERMS];
        umma::split_bf16<TERMS>(val[j], sp);
#pragma unroll
        for (int i = 0; i < TERMS; ++i) pl[i][j] = sp[i];
      }
#pragma unroll
      for (int i = 0; i < TERMS; ++i)
        *reinterpret_cast<uint2*>(dst + i * plane_stride + (long long)warp * D + col) = *reinterpret_cast<const uint2*>(pl[i]);
    };
    if (out_h) emit(h, out_h);
    if (out_r) {
      float2 cs0 = tab[(long long)pos * half + col / 2], cs1 = tab[(long long)pos * half + col / 2 + 1];
      float r[4];
      r[0] = h[0] * cs0.x - h[1] * cs0.y; r[1] = h[1] * cs0.x + h[0] * cs0.y;
      r[2] = h[2] * cs1.x - h[3] * cs1.y; r[3] = h[3] * cs1.x + h[2] * cs1.y;
      emit(r, out_r);
    }
  }
}

inline int launch_ln_rope_planes(int D, int terms, const float* x, long long ldx, const float* w, const float* b,
                                 __nv_bfloat16* out_h, __nv_bfloat16* out_r, long long plane_stride, const float2* tab,
                                 int pos_mod, int pos_base, int rows, cudaStream_t st) {
  int blocks = ceil_div(rows, 8);
#define A2P_LNP(DD, TT) A2P_CUDA(launch_pdl(ln_rope_planes_kernel<DD, TT>, dim3(blocks), dim3(256), (size_t)0, st, x, ldx, w, b, out_h, out_r, plane_stride, tab, DD / 2, pos_mod, pos_base, rows))
  if (D == 256 && terms == 1) A2P_LNP(256, 1);
  else if (D == 256 && terms == 2) A2P_LNP(256, 2);
  else if (D == 256 && terms == 3) A2P_LNP(256, 3);
  else if (D == 512 && terms == 1) A2P_LNP(512, 1);
  else if (D == 512 && terms == 2) A2P_LNP(512, 2);
  else if (D == 512 && terms == 3) A2P_LNP(512, 3);
  else A2P_FAIL("ln_rope_planes: D=%d terms=%d unsupported", D, terms);
#undef A2P_LNP
  A2P_CUDA(cudaGetLastError());
  return 0;
}

inline int launch_ln_rope(int D, const float* x, long long ldx, const float* w, const float* b, float* out_h,
                          float* out_r, long long ldo, const float2* tab, int pos_mod, int pos_base, int rows,
                          cudaStream_t st) {
  int blocks = ceil_div(rows, 8);
  if (D == 256)
    ln_rope_kernel<256><<<blocks, 256, 0, st>>>(x, ldx, w, b, out_h, out_r, ldo, tab, D / 2, pos_mod, pos_base, rows);
  else if (D == 512)
    ln_rope_kernel<512><<<blocks, 256, 0, st>>>(x, ldx, w, b, out_h, out_r, ldo, tab, D / 2, pos_mod, pos_base, rows);
  else
    A2P_FAIL("ln_rope: D=%d unsupported", D);
  A2P_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ timestep embedding
// e[r] = [sin(t*f_k) | cos(t*f_k)], f_k = exp(-k*ln(1e4)/(half-1))  (model/utils.py:67-79).
// `ts` is either the per-row [B] int64 array (forward API) or, inside the graph-captured loop, the
// [n_steps] table indexed by the device-side step counter.
// The frequency table f_k is built on the host with the reference's own torch expression so that the
// fp32 angle t*f_k (up to ~1e3 rad) is bit-identical; only sinf/cosf run here.
__global__ void time_embed_kernel(const long long* __restrict__ ts, const int* __restrict__ step_counter, int B,
                                  int rows, int D, const float* __restrict__ freqs, float* __restrict__ e) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  int half = D / 2;
  if (idx >= rows * half) return;
  int r = idx / half, k = idx - r * half;
  long long t = step_counter ? ts[*step_counter] : ts[r % B];
  float a = __fmul_rn((float)t, freqs[k]);
  e[(long long)r * D + k] = sinf(a);
  e[(long long)r * D + half + k] = cosf(a);
}

// ------------------------------------------------------------------ [B,C,1,T] -> [B,T,C] (tile transpose)
__global__ void bct_to_btc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int T) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + (long long)b * C * T;
  float* d = dst + (long long)b * C * T;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, t = t0 + threadIdx.x;
    if (c < C && t < T) tile[i][threadIdx.x] = s[(long long)c * T + t];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int t = t0 + i, c = c0 + threadIdx.x;
    if (c < C && t < T) d[(long long)t * C + c] = tile[threadIdx.x][i];
  }
}

// copy rows [n] floats (duplicate the input-projected x for the uncond half, pad TCN input, ...)
__global__ void copy_rows_kernel(const float* __restrict__ src, long long src_ld, long long src_sample_stride,
                                 float* __restrict__ dst, long long dst_ld, long long dst_sample_stride, int rows_per_sample,
                                 int n4, int samples) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)samples * rows_per_sample * n4;
  if (idx >= total) return;
  int c = idx % n4;
  long long rr = idx / n4;
  int r = rr % rows_per_sample;
  int s = rr / rows_per_sample;
  reinterpret_cast<float4*>(dst + s * dst_sample_stride + (long long)r * dst_ld)[c] =
      reinterpret_cast<const float4*>(src + s * src_sample_stride + (long long)r * src_ld)[c];
}

__global__ void fill_zero_kernel(float4* p, long long n4) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------ K3: fused sampler epilogue
// Reads the two denoiser outputs [B,T,C], mixes them (cfg_sampler.py:33), transposes to [B,C,1,T]
// (gaussian_diffusion.py:312-313) and applies the DDIM (:699-718) or ancestral (:243-246,:471-476)
// update.  Arithmetic mirrors the reference's op ORDER with explicit non-contracted fp32 ops, so with
// the same x0 inputs the result is bit-identical to the PyTorch fp32 path.
struct K3Params {
  const float* x_t; const float* x0c; const float* x0u; const float* scale;
  const float* coeffs;        // [n_steps, 8] or one row
  const int* step_counter;    // null -> row 0 of coeffs
  const float* noise; long long noise_step_stride;   // noise + loop_iter * stride
  int n_steps;
  float* x_prev; float* pred;
  int B, C, T, kind, clip;
  long long x0_sample_stride;  // floats between consecutive samples of x0c / x0u (rows are C apart)
  int* step_counter_dec;      // if set, thread 0 of block 0 decrements after use (graph loop)
  int rng;                    // 1: noise = counter-based N(0,1) (Philox4x32-10 + Box-Muller) instead of a tape
  unsigned long long seed;    // Philox key
  long long rng_row0;         // global index of batch row 0 (sharded runs draw the rows they own from the same stream)
};

// Philox4x32-10 (Salmon et al., SC'11; the generator behind curand / torch CUDA).  counter = (element lo, element hi,
// loop iteration, 0), key = seed.  Statistically -- not bitwise -- equivalent to the reference's th.randn_like.
__host__ __device__ inline void philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
                                              unsigned out[4]) {
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0, p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0, n1 = (unsigned)p1, n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1, n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long elem, unsigned iter) {
  unsigned r[4];
  philox4x32_10((unsigned)elem, (unsigned)(elem >> 32), iter, 0u, (unsigned)seed, (unsigned)(seed >> 32), r);
  const float u1 = ((float)r[0] + 1.0f) * 2.3283064365386963e-10f;   // (0, 1]
  const float u2 = (float)r[1] * 2.3283064365386963e-10f;            // [0, 1)
  return sqrtf(-2.0f * logf(u1)) * cospif(2.0f * u2);
}

__global__ void __launch_bounds__(256) k3_sampler_kernel(K3Params p) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int step = p.step_counter ? *p.step_counter : 0;
  const float* co = p.coeffs + (long long)step * 8;
  const float g = (p.x0u && p.scale) ? p.scale[b] : 1.f;
  // phase 1: coalesced read of [T,C] tiles, CFG mix, stash transposed
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int t = t0 + i, c = c0 + threadIdx.x;
    if (t < p.T && c < p.C) {
      long long o = (long long)b * p.x0_sample_stride + (long long)t * p.C + c;
      float xc = p.x0c[o];
      float x0 = xc;
      if (p.x0u) {
        float xu = p.x0u[o];
        x0 = __fadd_rn(xu, __fmul_rn(g, __fsub_rn(xc, xu)));
      }
      if (p.clip) x0 = fminf(fmaxf(x0, -1.f), 1.f);
      tile[i][threadIdx.x] = x0;
    }
  }
  __syncthreads();
  const float* noise = nullptr;
  if (p.noise) {
    long long it = p.step_counter ? (long long)(p.n_steps - 1 - step) : 0;
    noise = p.noise + it * p.noise_step_stride;
  }
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, t = t0 + threadIdx.x;
    if (t < p.T && c < p.C) {
      long long o = ((long long)b * p.C + c) * p.T + t;
      float x0 = tile[threadIdx.x][i];
      float xt = p.x_t[o];
      float nz = noise ? noise[o] : 0.f;
      if (p.rng) {
        const long long it = p.step_counter ? (long long)(p.n_steps - 1 - step) : p.noise_step_stride;   // single step: iteration passed in noise_step_stride
        nz = philox_normal(p.seed, (unsigned long long)(((p.rng_row0 + b) * p.C + c) * (long long)p.T + t), (unsigned)it);
      }
      float out;
      if (p.kind == 0) {  // DDIM
        float eps = __fdiv_rn(__fsub_rn(__fmul_rn(co[0], xt), x0), co[1]);
        float mean = __fadd_rn(__fmul_rn(x0, co[2]), __fmul_rn(co[3], eps));
        out = (noise || p.rng) ? __fadd_rn(mean, __fmul_rn(co[4], nz)) : mean;
      } else {  // ancestral
        float mean = __fadd_rn(__fmul_rn(co[5], x0), __fmul_rn(co[6], xt));
        out = __fadd_rn(mean, __fmul_rn(co[7], nz));
      }
      p.pred[o] = x0;
      p.x_prev[o] = out;
    }
  }
}

// decrement the device-side step counter (own 1-thread kernel so that every block of K3 saw the old value)
__global__ void step_dec_kernel(int* counter) { *counter -= 1; }
__global__ void step_set_kernel(int* counter, int v) { *counter = v; }

}  // namespace a2p
