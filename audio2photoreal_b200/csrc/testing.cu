// Test / measurement hooks (include/a2p_b200_testing.h): built into liba2p_b200_testing.so, NOT into the product
// library.  They drive the same header-only kernels (umma_*.cuh, sgemm.cuh, attention_simt.cuh, elementwise.cuh) on
// caller-provided fp32 tensors so that tests/ can check every kernel in isolation against fp64 / the exact-fp32 arm.
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string>
#include <vector>

#include "../../include/a2p_b200.h"
#include "attention_simt.cuh"
#include "common.cuh"
#include "elementwise.cuh"
#include "sgemm.cuh"
#include "umma_gemm.cuh"
#include "umma_attention.cuh"
#include "umma_chain.cuh"
#include "umma_attention2.cuh"
#include "umma_attention_short.cuh"
#include "umma_microbench.cuh"
#include "../../include/a2p_b200_testing.h"

using namespace a2p;

extern "C" {

const char* a2p_test_last_error(void) { return a2p::last_error().c_str(); }

// ------------------------------------------------------------------ testing hooks (include/a2p_b200_testing.h)
size_t a2p_test_tc_gemm_scratch_bytes(int M, int N, int K, int taps) {
  return ((size_t)3 * M * K + (size_t)3 * taps * N * K) * 2 + 1024;
}

int a2p_test_tc_gemm(int terms, int M, int N, int K, int taps, int dil, const float* A, const float* W, const float* bias,
                     float* C, void* scratch, size_t scratch_bytes, int iters, float* ms_out, void* stream) {
  if (scratch_bytes < a2p_test_tc_gemm_scratch_bytes(M, N, K, taps)) A2P_FAIL("test_tc_gemm: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  A2P_TRY(init_umma_gemm());
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  __nv_bfloat16* Ap = (__nv_bfloat16*)scratch;
  __nv_bfloat16* Wp = Ap + align_up((size_t)3 * M * K, 64);
  A2P_TRY(launch_split_planes(terms, A, K, Ap, (long long)M * K, M, K, 1.f, st));
  A2P_TRY(launch_split_planes(terms, W, K, Wp, (long long)taps * N * K, (long long)taps * N, K, 1.f, st));
  TcOperands o{Ap, K, (long long)M * K, Wp, K, (long long)N * K};
  TcGemmParams p{};
  p.M = M; p.N = N; p.K = K; p.taps = taps; p.dil = dil; p.bias = bias; p.C = C; p.ldc = N; p.out_scale = 1.f;
  A2P_TRY(launch_umma_gemm(terms, o, p, TC_F32, sms, st));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters; ++i) A2P_TRY(launch_umma_gemm(terms, o, p, TC_F32, sms, st));
  cudaEventRecord(e1, st);
  A2P_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = iters > 0 ? ms / iters : 0.f;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

void a2p_test_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out4) {
  unsigned o[4];
  philox4x32_10(c0, c1, c2, c3, k0, k1, o);   // the same inline function the K3 kernel calls (host compilation)
  for (int i = 0; i < 4; ++i) out4[i] = o[i];
}

int a2p_test_mma_rate(int N, int a_from_tmem, int n_mma, long long* cycles_out_dev, void* stream) {
  return launch_mma_rate(N, a_from_tmem, n_mma, cycles_out_dev, (cudaStream_t)stream);
}

int a2p_test_tmem_ldst_rate(int store, int n_ops, int n_warps, long long* cycles_out_dev, void* stream) {
  return launch_tmem_ldst_rate(store, n_ops, n_warps, cycles_out_dev, (cudaStream_t)stream);
}

size_t a2p_test_tc_attention_scratch_bytes(int R, int T, int D, int S, int n_extra) {
  const size_t Sp = align_up((size_t)S, 8), Xp = 8;
  return (align_up((size_t)3 * R * T * D, 512) + align_up((size_t)3 * R * S * D, 512) + align_up((size_t)3 * D * R * Sp, 512) +
          align_up((size_t)3 * R * 8 * D, 512) + align_up((size_t)3 * D * R * Xp, 512)) * 2 + 8192 + 64 * 32 * 8 +
         attn2_split_scratch_floats() * 4 + attn2_split_counter_ints() * 4 + 1024;
}

int a2p_test_tc_attention(int terms, int R, int T, int D, int dh, int S, int n_extra, const float* Q, const float* K,
                          const float* V, const float* Kx, const float* Vx, float* O, void* scratch, size_t scratch_bytes,
                          int iters, float* ms_out, void* stream) {
  if (scratch_bytes < a2p_test_tc_attention_scratch_bytes(R, T, D, S, n_extra)) A2P_FAIL("test_tc_attention: scratch too small");
  if (n_extra > 8) A2P_FAIL("test_tc_attention: n_extra <= 8");
  cudaStream_t st = (cudaStream_t)stream;
  int variant = 0;   // terms 20 / 21: second-generation kernel with P planes in shared / tensor memory (two planes)
  if (terms >= 20) { variant = terms - 19; terms = 2; }
  A2P_TRY(init_umma_attn());
  A2P_TRY(init_umma_attn2());
  A2P_TRY(init_umma_attn_short());
  const long long Sp = (long long)align_up((size_t)S, 8), Xp = 8;
  __nv_bfloat16* Qp = (__nv_bfloat16*)scratch;
  __nv_bfloat16* Kp = Qp + align_up((size_t)3 * R * T * D, 512);
  __nv_bfloat16* Vt = Kp + align_up((size_t)3 * R * S * D, 512);
  __nv_bfloat16* Kxp = Vt + align_up((size_t)3 * D * R * Sp, 512);
  __nv_bfloat16* Vxt = Kxp + align_up((size_t)3 * R * 8 * D, 512);
  A2P_CUDA(cudaMemsetAsync(scratch, 0, scratch_bytes, st));
  const float sc = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  A2P_TRY(launch_split_planes(terms, Q, D, Qp, (long long)R * T * D, (long long)R * T, D, sc, st));
  A2P_TRY(launch_split_planes(terms, K, D, Kp, (long long)R * S * D, (long long)R * S, D, 1.f, st));
  A2P_TRY(launch_transpose_split(terms, V, D, Vt, (long long)D * R * Sp, (long long)R * Sp, R * S, D, S, Sp, 1.f, st));
  if (n_extra > 0) {
    for (int r = 0; r < R; ++r)
      A2P_TRY(launch_split_planes(terms, Kx + (size_t)r * n_extra * D, D, Kxp + (size_t)r * 8 * D, (long long)R * 8 * D, n_extra, D, 1.f, st));
    A2P_TRY(launch_transpose_split(terms, Vx, D, Vxt, (long long)D * R * Xp, (long long)R * Xp, R * n_extra, D, n_extra, Xp, 1.f, st));
  }
  TcAttnOperands o{};
  o.Q = Qp; o.q_rows = (long long)R * T; o.q_ld = D; o.q_plane_stride = (long long)R * T * D;
  o.K[0] = Kp; o.k_rows[0] = (long long)R * S; o.k_ld[0] = D; o.k_plane_stride[0] = (long long)R * S * D; o.K[1] = nullptr;
  o.Vt[0] = Vt; o.vt_cols[0] = (long long)R * Sp; o.vt_ld[0] = (long long)R * Sp; o.vt_plane_stride[0] = (long long)D * R * Sp; o.Vt[1] = nullptr;
  o.vt_rows = D;
  if (n_extra > 0) {
    o.Kx = Kxp; o.kx_rows = (long long)R * 8; o.kx_ld = D; o.kx_plane_stride = (long long)R * 8 * D;
    o.Vx = Vxt; o.vx_rows = D; o.vx_cols = (long long)R * Xp; o.vx_ld = (long long)R * Xp; o.vx_plane_stride = (long long)D * R * Xp;
  }
  TcAttnParams p{};
  p.T = T; p.R = R; p.D = D; p.dh = dh; p.rows_per_branch = R; p.q_col0 = 0; p.k_col0 = 0; p.n_keys = S; p.n_extra = n_extra;
  p.k_row_stride[0] = S; p.v_col_stride[0] = Sp; p.kx_col0 = 0; p.kx_row_stride = 8; p.vx_row0 = 0; p.vx_col_stride = (int)Xp;
  p.O = O; p.o_ld = D; p.Op = nullptr;
  p.skew_ns = getenv("A2P_ATTN_SKEW_NS") ? atoi(getenv("A2P_ATTN_SKEW_NS")) : 0;
  p.trace = (iters < 0) ? reinterpret_cast<long long*>(Vxt + align_up((size_t)3 * D * R * Xp, 512)) : nullptr;   // iters < 0: trace mode
  {  // split-KV scratch behind the trace area (the whole scratch buffer was zeroed above, counters included)
    char* tail = reinterpret_cast<char*>(Vxt + align_up((size_t)3 * D * R * Xp, 512)) + 64 * 32 * 8;
    tail += 512 - (reinterpret_cast<uintptr_t>(tail) & 255);
    p.split_scratch = reinterpret_cast<float*>(tail);
    p.split_counters = reinterpret_cast<int*>(tail + attn2_split_scratch_floats() * 4);
  }
  auto launch = [&]() -> int {   // terms 24 (variant 5): the short-key-set kernel (umma_attention_short.cuh)
    if (variant == 5) return launch_umma_attn_short(o, p, st);
    return variant ? launch_umma_attn2(variant, o, p, st) : launch_umma_attn(terms, o, p, st);
  };
  A2P_TRY(launch());
  if (iters < 0) {
    A2P_CUDA(cudaStreamSynchronize(st));
    A2P_CUDA(cudaMemcpy(O, p.trace, 64 * 32 * sizeof(long long), cudaMemcpyDeviceToDevice));   // trace returned in the O buffer
    return 0;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters; ++i) A2P_TRY(launch());
  cudaEventRecord(e1, st);
  A2P_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = iters > 0 ? ms / iters : 0.f;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

void a2p_test_attn2_set_persist(int on) { attn2_persist_override() = on < 0 ? -1 : (on != 0); }

void a2p_test_chain_set_mode(int cl) { chain_mode_override() = (cl == 1 || cl == 2) ? cl : 0; }
void a2p_test_chain_set_nsplit(int n) { chain_nsplit_override() = n > 0 ? n : 0; }
int a2p_test_chain_nsplit_policy(int tiles, int n_acc, int concurrent, int K0) { return chain_nsplit_for(tiles, n_acc, concurrent, K0); }

size_t a2p_test_chain_scratch_bytes(int M, int K0, int N1, int T) {
  return ((size_t)2 * align_up((size_t)M, 128) * K0 + (size_t)2 * 256 * K0 + (size_t)2 * N1 * 256 + (size_t)2 * 256 * 256) * 2 +
         (size_t)T * 128 * 8 + (size_t)(T + 128) * 128 * 8 + 4096 + 1024;
}

int a2p_test_chain(int M, int T, int K0, int N1, int film_mode, int ln_mode, int rope, int gelu, int vjob, float out_scale,
                   int scale_ncols, const float* A0, const float* W0, const float* bias0, const float* film, float* x,
                   const float* ln_w, const float* ln_b, const float* rope_freqs, const float* W1, const float* bias1,
                   const float* W2, const float* bias2, void* Cp_out, void* Vt_out, void* scratch, size_t scratch_bytes,
                   int iters, float* ms_out, void* stream) {
  if (scratch_bytes < a2p_test_chain_scratch_bytes(M, K0, N1, T)) A2P_FAIL("test_chain: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  A2P_TRY(init_umma_chain());
  __nv_bfloat16* A0p = (__nv_bfloat16*)scratch;
  __nv_bfloat16* W0p = A0p + (size_t)2 * align_up((size_t)M, 128) * K0;
  __nv_bfloat16* W1p = W0p + (size_t)2 * 256 * K0;
  __nv_bfloat16* W2p = W1p + (size_t)2 * N1 * 256;
  float2* tab = reinterpret_cast<float2*>(reinterpret_cast<char*>(W2p + (size_t)2 * 256 * 256) + 1024 -
                                          (reinterpret_cast<uintptr_t>(W2p + (size_t)2 * 256 * 256) & 1023));
  A2P_TRY(launch_split_planes(2, A0, K0, A0p, (long long)M * K0, M, K0, 1.f, st));
  A2P_TRY(launch_split_planes(2, W0, K0, W0p, (long long)256 * K0, 256, K0, 1.f, st));
  A2P_TRY(launch_split_planes(2, W1, 256, W1p, (long long)N1 * 256, N1, 256, 1.f, st));
  if (vjob) A2P_TRY(launch_split_planes(2, W2, 256, W2p, (long long)256 * 256, 256, 256, 1.f, st));
  float2* ext = tab + (size_t)T * 128;
  rope_table_kernel<<<ceil_div(T * 128, 256), 256, 0, st>>>(rope_freqs, tab, T, 128);
  rope_ext_kernel<<<ceil_div((T + 128) * 128, 256), 256, 0, st>>>(tab, ext, T, 128);
  A2P_CUDA(cudaGetLastError());
  const long long M8 = (long long)align_up((size_t)M, 8);
  ChainOperands o{};
  o.A0 = A0p; o.a0_rows = M; o.a0_ld = K0; o.a0_plane_stride = (long long)M * K0;
  o.W0 = W0p; o.w0_plane_stride = (long long)256 * K0;
  o.W1 = W1p; o.w1_plane_stride = (long long)N1 * 256;
  o.W2 = vjob ? W2p : nullptr; o.w2_plane_stride = (long long)256 * 256;
  o.x = x; o.rope_ext = reinterpret_cast<const float*>(ext); o.rope_ext_rows = T + 128;
  ChainParams cp{};
  cp.M = M; cp.T = T; cp.K0 = K0; cp.bias0 = bias0; cp.film_mode = film_mode; cp.film = film; cp.film_ld = 512;
  cp.film_scale_off = 0; cp.film_shift_off = 256; cp.ln_mode = ln_mode; cp.ln_w = ln_w; cp.ln_b = ln_b;
  cp.rope = rope; cp.N1 = N1; cp.bias1 = bias1; cp.out_scale = out_scale; cp.scale_ncols = scale_ncols;
  cp.gelu = gelu; cp.Cp = (__nv_bfloat16*)Cp_out; cp.cp_plane_stride = (long long)M * N1; cp.ldcp = N1;
  cp.vjob = vjob; cp.bias2 = bias2; cp.Vt = (__nv_bfloat16*)Vt_out; cp.vt_plane_stride = 256 * M8; cp.ldvt = M8;
  if (getenv("A2P_CHAIN_TRACE")) {   // clock64 timeline of CTA 0 (diagnostics): stored behind the RoPE tables in the scratch buffer
    cp.trace = reinterpret_cast<long long*>(ext + (size_t)(T + 128) * 128);
    A2P_CUDA(cudaMemsetAsync(cp.trace, 0, 64 * sizeof(long long), st));
  }
  // N split (a2p_test_chain_set_nsplit): the parts of a tile read x and write the updated stream to a second buffer, which
  // is copied back so that the caller sees the same in-place contract
  float* xtmp = nullptr;
  {
    const int n_acc = ceil_div(N1, 128) + (vjob ? 2 : 0);
    cp.nsplit = chain_nsplit_override() > n_acc ? n_acc : chain_nsplit_override();
    if (cp.nsplit > 1) {
      A2P_CUDA(cudaMalloc(&xtmp, (size_t)M * 256 * sizeof(float)));
      o.x_out = xtmp;
    }
  }
  A2P_TRY(launch_umma_chain(o, cp, st));
  if (xtmp) A2P_CUDA(cudaMemcpyAsync(x, xtmp, (size_t)M * 256 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (cp.trace) {
    long long tr[64];
    A2P_CUDA(cudaStreamSynchronize(st));
    A2P_CUDA(cudaMemcpy(tr, cp.trace, sizeof(tr), cudaMemcpyDeviceToHost));
    fprintf(stderr, "a2p chain trace (cycles since start):");
    for (int i = 0; i < 64; ++i) if (tr[i]) fprintf(stderr, " [%d]=%lld", i, tr[i] - tr[0]);
    fprintf(stderr, "\n");
    cp.trace = nullptr;
  }
  if (iters <= 0) { A2P_CUDA(cudaStreamSynchronize(st)); if (xtmp) cudaFree(xtmp); return 0; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters; ++i) A2P_TRY(launch_umma_chain(o, cp, st));
  cudaEventRecord(e1, st);
  A2P_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = ms / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (xtmp) cudaFree(xtmp);
  return 0;
}

int a2p_test_simt_attention(int R, int T, int D, int dh, int S, int n_extra, const float* Q, const float* K, const float* V,
                            const float* Kx, const float* Vx, float* O, int iters, float* ms_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  A2P_TRY(init_attn_simt());
  AttnParams a{};
  a.Q = Q; a.q_ld = D; a.q_sample_stride = (long long)T * D;
  a.K.base[0] = K; a.K.stride[0] = (long long)S * D; a.K.rows_per_branch = R; a.V = a.K; a.V.base[0] = V;
  a.kv_ld = D; a.S_main = S; a.Kx = n_extra ? Kx : nullptr; a.Vx = n_extra ? Vx : nullptr; a.x_ld = D;
  a.x_sample_stride = (long long)n_extra * D; a.S_extra = n_extra;
  a.O = O; a.o_ld = D; a.o_sample_stride = (long long)T * D; a.T = T; a.H = D / dh; a.R = R;
  a.scale_log2e = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  A2P_TRY(launch_attn_simt(a, dh, st));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters; ++i) A2P_TRY(launch_attn_simt(a, dh, st));
  cudaEventRecord(e1, st);
  A2P_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = iters > 0 ? ms / iters : 0.f;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

int a2p_test_sgemm(int M, int N, int K, int taps, int dil, const float* A, const float* W, const float* bias, float* C,
                   int iters, float* ms_out, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  GemmParams p{};
  p.A = A; p.lda = K; p.W = W; p.ldw = (long long)taps * K; p.bias = bias; p.C = C; p.ldc = N; p.M = M; p.N = N; p.K = taps * K;
  p.taps = taps; p.dil = dil; p.Kc = K; p.epi = EPI_BIAS;
  A2P_TRY(launch_sgemm(p, st));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int i = 0; i < iters; ++i) A2P_TRY(launch_sgemm(p, st));
  cudaEventRecord(e1, st);
  A2P_CUDA(cudaStreamSynchronize(st));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (ms_out) *ms_out = iters > 0 ? ms / iters : 0.f;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return 0;
}

}  // extern "C"
