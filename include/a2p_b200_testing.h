/* Test / measurement hooks, built into liba2p_b200_testing.so (NOT part of the drop-in ABI and NOT in the product
 * library liba2p_b200.so; used by tests/ and scripts/). */
#ifndef A2P_B200_TESTING_H
#define A2P_B200_TESTING_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* message of the last failed a2p_test_* call of this thread */
const char* a2p_test_last_error(void);

/* C[M,N] = A[M,K] * W[taps][N][K]^T (+ bias) with the split-bf16 tcgen05 GEMM (terms 1..3); fp32 in/out.
 * scratch holds the bf16 planes (size from a2p_test_tc_gemm_scratch_bytes).  Runs `iters` timed launches
 * after one warm-up and returns the average kernel time (CUDA events) in *ms_out.  Synchronises. */
size_t a2p_test_tc_gemm_scratch_bytes(int M, int N, int K, int taps);
int a2p_test_tc_gemm(int terms, int M, int N, int K, int taps, int dil, const float* A, const float* W, const float* bias,
                     float* C, void* scratch, size_t scratch_bytes, int iters, float* ms_out, void* stream);
/* same product with the exact-fp32 FFMA GEMM (W as [N][taps*K]) */
int a2p_test_sgemm(int M, int N, int K, int taps, int dil, const float* A, const float* W, const float* bias, float* C,
                   int iters, float* ms_out, void* stream);

/* softmax(QK^T/sqrt(dh))V with the split-bf16 tcgen05 attention kernel.  fp32 inputs Q [R,T,D], K,V [R,S,D],
 * optional extra keys Kx,Vx [R,n_extra,D]; fp32 output O [R,T,D].  The hook does the splits / transposes the
 * engine does (Q pre-scaled).  Average kernel time over `iters` launches in *ms_out. */
size_t a2p_test_tc_attention_scratch_bytes(int R, int T, int D, int S, int n_extra);
int a2p_test_tc_attention(int terms, int R, int T, int D, int dh, int S, int n_extra, const float* Q, const float* K,
                          const float* V, const float* Kx, const float* Vx, float* O, void* scratch, size_t scratch_bytes,
                          int iters, float* ms_out, void* stream);
/* umma_attn2_kernel scheduling: 1 = persistent CTAs (one per SM walks the work items), 0 = one CTA per work item, -1 = library default */
void a2p_test_attn2_set_persist(int on);
/* the exact-fp32 FFMA attention on the same inputs */
int a2p_test_simt_attention(int R, int T, int D, int dh, int S, int n_extra, const float* Q, const float* K, const float* V,
                            const float* Kx, const float* Vx, float* O, int iters, float* ms_out, void* stream);

/* the fused row-chain kernel (umma_chain.cuh; D = 256, split-bf16 x2): GEMM0 (A0[M,K0] * W0[256,K0]^T) -> FiLM/residual on
 * x[M,256] (film[M/T][512] = scale | shift) -> LayerNorm -> optional RoPE -> GEMM1 (W1[N1,256]) -> bf16 planes Cp_out
 * [2][M][N1]; optional V job (W2[256,256] on the un-rotated LayerNorm output) -> Vt_out [2][256][align8(M)].
 * fp32 inputs are split by the hook.  iters <= 0: one launch (x is updated once); iters > 0: timing. */
size_t a2p_test_chain_scratch_bytes(int M, int K0, int N1, int T);
/* 1: one CTA per 128-row tile; 2: CTA pairs (cta_group::2 MMAs, half of every weight tile per CTA); 0: the library default */
void a2p_test_chain_set_mode(int cl);
void a2p_test_chain_set_nsplit(int n);
int a2p_test_chain_nsplit_policy(int tiles, int n_acc, int concurrent, int K0);   /* host logic only: parts per tile the engine would choose */   /* > 0: every tile is worked on by n CTAs (N split of GEMM1 / the V job); 0 = off */
int a2p_test_chain(int M, int T, int K0, int N1, int film_mode, int ln_mode, int rope, int gelu, int vjob, float out_scale,
                   int scale_ncols, const float* A0, const float* W0, const float* bias0, const float* film, float* x,
                   const float* ln_w, const float* ln_b, const float* rope_freqs, const float* W1, const float* bias1,
                   const float* W2, const float* bias2, void* Cp_out, void* Vt_out, void* scratch, size_t scratch_bytes,
                   int iters, float* ms_out, void* stream);

/* Philox4x32-10 block function used by the in-kernel noise of K3 (host build of the same inline code; known-answer tests) */
void a2p_test_philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t* out4);

/* micro-benchmark: total SM cycles for n_mma back-to-back tcgen05.mma (M=128, K=16, bf16) with the given N, A operand
 * from shared memory (0) or tensor memory (1); result written to the DEVICE pointer cycles_out_dev[0]. */
int a2p_test_mma_rate(int N, int a_from_tmem, int n_mma, long long* cycles_out_dev, void* stream);

/* micro-benchmark: cycles for n_warps warps (1..8; warps 0-3 cover the four TMEM quarters, 4-7 share them) to each issue n_ops
 * tcgen05.ld (store = 0) or tcgen05.st (store = 1) of 32 lanes x 32 columns (4 KB per instruction); cycles_out_dev[0]. */
int a2p_test_tmem_ldst_rate(int store, int n_ops, int n_warps, long long* cycles_out_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
