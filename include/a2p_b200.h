/*
 * a2p_b200.h -- C-ABI of the B200-native diffusion-sampling hot path (liba2p_b200.so).
 *
 * The reference (facebookresearch/audio2photoreal) has no FFI layer: its seam for this path is the
 * Python callable protocol between the sampler and the denoiser.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference tree).  INTEGRATION.md shows the
 * ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch types.  All pointers are DEVICE pointers unless
 *     stated otherwise.  The library never allocates tensors: the caller (PyTorch) owns all device
 *     memory and passes workspaces whose sizes come from the *_bytes() queries.
 *   - every function returns 0 on success, non-zero on error; a2p_last_error() gives the message.
 *     Nothing throws across the boundary.
 *   - all work is enqueued on the caller's cudaStream_t (passed as void*); no implicit sync.
 *   - a handle is not thread-safe: one per (device, stream).
 *   - fp32 everywhere at the boundary; timesteps are int64 like the reference's `times`.
 */
#ifndef A2P_B200_H
#define A2P_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define A2P_ABI_VERSION 1

#define A2P_FMT_POSE 0
#define A2P_FMT_FACE 1

#define A2P_BRANCH_COND 0   /* cond_drop_prob = 0.0 */
#define A2P_BRANCH_UNCOND 1 /* cond_drop_prob = 1.0 */

#define A2P_MASK_COND 1
#define A2P_MASK_UNCOND 2
#define A2P_MASK_BOTH 3

#define A2P_SAMPLER_DDIM 0      /* diffusion/gaussian_diffusion.py:667-718 */
#define A2P_SAMPLER_ANCESTRAL 1 /* diffusion/gaussian_diffusion.py:434-477 (with the noise repair) */

#define A2P_LAYOUT_BC1T 0 /* x as [B, C, 1, T] (what the sampler passes) */
#define A2P_LAYOUT_BTC 1  /* x as [B, T, C] */

/* columns of one row of the [n_steps, 8] fp32 coefficient table (host builds it in fp64 -> fp32,
 * audio2photoreal_b200/schedule.py: step_coefficients) */
#define A2P_COL_A 0      /* float(sqrt_recip_alphas_cumprod[i])            gaussian_diffusion.py:347-351 */
#define A2P_COL_B 1      /* float(sqrt_recipm1_alphas_cumprod[i]) */
#define A2P_COL_CX0 2    /* sqrt(float(alphas_cumprod_prev[i]))           :710-713 */
#define A2P_COL_CEPS 3   /* sqrt(1 - abar_prev - sigma^2) */
#define A2P_COL_SIGMA 4  /* [i != 0] * sigma(eta)                         :703-718 */
#define A2P_COL_COEF1 5  /* float(posterior_mean_coef1[i])                :243-246 */
#define A2P_COL_COEF2 6  /* float(posterior_mean_coef2[i]) */
#define A2P_COL_STD 7    /* [i != 0] * exp(0.5 * float(log_variance[i]))  :471-476 */

typedef struct a2p_denoiser a2p_denoiser_t;

/* Geometry of one FiLMTransformer (model/diffusion.py:83-199, utils/model_util.py:49-76). */
typedef struct {
  int32_t fmt;         /* A2P_FMT_POSE | A2P_FMT_FACE */
  int32_t C;           /* nfeats: 104 pose / 256 face */
  int32_t D;           /* latent_dim: 256 / 512 */
  int32_t L;           /* decoder layers */
  int32_t H;           /* heads; D/H must be 32 or 64 */
  int32_t FF;          /* 1024 */
  int32_t S2;          /* keyframe tokens (pose: ceil(max_seq_length/30) = 20); 0 for face */
  int32_t max_pos;     /* rows of the RoPE cos/sin table, >= max(T, S_audio + 2) */
  int32_t split_terms; /* 0: exact-fp32 FFMA kernels; 2 or 3: split-bf16 tcgen05 tensor-core kernels */
  int32_t reserved;
} a2p_model_cfg;

/* One entry of the weight table handed to bind_weights: reference state_dict name -> fp32 device ptr. */
typedef struct {
  const char* name;
  const float* ptr;
  int64_t numel;
} a2p_weight_t;

/* ---- library ------------------------------------------------------------------------------------ */
int a2p_abi_version(void);
const char* a2p_last_error(void);
/* 1 if the build contains the sm_100a tcgen05 kernels (always 1 for a normal build). */
int a2p_has_tcgen05(void);

/* ---- denoiser: replaces FiLMTransformer.forward (model/diffusion.py:338-403) and, with
 *      A2P_MASK_BOTH, both calls of ClassifierFreeSampleModel.forward (model/cfg_sampler.py:30-33) */
int a2p_denoiser_create(a2p_denoiser_t** out, const a2p_model_cfg* cfg);
void a2p_denoiser_destroy(a2p_denoiser_t* h);

/* bytes of the derived-weight arena (stacked FiLM / time-token projections, permuted conv taps,
 * RoPE table, split-bf16 planes) the caller must provide to bind_weights. */
size_t a2p_packed_weight_bytes(const a2p_model_cfg* cfg);

/* Borrow the fp32 parameters (reference state_dict layout, utils/model_util.py:30-38) for the life
 * of the handle and build the derived arena.  Unknown names are ignored; a missing required name is
 * an error naming it. */
int a2p_denoiser_bind_weights(a2p_denoiser_t* h, const a2p_weight_t* table, int n_entries, void* packed,
                              size_t packed_bytes, void* stream);

/* Per-branch cache of the step-invariant attention memories (rotated-K / V projections of the audio
 * tokens and of the keyframe tokens for every layer). */
size_t a2p_kv_cache_bytes(const a2p_model_cfg* cfg, int Bc, int S);

/* Step-invariant conditioning of one CFG branch -- everything FiLMTransformer.forward recomputes per
 * call before the decoder stack (model/diffusion.py:372-393) minus the frozen audio encoders, which
 * stay in PyTorch:
 *   cond_tokens [Bc, S, D]  tokens AFTER cond_projection / cond_encoder / null-embedding select and
 *                           BEFORE norm_cond (norm_cond is applied here; per-row LayerNorm)
 *   cond_hidden [Bc, D]     non_attn_cond_projection(mean-pooled tokens) or null_cond_hidden
 *   pose_tokens [Bc, S2, D] encode_keyframes output or null_pose_embed[:, :S2] (pose only, 0 < S2 <=
 *                           cfg.S2; NULL and S2 = 0 for face)
 * Bc is the batch size B, or 1 when the conditioning is shared by every row (the uncond branch).
 * kv_cache must stay alive and unmodified until the next set_conditioning of the same branch. */
int a2p_denoiser_set_conditioning(a2p_denoiser_t* h, int branch, int Bc, int S, int S2, const float* cond_tokens,
                                  const float* cond_hidden, const float* pose_tokens, void* kv_cache,
                                  size_t kv_bytes, void* ws, size_t ws_bytes, void* stream);

/* The step-invariant conditioning ENCODERS of one batch, native (exact fp32 FFMA GEMMs / fp32 attention): what
 * FiLMTransformer.forward computes from the frozen encoders' features before the null-embedding select
 * (model/diffusion.py:355-381, 316-336; transformer_modules.py:69-102):
 *   feats     [Bc, S, feat_dim]  encode_audio output (+ lip features for face): 1024 (pose) / 2038 (face) columns
 *   keyframes [Bc, S2, C]        pose only: y["keyframes"] with the unknown keyframes already zeroed; NULL / S2 = 0 for face
 *   cond_tokens [Bc, S, D]   <-  cond_projection (-> face: the two rotary pre-LN layers of cond_encoder)
 *   cond_hidden [Bc, D]      <-  non_attn_cond_projection(mean over S of cond_tokens)
 *   pose_tokens [Bc, S2, D]  <-  frame_norm_cond(frame_cond_projection(keyframes))      (pose only)
 * The outputs are exactly the inputs of a2p_denoiser_set_conditioning(branch 0).  Needs cond_projection.*,
 * non_attn_cond_projection.{0,1,3}.*, and frame_cond_projection.* / frame_norm_cond.* (pose) or cond_encoder.{0,1}.*
 * (face) in the bound weight table; fails if they are missing.  A row's result does not depend on Bc. */
size_t a2p_encode_workspace_bytes(const a2p_model_cfg* cfg, int Bc, int S, int feat_dim);
int a2p_denoiser_encode_conditioning(a2p_denoiser_t* h, int Bc, int S, int S2, int feat_dim, const float* feats,
                                     const float* keyframes, float* cond_tokens, float* cond_hidden,
                                     float* pose_tokens, void* ws, size_t ws_bytes, void* stream);

/* scratch needed by set_conditioning (normalised + rotated copies of the memories). */
size_t a2p_conditioning_workspace_bytes(const a2p_model_cfg* cfg, int Bc, int S);

/* workspace for forward / sample_loop with B samples of T frames (both branches). */
size_t a2p_workspace_bytes(const a2p_model_cfg* cfg, int B, int T);

/* One denoiser evaluation.  x: [B,C,1,T] or [B,T,C] per x_layout; timesteps: [B] int64 original-scale
 * (already mapped through timestep_map, diffusion/respace.py:140-145); outputs [B,T,C] (NULL allowed
 * for a branch not in branch_mask). */
int a2p_denoiser_forward(a2p_denoiser_t* h, int B, int T, const float* x, int x_layout, const int64_t* timesteps,
                         int branch_mask, float* out_cond, float* out_uncond, void* ws, size_t ws_bytes,
                         void* stream);

/* ---- sampler epilogue (K3): CFG mix + layout permute + x0 -> x_{t-1}.  Replaces
 *      cfg_sampler.py:33 + gaussian_diffusion.py:302-316 + :699-718 (DDIM) / :243-246,:471-476.
 *   x_t, x_prev, pred_xstart: [B,C,1,T];  x0_cond/x0_uncond: [B,T,C] denoiser outputs
 *   x0_uncond == NULL -> no guidance (x0 = x0_cond);  scale: [B] (y["scale"])
 *   coeffs: DEVICE pointer to one 8-float row (A2P_COL_*);  noise: [B,C,1,T] or NULL (treated as 0)
 *   clip_denoised: clamp x0 to [-1,1] (gaussian_diffusion.py:305-310) */
int a2p_sampler_step(int kind, int B, int C, int T, const float* x_t, const float* x0_cond, const float* x0_uncond,
                     const float* scale, const float* coeffs, const float* noise, int clip_denoised,
                     float* x_prev, float* pred_xstart, void* stream);

/* ---- whole reverse loop: replaces ddim_sample_loop / p_sample_loop bodies
 *      (gaussian_diffusion.py:815-936 / :525-665) for the CFG-wrapped denoiser.  One diffusion step
 *      is captured into a CUDA graph and replayed n_steps times; per-step scalars are read on the
 *      device from `coeffs`/`timesteps` through a device-side step counter.
 *   coeffs    [n_steps, 8] fp32, row i = respaced step index i
 *   timesteps [n_steps] int64, timestep_map[i]
 *   x         [B,C,1,T] in: x_T, out: final `sample`;  pred_xstart [B,C,1,T] out: last x0 prediction
 *   noise_tape [n_steps, B,C,1,T] consumed in loop order (first row = step n_steps-1) or NULL
 *   branch_mask A2P_MASK_BOTH (CFG) or A2P_MASK_COND (bare denoiser, scale ignored) */
int a2p_sample_loop(a2p_denoiser_t* h, int kind, int B, int T, int n_steps, const float* coeffs,
                    const int64_t* timesteps, const float* scale, float* x, float* pred_xstart,
                    const float* noise_tape, int clip_denoised, int branch_mask, int use_graph, void* ws,
                    size_t ws_bytes, void* stream);

/* Same loop / step with IN-KERNEL noise instead of a tape: every (batch row, channel, frame, loop
 * iteration) draws N(0,1) from Philox4x32-10 keyed by `seed` + Box-Muller inside K3.  This replaces the per-step
 * th.randn_like(x) of p_sample (gaussian_diffusion.py:471-476, upstream form) and of ddim_sample at eta > 0 (:708-713)
 * STATISTICALLY, not bitwise (parity runs use the tape).  row0 = global index of this call's batch row 0, so a batch
 * sharded over ranks draws exactly the numbers the unsharded batch would.  No [n_steps,B,C,1,T] tape (2 GB at B = 8). */
int a2p_sample_loop_rng(a2p_denoiser_t* h, int kind, int B, int T, int n_steps, const float* coeffs,
                        const int64_t* timesteps, const float* scale, float* x, float* pred_xstart, uint64_t seed,
                        int64_t row0, int clip_denoised, int branch_mask, int use_graph, void* ws, size_t ws_bytes,
                        void* stream);
int a2p_sampler_step_rng(int kind, int B, int C, int T, const float* x_t, const float* x0_cond, const float* x0_uncond,
                         const float* scale, const float* coeffs, uint64_t seed, int64_t iteration, int64_t row0,
                         int clip_denoised, float* x_prev, float* pred_xstart, void* stream);

/* Measurement aid (bench.py roofline): one un-captured denoiser evaluation with a CUDA-event pair around
 * every kernel; ms_by_cat / launches_by_cat are HOST arrays of ncat >= 9 entries, categories:
 * 0 time-conditioning GEMMs, 1 LayerNorm+RoPE, 2 attention projections, 3 self-attention core,
 * 4 audio cross-attention core, 5 keyframe cross-attention core, 6 feed-forward, 7 in/out projection + TCN,
 * 8 misc.  x_btc: [B,T,C].  Synchronises the stream. */
int a2p_profile_forward(a2p_denoiser_t* h, int B, int T, const float* x_btc, const int64_t* timesteps, int branch_mask,
                        void* ws, size_t ws_bytes, void* stream, float* ms_by_cat, int64_t* launches_by_cat, int ncat);

/* Same for batch rows [b0, b0 + Bs) of a batch of B_total rows (x_btc / timesteps point at row b0): the launch shapes of the
 * sampling loop of the fused arm, which cuts a CFG step into concurrent forwards over groups of rows. */
int a2p_profile_forward_rows(a2p_denoiser_t* h, int B_total, int b0, int Bs, int T, const float* x_btc, const int64_t* timesteps,
                             int branch_mask, void* ws, size_t ws_bytes, void* stream, float* ms_by_cat,
                             int64_t* launches_by_cat, int ncat);
/* Number of row groups per CFG branch the sampling loop uses for a batch of B rows (1 = whole branch per forward). */
int a2p_loop_row_groups(const a2p_denoiser_t* h, int B, int T);

/* kernels launched by this handle since creation (for bench.py's gpu_launches). */
int64_t a2p_launch_count(const a2p_denoiser_t* h);

#ifdef __cplusplus
}
#endif
#endif /* A2P_B200_H */
