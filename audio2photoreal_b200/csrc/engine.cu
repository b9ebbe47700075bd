// a2p_b200 engine: the denoiser handle, conditioning caches, one denoiser evaluation, the fused sampler
// epilogue and the graph-captured reverse loop, exported through the C-ABI of include/a2p_b200.h.
//
// Reference behaviour implemented (paths relative to the reference tree):
//   model/diffusion.py:338-403                       FiLMTransformer.forward (step-dependent part)
//   model/modules/transformer_modules.py:190-267     FiLMTransformerDecoderLayer (pre-LN branch)
//   model/cfg_sampler.py:30-33                       CFG: both branches batched as 2B rows
//   diffusion/gaussian_diffusion.py:667-718,434-477  DDIM / ancestral update (K3)
//   diffusion/respace.py:140-145                     timestep_map lookup (host builds the table)
#include <cuda_runtime.h>
#include <stdlib.h>
#include <map>
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

using namespace a2p;

namespace {

constexpr int MAXL = 16;
constexpr int TCN_PAD = 24;  // receptive_field - 1 (model/diffusion.py:153,215)
const int TCN_DIL[6] = {1, 2, 3, 1, 2, 3};

struct AttnW { const float *in_w, *in_b, *out_w, *out_b; };
struct LayerW {
  AttnW sa, ca, c2;
  const float *l1w, *l1b, *l2w, *l2b;
  const float *n1w, *n1b, *n2w, *n2b, *n2aw, *n2ab, *n3w, *n3b;
};

struct CondSet {
  bool set = false;
  int Bc = 0, S = 0, S2 = 0;
  float* base = nullptr;      // kv cache arena
  const float* hidden = nullptr;  // [Bc, D] copy inside the arena
};

struct GraphKey {
  int B, T, kind, n_steps, clip, mask;
  const void *coeffs, *ts, *scale, *x, *pred, *noise, *ws;
  const void *kv0, *kv1;
  unsigned long long seed; long long row0; int rng;
  int units;   // how the step was cut into concurrent forwards when the graph was captured
  int pipes;   // independent per-group pipelines (1: all groups joined every step)
  bool operator==(const GraphKey& o) const {
    return B == o.B && T == o.T && kind == o.kind && n_steps == o.n_steps && clip == o.clip && mask == o.mask &&
           coeffs == o.coeffs && ts == o.ts && scale == o.scale && x == o.x && pred == o.pred && noise == o.noise &&
           ws == o.ws && kv0 == o.kv0 && kv1 == o.kv1 && seed == o.seed && row0 == o.row0 && rng == o.rng && units == o.units && pipes == o.pipes;
  }
};

}  // namespace

struct a2p_denoiser {
  a2p_model_cfg cfg{};
  int dh = 0, nf = 0;
  bool bound = false;
  LayerW lw[MAXL]{};
  const float *time_w1, *time_b1, *time_w2, *time_b2, *time_w3, *time_b3;
  const float *normc_w, *normc_b, *inp_w, *inp_b, *fin_w, *fin_b, *fconv_w, *fconv_b;
  // step-invariant conditioning encoders (a2p_denoiser_encode_conditioning; optional in the weight table)
  struct EncLayerW { const float *n1w, *n1b, *n2w, *n2b, *in_w, *in_b, *out_w, *out_b, *l1w, *l1b, *l2w, *l2b; };
  struct EncW {
    const float *cp_w = nullptr, *cp_b = nullptr;                                   // cond_projection [D, feat_dim]
    const float *p0w = nullptr, *p0b = nullptr, *p1w = nullptr, *p1b = nullptr, *p3w = nullptr, *p3b = nullptr;   // non_attn_cond_projection
    const float *fp_w = nullptr, *fp_b = nullptr, *fn_w = nullptr, *fn_b = nullptr; // frame_cond_projection, frame_norm_cond (pose)
    EncLayerW enc[2]{};                                                             // cond_encoder (face)
    int feat_dim = 0;
    bool ok = false;
  } encw;
  const float* conv_b[6];
  const float* time_freqs;
  // derived arena
  float *film_w = nullptr, *film_b = nullptr, *ttk_w = nullptr, *ttk_b = nullptr, *ttv_w = nullptr, *ttv_b = nullptr;
  float* conv_w[6]{};
  float* conv_t[6]{};   // tap-major fp32 copies, keys of the split planes
  float2* rope_tab = nullptr;
  std::map<const float*, __nv_bfloat16*> wplanes;  // fp32 weight -> split-bf16 planes [P][rows][cols] (plane stride = numel)
  std::map<const float*, long long> wnumel;
  int num_sms = 148;
  int attn_skew_ns = 0;   // start delay of head 1's softmax warpgroup: no effect (profiles/r01f, r02_attn2_decoupled_mma_ab.txt)
  CondSet cond[2];
  int64_t launches = 0;
  int64_t graph_nodes = 0;
  static constexpr int MAXG = 4;      // row groups of a CFG step
  cudaGraphExec_t gexec[MAXG] = {};   // one graph per independent pipeline (one in total when the groups are joined every step)
  int n_gexec = 0;
  cudaStream_t group_stream[MAXG] = {};   // pipelines 1.. run on their own stream (pipeline 0: the caller's)
  cudaEvent_t ev_gstart = nullptr, ev_gdone[MAXG] = {}, ev_gfork[MAXG] = {};
  cudaStream_t cap_stream = nullptr;  // private stream used only to CAPTURE a step (the legacy default stream cannot capture)
  // A sampling step of the fused arm runs as up to MAXU concurrent forwards ("units" = CFG branch x group of batch rows,
  // see sample_loop_impl).  Per unit: a side stream on which the per-step conditioning chain runs beside the first chain /
  // self-attention launches, and (units > 0) the stream the unit's own launches go to.
  static constexpr int MAXU = 8;
  cudaStream_t cond_stream[MAXU] = {};
  cudaEvent_t ev_fork[MAXU] = {}, ev_join[MAXU] = {};
  cudaStream_t unit_stream[MAXU] = {};
  cudaEvent_t ev_ujoin[MAXU] = {};
  cudaEvent_t ev_bfork = nullptr;
  GraphKey gkey{};
  bool gvalid = false;
};

namespace {

// ---------------------------------------------------------------- arena layouts
struct PackedLayout {
  size_t film_w, film_b, ttk_w, ttk_b, ttv_w, ttv_b, conv_w[6], conv_t[6], rope, planes, planes_bytes, total;
};
// elements of every weight that gets split-bf16 planes
size_t split_weight_elems(const a2p_model_cfg& c) {
  const size_t D = c.D, FF = c.FF, C = c.C;
  size_t per_layer = 2 * (3 * D * D + D * D) + 2 * FF * D + (c.fmt == A2P_FMT_POSE ? 4 * D * D : 0);
  size_t conv = 0;
  if (c.fmt == A2P_FMT_POSE) {
    const size_t cm = C > 256 ? C : 256;
    conv = 3 * (cm * C * 2 + C * C * 4) + C * C;   // [3][Cout][Cin] per TCN layer + the 1x1 final conv
  }
  return per_layer * c.L + 2 * D * C + conv;
}
PackedLayout packed_layout(const a2p_model_cfg& c) {
  PackedLayout p{};
  size_t off = 0;
  auto take = [&](size_t nfloats) { size_t o = off; off = align_up(off + nfloats * 4, 256); return o; };
  const int nf = c.fmt == A2P_FMT_POSE ? 4 : 3;
  p.film_w = take((size_t)c.L * nf * 2 * c.D * c.D);
  p.film_b = take((size_t)c.L * nf * 2 * c.D);
  p.ttk_w = take((size_t)c.L * c.D * c.D);
  p.ttk_b = take((size_t)c.L * c.D);
  p.ttv_w = take((size_t)c.L * c.D * c.D);
  p.ttv_b = take((size_t)c.L * c.D);
  if (c.fmt == A2P_FMT_POSE) {
    const int cm = c.C > 256 ? c.C : 256;
    const int chans[6][2] = {{cm, c.C}, {c.C, cm}, {c.C, c.C}, {c.C, c.C}, {c.C, c.C}, {c.C, c.C}};
    for (int i = 0; i < 6; ++i) p.conv_w[i] = take((size_t)chans[i][0] * chans[i][1] * 3);
    for (int i = 0; i < 6; ++i) p.conv_t[i] = take((size_t)chans[i][0] * chans[i][1] * 3);   // tap-major [3][Cout][Cin] copy (TC arm)
  }
  p.rope = take((size_t)c.max_pos * (c.D / 2) * 2);
  p.planes = off;
  p.planes_bytes = c.split_terms > 0 ? align_up(split_weight_elems(c) * 2 * c.split_terms + 64 * 256, 256) : 0;
  off += p.planes_bytes;
  p.total = off;
  return p;
}

struct KvLayout {
  size_t per_layer, ka, va, k2, v2, hidden, total;  // float offsets
  size_t kaP, vtaP, k2P, vt2P;                      // tensor-core arm: split-bf16 K planes / V^T planes (float offsets)
  size_t Sp, S2p;                                   // per-sample column pitch of the V^T planes (multiple of 8)
};
KvLayout kv_layout(const a2p_model_cfg& c, int Bc, int S, int S2) {
  KvLayout k{};
  size_t a = (size_t)Bc * S * c.D, b = (size_t)Bc * S2 * c.D;
  k.ka = 0; k.va = a; k.k2 = 2 * a; k.v2 = 2 * a + b;
  k.per_layer = 2 * a + 2 * b;
  k.Sp = align_up((size_t)S, 8); k.S2p = align_up((size_t)(S2 > 0 ? S2 : 1), 8);
  if (c.split_terms > 0) {
    const size_t P = c.split_terms;
    auto take = [&](size_t bf16s) { size_t o = k.per_layer; k.per_layer += align_up(bf16s / 2 + 1, 64); return o; };
    k.per_layer = align_up(k.per_layer, 64);
    k.kaP = take(P * Bc * S * c.D);
    k.vtaP = take(P * c.D * Bc * k.Sp);
    k.k2P = take(P * (b ? b : 1));
    k.vt2P = take(P * c.D * Bc * k.S2p);
  }
  k.hidden = k.per_layer * c.L;
  k.total = k.hidden + (size_t)Bc * c.D;
  return k;
}

struct WsLayout {
  size_t counter, e, th, mt, ttok, tt, ttr, ktt, vtt, film, xin, x, h, hr, qkv, att, u, out, tcnA, tcnB, tcnC, total;
  size_t hP, hrP, attP, uP, xinP;   // split-bf16 activation planes (tensor-core arm)
  size_t tcnUP, tcnVP;              // TCN activation planes in the left-padded layout
  size_t qkP, vtS, kttP, vttT;      // tensor-core attention operands: Q|K planes, self V^T planes, time-token K / V^T planes
  size_t ropeX;                     // chain arm: RoPE table extended to T + 128 rows (one TMA box per 128-row tile)
  size_t splitS, splitC;            // attention split-KV tail: partial results and per-tile arrival counters
};
WsLayout ws_layout(const a2p_model_cfg& c, int B, int T) {
  WsLayout w{};
  size_t off = 0;
  auto take = [&](size_t nfloats) { size_t o = off; off = align_up(off + nfloats * 4, 256); return o; };
  const size_t R = 2 * (size_t)B, D = c.D;
  const int nf = c.fmt == A2P_FMT_POSE ? 4 : 3;
  w.counter = take(64);
  w.e = take(R * D); w.th = take(R * 4 * D); w.mt = take(R * D); w.ttok = take(R * 2 * D);
  w.tt = take(2 * R * D); w.ttr = take(2 * R * D);
  w.ktt = take(2 * R * c.L * D); w.vtt = take(2 * R * c.L * D);
  w.film = take(R * c.L * nf * 2 * D);
  w.xin = take((size_t)B * T * c.C);
  w.x = take(R * T * D); w.h = take(R * T * D); w.hr = take(R * T * D);
  w.qkv = take(R * T * 3 * D); w.att = take(R * T * D); w.u = take(R * T * c.FF);
  w.out = take(R * T * c.C);
  if (c.fmt == A2P_FMT_POSE) {
    const size_t cm = c.C > 256 ? c.C : 256;
    w.tcnA = take(R * (T + TCN_PAD) * c.C); w.tcnB = take(R * (T + TCN_PAD) * cm); w.tcnC = take(R * (T + TCN_PAD) * c.C);
  }
  if (c.split_terms > 0) {
    const size_t P = c.split_terms;   // bf16 = half a float
    w.hP = take(P * R * T * D / 2 + 64); w.hrP = take(P * R * T * D / 2 + 64); w.attP = take(P * R * T * D / 2 + 64);
    w.uP = take(P * R * T * c.FF / 2 + 64); w.xinP = take(P * R * T * (D > c.C ? D : c.C) / 2 + 64);
    w.qkP = take(P * R * T * 2 * D / 2 + 64); w.vtS = take(P * D * align_up(R * T, 8) / 2 + 64);
    w.kttP = take(P * (2 * R + 64) * c.L * D / 2 + 64); w.vttT = take(P * c.L * D * (8 * R) / 2 + 64);
    w.ropeX = take((size_t)(T + 128) * D);
    w.splitS = take(attn2_split_scratch_floats()); w.splitC = take(attn2_split_counter_ints());
    if (c.fmt == A2P_FMT_POSE) {
      const size_t cm = c.C > 256 ? c.C : 256;
      w.tcnUP = take(P * R * (T + TCN_PAD) * cm / 2 + 64); w.tcnVP = take(P * R * (T + TCN_PAD) * cm / 2 + 64);
    }
  }
  w.total = off;
  return w;
}

// ---------------------------------------------------------------- small kernels local to the engine
__global__ void rope_only_kernel(const float* __restrict__ x, float* __restrict__ out, const float2* __restrict__ tab,
                                 int D, int pos_mod, long long rows) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  int half = D / 2;
  if (idx >= rows * half) return;
  long long r = idx / half;
  int i = idx - r * half;
  int pos = r % pos_mod;
  float2 cs = tab[(long long)pos * half + i];
  float a = x[r * D + 2 * i], b = x[r * D + 2 * i + 1];
  out[r * D + 2 * i] = a * cs.x - b * cs.y;
  out[r * D + 2 * i + 1] = b * cs.x + a * cs.y;
}

// conv weight [Cout, Cin, 3] -> [Cout, 3*Cin] with k = tap*Cin + ci
__global__ void permute_conv_w_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * Cin * 3) return;
  int co = idx / (Cin * 3), rem = idx - co * Cin * 3, tap = rem / Cin, ci = rem - tap * Cin;
  out[idx] = w[((long long)co * Cin + ci) * 3 + tap];
}

// conv weight [Cout, Cin, 3] -> tap-major [3][Cout][Cin]
__global__ void permute_conv_w_tapmajor_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Cout * Cin * 3) return;
  int tap = idx / (Cout * Cin), rem = idx - tap * Cout * Cin, co = rem / Cin, ci = rem - co * Cin;
  out[idx] = w[((long long)co * Cin + ci) * 3 + tap];
}

// dst[r][p][c] = p < pad ? 0 : src[r][p-pad][c]
__global__ void pad_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int T, int pad, int C4, long long total) {
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int c = idx % C4;
  long long rp = idx / C4;
  int pp = rp % (T + pad);
  long long r = rp / (T + pad);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (pp >= pad) v = reinterpret_cast<const float4*>(src)[(r * T + (pp - pad)) * C4 + c];
  reinterpret_cast<float4*>(dst)[idx] = v;
}

// optional per-category CUDA-event profiler (bench.py's live roofline measurement; never used inside a capture)
enum Cat : int { CAT_COND = 0, CAT_LN, CAT_PROJ, CAT_ATT_SELF, CAT_ATT_CROSS, CAT_ATT_CROSS2, CAT_FFN, CAT_IO_TCN, CAT_MISC, CAT_N };
struct Prof {
  std::vector<cudaEvent_t> ev;
  std::vector<int> cat;
  std::vector<std::string> info;
};
struct Ctx {
  a2p_denoiser* h;
  cudaStream_t st;
  Prof* prof = nullptr;
  int cat = CAT_MISC;
  bool skinny = false;   // route the next FFMA GEMMs to the warp-per-column kernel (per-step conditioning linears)
  int slot = 0;          // which set of side streams / events of the handle this forward uses
  int concurrent = 1;    // how many forwards of this size share the GPU (the units of a sampling step): sizes the chain N split
  cudaEvent_t stagger_ev = nullptr;   // recorded after the stagger_after-th chain / attention launch of the fused arm
  int stagger_after = 0, n_main = 0;
  std::string tag;
  int mark() {           // one more chain / attention launch is in the stream
    if (stagger_ev && ++n_main == stagger_after) { A2P_CUDA(cudaEventRecord(stagger_ev, st)); stagger_ev = nullptr; }
    return 0;
  }
  void begin() {
    if (!prof) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); prof->ev.push_back(e); prof->cat.push_back(cat);
    prof->info.push_back(tag); tag.clear();
  }
  void end() {
    if (!prof) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); prof->ev.push_back(e);
  }
};

int gemm(Ctx& c, const float* A, long long lda, int M, const float* W, long long ldw, const float* bias, int N, int K,
         float* C, long long ldc, int epi = EPI_BIAS, GemmParams* extra = nullptr) {
  GemmParams p{};
  if (extra) p = *extra;
  p.A = A; p.lda = lda; p.W = W; p.ldw = ldw; p.bias = bias; p.C = C; p.ldc = ldc; p.M = M; p.N = N; p.K = K;
  if (p.taps == 0) { p.taps = 1; p.dil = 0; p.Kc = K; }
  p.epi = epi;
  c.h->launches++;
  if (c.prof) { char b_[96]; snprintf(b_, sizeof(b_), "ffma_gemm M=%d N=%d K=%d epi=%d", M, N, K, epi); c.tag = b_; }
  c.begin();
  int rc = (c.skinny && skinny_ok(p)) ? launch_skinny_gemm(p, c.st) : launch_sgemm(p, c.st);
  c.end();
  return rc;
}

int film_gemm(Ctx& c, const float* A, long long lda, int M, const float* W, const float* bias, int N, int K, float* x,
              long long ldx, const float* film, long long film_ld, int film_off, int D, int rows_per_sample) {
  GemmParams e{};
  e.film = film; e.film_ld = film_ld; e.film_scale_off = film_off; e.film_shift_off = film_off + D;
  e.rows_per_sample = rows_per_sample;
  return gemm(c, A, lda, M, W, K, bias, N, K, x, ldx, EPI_FILM_RESID, &e);
}

// tensor-core GEMM on pre-split operands: A planes [P][M][K]; W = rows [w_row0, w_row0+N) of the fp32 weight `wkey` [*,K]
int tc_gemm(Ctx& c, const __nv_bfloat16* Ap, int M, int K, const float* wkey, long long w_row0, int N, const float* bias, int epi,
            TcGemmParams p) {
  a2p_denoiser* h = c.h;
  auto it = h->wplanes.find(wkey);
  if (it == h->wplanes.end()) A2P_FAIL("tc_gemm: weight has no split planes");
  if (p.taps <= 0) { p.taps = 1; p.dil = 0; }
  TcOperands o{Ap, K, (long long)M * K, it->second + w_row0 * K, K, h->wnumel[wkey] / p.taps};
  p.M = M; p.N = N; p.K = K; p.bias = bias;
  if (p.out_scale == 0.f) p.out_scale = 1.f;
  h->launches++;
  if (c.prof) { char b_[96]; snprintf(b_, sizeof(b_), "tc_gemm M=%d N=%d K=%d taps=%d epi=%d", M, N, K, p.taps, epi); c.tag = b_; }
  c.begin();
  int rc = launch_umma_gemm(h->cfg.split_terms, o, p, epi, h->num_sms, c.st);
  c.end();
  return rc;
}

int tc_film_gemm(Ctx& c, const __nv_bfloat16* Ap, int M, int K, const float* wkey, const float* bias, int N, float* x, long long ldx,
                 const float* film, long long film_ld, int film_off, int D, int rows_per_sample) {
  TcGemmParams p{};
  p.C = x; p.ldc = ldx; p.film = film; p.film_ld = film_ld; p.film_scale_off = film_off; p.film_shift_off = film_off + D;
  p.rows_per_sample = rows_per_sample;
  return tc_gemm(c, Ap, M, K, wkey, 0, N, bias, TC_FILM, p);
}

const float* find(const std::map<std::string, std::pair<const float*, int64_t>>& m, const std::string& k, int64_t numel,
                  std::string& err) {
  auto it = m.find(k);
  if (it == m.end()) { if (err.empty()) err = "missing weight '" + k + "'"; return nullptr; }
  if (numel >= 0 && it->second.second != numel) {
    if (err.empty()) err = "weight '" + k + "' has " + std::to_string(it->second.second) + " elements, expected " + std::to_string(numel);
    return nullptr;
  }
  return it->second.first;
}

// A2P_NO_CHAIN=1 keeps the unfused GEMM / LayerNorm kernels (A/B measurements and the P = 3 / face arms use them anyway)
// A2P_ATTN2: 0 = first-generation attention kernel, 1 = head-parallel kernel with P planes in shared memory,
// 2 (default) = head-parallel kernel with P and Q planes in tensor memory (umma_attention2.cuh; head dim 32, two planes),
// 8 = as 2 with the Q planes in shared memory, 3 / 4 = as 8 with 1 / 2 of every 4 exponentials on the FMA pipe
int attn2_variant() {
  static int v = -1;
  if (v < 0) v = getenv("A2P_ATTN2") ? atoi(getenv("A2P_ATTN2")) : 2;
  return v;
}

// CFG sampling: run the conditional and the unconditional forward as two concurrent streams of launches (env
// A2P_NO_BRANCH_STREAMS=1 keeps the single stacked forward).  A2P_BRANCH_STAGGER=n starts the unconditional forward after
// the n-th chain / attention launch of the conditional one, so that one branch's 1-CTA-per-tile chain kernels (which
// leave half of the SMs idle at small batches) overlap the other branch's attention kernels.
// A2P_BRANCH_GROUPS=g (1, 2 or 4): additionally cut the batch rows of every branch into g groups (units = 2 g forwards).
// Default: 2 groups for 4..16 rows per branch, else 1 (measured, B = 8: 192.7 / 182.6 / 222.6 ms per 100 steps for g = 1 / 2 / 4;
// B = 32: g = 2 is 2 % slower than g = 1 -- profiles/r01v_unit_groups_sweep.txt)
int branch_groups(int B) {
  const char* e = getenv("A2P_BRANCH_GROUPS");
  const int v = e ? atoi(e) : ((B >= 4 && B <= 16) ? 2 : 1);
  return v >= 4 ? 4 : (v >= 2 ? 2 : 1);
}
bool branch_streams_disabled() { return getenv("A2P_NO_BRANCH_STREAMS") != nullptr; }
int branch_stagger() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("A2P_BRANCH_STAGGER"); v = e ? atoi(e) : 0; if (v < 0) v = 0; }
  return v;
}

int chain_priority() {   // A2P_CHAIN_PRIO=p: launch priority of the chain kernels (negative = higher than the attention kernels)
  static int v = 1 << 30;
  if (v == (1 << 30)) { const char* e = getenv("A2P_CHAIN_PRIO"); v = e ? atoi(e) : 0; }
  return v;
}

bool chain_disabled() {
  static int v = -1;
  if (v < 0) v = getenv("A2P_NO_CHAIN") ? 1 : 0;
  return v == 1;
}

int check_cfg(const a2p_model_cfg* c) {
  if (!c) A2P_FAIL("null cfg");
  if (c->fmt != A2P_FMT_POSE && c->fmt != A2P_FMT_FACE) A2P_FAIL("cfg.fmt must be 0 (pose) or 1 (face)");
  if (c->D != 256 && c->D != 512) A2P_FAIL("cfg.D=%d unsupported (256 or 512)", c->D);
  if (c->H <= 0 || c->D % c->H) A2P_FAIL("cfg.H=%d does not divide D=%d", c->H, c->D);
  int dh = c->D / c->H;
  if (dh != 32 && dh != 64) A2P_FAIL("head dim %d unsupported (32 or 64)", dh);
  if (c->L <= 0 || c->L > MAXL) A2P_FAIL("cfg.L=%d out of range", c->L);
  if (c->C % 8 || c->FF % 8) A2P_FAIL("cfg.C/FF must be multiples of 8");
  if (c->fmt == A2P_FMT_POSE && c->S2 <= 0) A2P_FAIL("pose needs S2 > 0");
  if (c->fmt == A2P_FMT_FACE && c->S2 != 0) A2P_FAIL("face must have S2 == 0");
  if (c->split_terms < 0 || c->split_terms > 3) A2P_FAIL("split_terms=%d must be 0 (exact fp32) or 1..3 bf16 planes", c->split_terms);
  if (c->max_pos < 2) A2P_FAIL("cfg.max_pos too small");
  return 0;
}

// ---------------------------------------------------------------- one denoiser evaluation
// xin: [B,T,C] (already transposed).  ts: per-row [B] int64 (counter == nullptr) or the step table.
// b0 / B_total: this call evaluates batch rows [b0, b0 + B) of a batch of B_total rows (xin / ts already point at row b0;
// only the per-sample conditioning caches are addressed with the offset).  b0 > 0 is supported by the fused arm only.
int forward_core(Ctx& c, int B, int T, const float* xin, const long long* ts, const int* counter, int mask, char* wsb,
                 const float** x0_cond, const float** x0_uncond, long long* x0_sample_stride, int b0 = 0, int B_total = 0) {
  if (B_total <= 0) B_total = B;
  a2p_denoiser* h = c.h;
  const a2p_model_cfg& cf = h->cfg;
  const int D = cf.D, L = cf.L, C = cf.C, nf = h->nf, H = cf.H, dh = h->dh;
  const int nb = (mask == A2P_MASK_BOTH) ? 2 : 1;
  const int R = nb * B;
  const int br0 = (mask == A2P_MASK_UNCOND) ? 1 : 0;  // branch of rows [0,B)
  const CondSet& c0 = h->cond[br0];
  const CondSet& c1 = h->cond[1];
  if (!c0.set || (nb == 2 && !c1.set)) A2P_FAIL("conditioning not set for the requested branch(es)");
  if (nb == 2 && c0.S != c1.S) A2P_FAIL("cond/uncond token counts differ (%d vs %d)", c0.S, c1.S);
  if ((c0.Bc != 1 && c0.Bc != B_total) || (nb == 2 && c1.Bc != 1 && c1.Bc != B_total)) A2P_FAIL("conditioning batch does not match B=%d", B_total);
  const int S = c0.S;
  if (T > cf.max_pos || S + 2 > cf.max_pos) A2P_FAIL("T=%d / S+2=%d exceed cfg.max_pos=%d", T, S + 2, cf.max_pos);
  const WsLayout w = ws_layout(cf, B, T);
  auto F = [&](size_t off) { return reinterpret_cast<float*>(wsb + off); };
  float *e = F(w.e), *th = F(w.th), *mt = F(w.mt), *ttok = F(w.ttok), *tt = F(w.tt), *ttr = F(w.ttr);
  float *ktt = F(w.ktt), *vtt = F(w.vtt), *film = F(w.film), *x = F(w.x), *hh = F(w.h), *hr = F(w.hr);
  float *qkv = F(w.qkv), *att = F(w.att), *u = F(w.u), *out = F(w.out);
  const KvLayout k0 = kv_layout(cf, c0.Bc, S, c0.S2), k1 = kv_layout(cf, c1.Bc, S, c1.S2);
  if (nb == 2 && c0.S2 != c1.S2) A2P_FAIL("cond/uncond keyframe token counts differ (%d vs %d)", c0.S2, c1.S2);
  const int S2 = c0.S2;
  const long long film_ld = (long long)L * nf * 2 * D;
  cudaStream_t st = c.st;

  // --- time conditioning (model/diffusion.py:384-389, model/utils.py:67-79)
  // It depends on the timestep only, not on x_t: on the fused arm it runs on a side stream beside the input-projection chain
  // launch and the first self-attention (fork / join by events; inside a graph capture these become parallel branches) and
  // is joined before the first kernel that reads the FiLM table.
  const bool chain_arm = cf.split_terms == 2 && D == 256 && (T % 8 == 0) && T >= 128 && !chain_disabled();
  if (b0 != 0 && !chain_arm) A2P_FAIL("forward: a batch-row offset needs the fused arm");
  const bool side = chain_arm && !c.prof && !getenv("A2P_NO_SIDE_STREAM");
  cudaStream_t st_main = st;
  if (side) {
    if (!h->cond_stream[c.slot]) A2P_CUDA(cudaStreamCreateWithFlags(&h->cond_stream[c.slot], cudaStreamNonBlocking));
    if (!h->ev_fork[c.slot]) A2P_CUDA(cudaEventCreateWithFlags(&h->ev_fork[c.slot], cudaEventDisableTiming));
    if (!h->ev_join[c.slot]) A2P_CUDA(cudaEventCreateWithFlags(&h->ev_join[c.slot], cudaEventDisableTiming));
    A2P_CUDA(cudaEventRecord(h->ev_fork[c.slot], st_main));
    A2P_CUDA(cudaStreamWaitEvent(h->cond_stream[c.slot], h->ev_fork[c.slot], 0));
    st = h->cond_stream[c.slot]; c.st = st;
  }
  cudaStream_t st_side = side ? h->cond_stream[c.slot] : st_main;
  c.cat = CAT_COND;
  c.skinny = true;
  time_embed_kernel<<<ceil_div(R * D / 2, 256), 256, 0, st>>>(ts, counter, B, R, D, h->time_freqs, e);
  h->launches++;
  A2P_TRY(gemm(c, e, D, R, h->time_w1, D, h->time_b1, 4 * D, D, th, 4 * D, EPI_MISH));
  {
    GemmParams ex{};
    ex.rowvec.base[0] = c0.hidden + (c0.Bc == 1 ? 0 : (size_t)b0 * D); ex.rowvec.stride[0] = c0.Bc == 1 ? 0 : D;
    ex.rowvec.base[1] = c1.hidden ? c1.hidden + (c1.Bc == 1 ? 0 : (size_t)b0 * D) : nullptr; ex.rowvec.stride[1] = c1.Bc == 1 ? 0 : D;
    ex.rowvec.rows_per_branch = B;
    A2P_TRY(gemm(c, th, 4 * D, R, h->time_w2, 4 * D, h->time_b2, D, 4 * D, mt, D, EPI_ADDROW_MISH, &ex));
  }
  A2P_TRY(gemm(c, th, 4 * D, R, h->time_w3, 4 * D, h->time_b3, 2 * D, 4 * D, ttok, 2 * D));
  // time tokens -> norm_cond rows at positions S, S+1 (model/diffusion.py:392-393) -> per-layer K/V rows
  { int _c = c.cat; c.cat = CAT_LN; c.begin(); A2P_TRY(launch_ln_rope(D, ttok, D, h->normc_w, h->normc_b, tt, ttr, D, h->rope_tab, 2, S, 2 * R, st)); c.end(); c.cat = _c; }
  h->launches++;
  A2P_TRY(gemm(c, ttr, D, 2 * R, h->ttk_w, D, h->ttk_b, L * D, D, ktt, (long long)L * D));
  A2P_TRY(gemm(c, tt, D, 2 * R, h->ttv_w, D, h->ttv_b, L * D, D, vtt, (long long)L * D));
  // all FiLM (scale, shift) pairs of all layers in one GEMM (transformer_modules.py:105-119)
  A2P_TRY(gemm(c, mt, D, R, h->film_w, D, h->film_b, (int)film_ld, D, film, film_ld));
  // --- input projection (identical for both branches: computed once, duplicated)
  c.cat = CAT_IO_TCN;
  c.skinny = false;
  st = st_main; c.st = st_main;
  if (chain_arm) {   // input projection fused into the first chain launch below
  } else if (cf.split_terms > 0) {
    __nv_bfloat16* xinP = reinterpret_cast<__nv_bfloat16*>(wsb + w.xinP);
    A2P_TRY(launch_split_planes(cf.split_terms, xin, C, xinP, (long long)B * T * C, (long long)B * T, C, 1.f, st));
    h->launches++;
    TcGemmParams g{};
    g.C = x; g.ldc = D;
    A2P_TRY(tc_gemm(c, xinP, B * T, C, h->inp_w, 0, D, h->inp_b, TC_F32, g));
  } else {
    A2P_TRY(gemm(c, xin, C, B * T, h->inp_w, C, h->inp_b, D, C, x, D));
  }
  if (nb == 2 && !chain_arm) {
    A2P_CUDA(cudaMemcpyAsync(x + (size_t)B * T * D, x, sizeof(float) * (size_t)B * T * D, cudaMemcpyDeviceToDevice, st));
  }
  const float scale_log2e = (1.0f / sqrtf((float)dh)) * 1.4426950408889634f;
  const long long sT = (long long)T * D;
  const int P = cf.split_terms;
  __nv_bfloat16* hP = P ? reinterpret_cast<__nv_bfloat16*>(wsb + w.hP) : nullptr;
  __nv_bfloat16* hrP = P ? reinterpret_cast<__nv_bfloat16*>(wsb + w.hrP) : nullptr;
  __nv_bfloat16* attP = P ? reinterpret_cast<__nv_bfloat16*>(wsb + w.attP) : nullptr;
  __nv_bfloat16* uP = P ? reinterpret_cast<__nv_bfloat16*>(wsb + w.uP) : nullptr;
  const int MT = R * T;
  const long long pstrideD = (long long)MT * D;
  // ---- tensor-core attention operands (needs T % 8 == 0 for the 16-byte TMA strides of the V^T planes)
  const bool tc_attn = P > 0 && (T % 8 == 0);
  __nv_bfloat16* qkP = P ? reinterpret_cast<__nv_bfloat16*>(wsb + w.qkP) : nullptr;
  __nv_bfloat16* vtS = P ? reinterpret_cast<__nv_bfloat16*>(wsb + w.vtS) : nullptr;
  __nv_bfloat16* kttP = P ? reinterpret_cast<__nv_bfloat16*>(wsb + w.kttP) : nullptr;
  __nv_bfloat16* vttT = P ? reinterpret_cast<__nv_bfloat16*>(wsb + w.vttT) : nullptr;
  const long long MT8 = (long long)align_up((size_t)MT, 8);
  const long long XP = 8LL * R;   // time-token V^T: 8 columns per sample (TMA needs 16-byte aligned box starts), 2 used
  if (tc_attn) {
    if (side) { st = st_side; c.st = st; }
    c.cat = CAT_COND;
    c.begin();
    A2P_TRY(launch_split_planes(P, ktt, (long long)L * D, kttP, (long long)2 * R * L * D, 2 * R, L * D, 1.f, st));
    A2P_CUDA(cudaMemsetAsync(vttT, 0, sizeof(__nv_bfloat16) * (size_t)P * L * D * XP, st));
    A2P_TRY(launch_transpose_split(P, vtt, (long long)L * D, vttT, (long long)L * D * XP, XP, 2 * R, L * D, 2, 8, 1.f, st));
    c.end();
    h->launches += 3;
    if (side) {
      A2P_CUDA(cudaEventRecord(h->ev_join[c.slot], st_side));
      st = st_main; c.st = st_main;
    }
  }
  auto attn_tc = [&](int l, int kind) -> int {   // kind 0 self, 1 audio cross (+2 time tokens), 2 keyframe cross
    TcAttnOperands o{};
    TcAttnParams ap{};
    o.Q = qkP; o.q_rows = MT; o.q_ld = 2 * D; o.q_plane_stride = (long long)MT * 2 * D;
    o.vt_rows = D;
    ap.skew_ns = h->attn_skew_ns;
    if (P == 2) { ap.split_scratch = F(w.splitS); ap.split_counters = reinterpret_cast<int*>(wsb + w.splitC); }
    ap.T = T; ap.R = R; ap.D = D; ap.dh = dh; ap.q_col0 = 0; ap.Op = attP; ap.op_plane_stride = pstrideD; ap.o_ld = D; ap.O = nullptr;
    if (kind == 0) {
      o.K[0] = qkP; o.k_rows[0] = MT; o.k_ld[0] = 2 * D; o.k_plane_stride[0] = (long long)MT * 2 * D;
      o.Vt[0] = vtS; o.vt_cols[0] = MT; o.vt_ld[0] = MT8; o.vt_plane_stride[0] = (long long)D * MT8;
      ap.rows_per_branch = R; ap.k_col0 = D; ap.n_keys = T; ap.n_extra = 0; ap.k_row_stride[0] = T; ap.v_col_stride[0] = T;
    } else {
      const CondSet* cs[2] = {&c0, &c1};
      const KvLayout* kl[2] = {&k0, &k1};
      for (int b = 0; b < nb; ++b) {
        float* base = cs[b]->base + kl[b]->per_layer * l;
        const long long Bc = cs[b]->Bc;
        const long long bo = Bc == 1 ? 0 : b0;   // first sample of this call inside the per-sample caches
        if (kind == 1) {
          const long long Sp = kl[b]->Sp;
          o.K[b] = reinterpret_cast<__nv_bfloat16*>(base + kl[b]->kaP) + bo * S * D; o.k_rows[b] = (Bc - bo) * S; o.k_ld[b] = D; o.k_plane_stride[b] = Bc * S * D;
          o.Vt[b] = reinterpret_cast<__nv_bfloat16*>(base + kl[b]->vtaP) + bo * Sp; o.vt_cols[b] = (Bc - bo) * Sp; o.vt_ld[b] = Bc * Sp;
          o.vt_plane_stride[b] = (long long)D * Bc * Sp;
          ap.k_row_stride[b] = Bc == 1 ? 0 : S; ap.v_col_stride[b] = Bc == 1 ? 0 : Sp;
        } else {
          const long long S2p = kl[b]->S2p;
          o.K[b] = reinterpret_cast<__nv_bfloat16*>(base + kl[b]->k2P) + bo * S2 * D; o.k_rows[b] = (Bc - bo) * S2; o.k_ld[b] = D; o.k_plane_stride[b] = Bc * S2 * D;
          o.Vt[b] = reinterpret_cast<__nv_bfloat16*>(base + kl[b]->vt2P) + bo * S2p; o.vt_cols[b] = (Bc - bo) * S2p; o.vt_ld[b] = Bc * S2p;
          o.vt_plane_stride[b] = (long long)D * Bc * S2p;
          ap.k_row_stride[b] = Bc == 1 ? 0 : S2; ap.v_col_stride[b] = Bc == 1 ? 0 : S2p;
        }
      }
      ap.rows_per_branch = B; ap.k_col0 = 0; ap.n_keys = kind == 1 ? S : S2; ap.n_extra = 0;
      if (kind == 1) {
        o.Kx = kttP; o.kx_rows = 2 * R; o.kx_ld = (long long)L * D; o.kx_plane_stride = (long long)2 * R * L * D;
        o.Vx = vttT; o.vx_rows = (long long)L * D; o.vx_cols = XP; o.vx_ld = XP; o.vx_plane_stride = (long long)L * D * XP;
        ap.n_extra = 2; ap.kx_col0 = l * D; ap.kx_row_stride = 2; ap.vx_row0 = l * D; ap.vx_col_stride = 8;
      }
    }
    c.cat = kind == 0 ? CAT_ATT_SELF : (kind == 1 ? CAT_ATT_CROSS : CAT_ATT_CROSS2);
    c.begin();
    const int av = attn2_variant();
    int rc;
    if (P == 2 && av > 0 && kind == 2 && attn_short_ok(ap) && !attn_short_disabled()) rc = launch_umma_attn_short(o, ap, st);
    else rc = (P == 2 && dh == 32 && av > 0) ? launch_umma_attn2(av, o, ap, st) : launch_umma_attn(P, o, ap, st);
    c.end();
    c.cat = CAT_PROJ;
    h->launches++;
    if (rc == 0) rc = c.mark();
    return rc;
  };
  // ===== fused row-chain arm (split_terms == 2, D == 256): per layer 4 chain launches + 3 attention launches =====
  const bool chain = P == 2 && D == 256 && tc_attn && T >= 128 && !chain_disabled();
  if (chain) {
    A2P_CUDA(cudaMemsetAsync(wsb + w.splitC, 0, attn2_split_counter_ints() * sizeof(int), st));   // arrival counters of the split-KV tail
    float* ropeX = F(w.ropeX);
    rope_ext_kernel<<<ceil_div((T + 128) * (D / 2), 256), 256, 0, st>>>(h->rope_tab, reinterpret_cast<float2*>(ropeX), T, D / 2);
    h->launches++;
    auto planes = [&](const float* key, long long row0, long long cols, const __nv_bfloat16** base, long long* pstride) -> int {
      auto it = h->wplanes.find(key);
      if (it == h->wplanes.end()) A2P_FAIL("chain: weight has no split planes");
      *base = it->second + row0 * cols; *pstride = h->wnumel[key];
      return 0;
    };
    // GEMM0 = A0 * W0^T (+ FiLM / residual), then LayerNorm (+RoPE), then GEMM1 (and optionally the V^T job of a self-attention)
    struct Next { const float* lnw; const float* lnb; int rope; const float* w1; long long w1_row0; int N1; const float* b1;
                  float oscale; int scale_ncols; int gelu; __nv_bfloat16* Cp; long long cp_ps, ldcp; int remap_rps, remap_pad;
                  const float* w2; long long w2_row0; const float* b2; };
    // residual stream: in place in `x`, except for N-split launches with a FiLM / residual update, which read the current
    // buffer and write the other one (the LayerNorm scratch of the unfused arm is free here)
    float *xcur = x, *xalt = hh;
    auto run_chain = [&](int cat, const char* tag, const __nv_bfloat16* A0, long long a0_rows, int K0, const float* w0, const float* b0,
                         int film_off, const Next& nx) -> int {
      ChainOperands o{};
      ChainParams cp{};
      o.A0 = A0; o.a0_rows = a0_rows; o.a0_ld = K0; o.a0_plane_stride = a0_rows * K0;
      A2P_TRY(planes(w0, 0, K0, &o.W0, &o.w0_plane_stride));
      A2P_TRY(planes(nx.w1, nx.w1_row0, D, &o.W1, &o.w1_plane_stride));
      if (nx.w2) A2P_TRY(planes(nx.w2, nx.w2_row0, D, &o.W2, &o.w2_plane_stride));
      cp.M = MT; cp.T = T; cp.K0 = K0; cp.bias0 = b0;
      cp.film_mode = film_off >= 0 ? 1 : 0; cp.film = film; cp.film_ld = film_ld; cp.film_scale_off = film_off; cp.film_shift_off = film_off + D;
      o.x = xcur; o.rope_ext = ropeX; o.rope_ext_rows = T + 128;
      cp.nsplit = chain_nsplit_for(ceil_div(MT, 128), ceil_div(nx.N1, 128) + (nx.w2 ? 2 : 0), c.concurrent, K0);
      if (cp.nsplit > 1 && cp.film_mode) { o.x_out = xalt; std::swap(xcur, xalt); }
      cp.ln_mode = nx.lnw ? 1 : 0; cp.ln_w = nx.lnw; cp.ln_b = nx.lnb; cp.rope = nx.rope;
      cp.N1 = nx.N1; cp.bias1 = nx.b1; cp.out_scale = nx.oscale; cp.scale_ncols = nx.scale_ncols; cp.gelu = nx.gelu;
      cp.Cp = nx.Cp; cp.cp_plane_stride = nx.cp_ps; cp.ldcp = nx.ldcp; cp.remap_rps = nx.remap_rps; cp.remap_pad = nx.remap_pad;
      cp.vjob = nx.w2 ? 1 : 0; cp.bias2 = nx.b2; cp.Vt = vtS; cp.vt_plane_stride = (long long)D * MT8; cp.ldvt = MT8;
      c.cat = cat;
      if (c.prof) c.tag = tag;
      c.begin();
      launch_priority() = chain_priority();
      int rc = launch_umma_chain(o, cp, st);
      launch_priority() = 0;
      c.end();
      h->launches++;
      if (rc == 0) rc = c.mark();
      return rc;
    };
    auto next_self = [&](int l) {   // LN1 + RoPE -> Q|K planes (Q pre-scaled), un-rotated LN1 -> V^T planes
      const LayerW& lw = h->lw[l];
      return Next{lw.n1w, lw.n1b, 1, lw.sa.in_w, 0, 2 * D, lw.sa.in_b, scale_log2e, D, 0, qkP, (long long)MT * 2 * D, 2 * D, 0, 0,
                  lw.sa.in_w, 2 * D, lw.sa.in_b + 2 * D};
    };
    auto next_q = [&](const float* nw, const float* nb_, const AttnW& a) {   // LN + RoPE -> Q planes of a cross attention
      return Next{nw, nb_, 1, a.in_w, 0, D, a.in_b, scale_log2e, 0, 0, qkP, (long long)MT * 2 * D, 2 * D, 0, 0, nullptr, 0, nullptr};
    };
    // input projection (both CFG branches see the same x_t: the planes are duplicated) -> layer 0 self-attention operands
    __nv_bfloat16* xinP = reinterpret_cast<__nv_bfloat16*>(wsb + w.xinP);
    c.cat = CAT_IO_TCN;
    for (int b = 0; b < nb; ++b)
      A2P_TRY(launch_split_planes(2, xin, C, xinP + (size_t)b * B * T * C, (long long)MT * C, (long long)B * T, C, 1.f, st));
    h->launches += nb;
    A2P_TRY(run_chain(CAT_IO_TCN, "chain in_proj->ln1->qkv", xinP, MT, C, h->inp_w, h->inp_b, -1, next_self(0)));
    for (int l = 0; l < L; ++l) {
      const LayerW& lw = h->lw[l];
      const int fo = l * nf * 2 * D;
      A2P_TRY(attn_tc(l, 0));
      if (l == 0 && side) A2P_CUDA(cudaStreamWaitEvent(st, h->ev_join[c.slot], 0));   // FiLM table + time-token K/V rows are ready
      A2P_TRY(run_chain(CAT_PROJ, "chain sa_out->ln2->q", attP, MT, D, lw.sa.out_w, lw.sa.out_b, fo + 0 * 2 * D, next_q(lw.n2w, lw.n2b, lw.ca)));
      A2P_TRY(attn_tc(l, 1));
      const AttnW* last = &lw.ca;
      int fidx = 1;
      if (cf.fmt == A2P_FMT_POSE) {
        A2P_TRY(run_chain(CAT_PROJ, "chain ca_out->ln2a->q", attP, MT, D, lw.ca.out_w, lw.ca.out_b, fo + 1 * 2 * D, next_q(lw.n2aw, lw.n2ab, lw.c2)));
        A2P_TRY(attn_tc(l, 2));
        last = &lw.c2; fidx = 2;
      }
      {
        Next nx{lw.n3w, lw.n3b, 0, lw.l1w, 0, cf.FF, lw.l1b, 1.f, 0, 1, uP, (long long)MT * cf.FF, cf.FF, 0, 0, nullptr, 0, nullptr};
        A2P_TRY(run_chain(CAT_FFN, "chain out->ln3->ffn1", attP, MT, D, last->out_w, last->out_b, fo + fidx * 2 * D, nx));
      }
      if (l + 1 < L) {
        A2P_TRY(run_chain(CAT_FFN, "chain ffn2->ln1->qkv", uP, MT, cf.FF, lw.l2w, lw.l2b, fo + (nf - 1) * 2 * D, next_self(l + 1)));
      } else if (cf.fmt == A2P_FMT_POSE) {
        // last layer: FFN2 + FiLM + residual, then final_layer straight into the left-padded TCN input planes
        const int Pl = T + TCN_PAD, Mp = R * Pl;
        __nv_bfloat16* U = reinterpret_cast<__nv_bfloat16*>(wsb + w.tcnUP);
        A2P_CUDA(cudaMemset2DAsync(U, (size_t)Pl * C * 2, 0, (size_t)TCN_PAD * C * 2, (size_t)P * R, st));
        h->launches++;
        Next nx{nullptr, nullptr, 0, h->fin_w, 0, C, h->fin_b, 1.f, 0, 0, U, (long long)Mp * C, C, T, TCN_PAD, nullptr, 0, nullptr};
        A2P_TRY(run_chain(CAT_FFN, "chain ffn2->final_layer", uP, MT, cf.FF, lw.l2w, lw.l2b, fo + (nf - 1) * 2 * D, nx));
      } else {
        A2P_FAIL("chain arm: face models are not supported (D must be 256)");
      }
    }
  }
  for (int l = 0; P > 0 && !chain && l < L; ++l) {
    // ===== tensor-core arm: every [R*T, *] linear runs as a split-bf16 tcgen05 GEMM; LN / RoPE / softmax stay fp32 =====
    const LayerW& lw = h->lw[l];
    const int fo = l * nf * 2 * D;
    auto lnp = [&](const float* nw, const float* nb_, __nv_bfloat16* oh, __nv_bfloat16* orr) -> int {
      int _c = c.cat; c.cat = CAT_LN; c.begin();
      int rc = launch_ln_rope_planes(D, P, x, D, nw, nb_, oh, orr, pstrideD, h->rope_tab, T, 0, MT, st);
      c.end(); c.cat = _c; h->launches++;
      return rc;
    };
    auto attn = [&](AttnParams& a, int cat) -> int {   // exact-fp32 attention core (used when T % 8 != 0)
      a.Q = qkv; a.q_ld = 3 * D; a.q_sample_stride = 3 * sT;
      a.O = nullptr; a.Op = attP; a.op_plane_stride = pstrideD; a.op_terms = P; a.o_ld = D; a.o_sample_stride = sT;
      a.T = T; a.H = H; a.R = R; a.scale_log2e = scale_log2e;
      c.cat = cat; c.begin();
      int rc = launch_attn_simt(a, dh, st);
      c.end(); c.cat = CAT_PROJ; h->launches++;
      return rc;
    };
    TcGemmParams f32out{};
    f32out.C = qkv; f32out.ldc = 3 * D;
    TcGemmParams qplanes{};
    qplanes.Cp = qkP; qplanes.cp_plane_stride = (long long)MT * 2 * D; qplanes.ldcp = 2 * D; qplanes.out_scale = scale_log2e;
    c.cat = CAT_PROJ;
    // ---- self attention
    A2P_TRY(lnp(lw.n1w, lw.n1b, hP, hrP));
    if (tc_attn) {
      { TcGemmParams q = qplanes; q.scale_ncols = D; A2P_TRY(tc_gemm(c, hrP, MT, D, lw.sa.in_w, 0, 2 * D, lw.sa.in_b, TC_PLANES, q)); }
      {  // V^T = Wv * LN(x)^T  (swapped operands: the result is already transposed for the PV product)
        TcOperands o{h->wplanes[lw.sa.in_w] + (size_t)2 * D * D, D, h->wnumel[lw.sa.in_w], hP, D, pstrideD};
        TcGemmParams v{};
        v.M = D; v.N = MT; v.K = D; v.taps = 1; v.bias = lw.sa.in_b + 2 * D; v.bias_per_row = 1; v.out_scale = 1.f;
        v.Cp = vtS; v.cp_plane_stride = (long long)D * MT8; v.ldcp = MT8;
        h->launches++;
        c.begin();
        A2P_TRY(launch_umma_gemm(P, o, v, TC_PLANES, h->num_sms, st));
        c.end();
      }
      A2P_TRY(attn_tc(l, 0));
    } else {
      A2P_TRY(tc_gemm(c, hrP, MT, D, lw.sa.in_w, 0, 2 * D, lw.sa.in_b, TC_F32, f32out));
      { TcGemmParams v = f32out; v.C = qkv + 2 * D; A2P_TRY(tc_gemm(c, hP, MT, D, lw.sa.in_w, 2 * D, D, lw.sa.in_b + 2 * D, TC_F32, v)); }
      AttnParams a{};
      a.K.base[0] = qkv + D; a.K.stride[0] = 3 * sT; a.K.base[1] = nullptr; a.K.stride[1] = 0; a.K.rows_per_branch = R;
      a.V = a.K; a.V.base[0] = qkv + 2 * D;
      a.kv_ld = 3 * D; a.S_main = T; a.S_extra = 0;
      A2P_TRY(attn(a, CAT_ATT_SELF));
    }
    A2P_TRY(tc_film_gemm(c, attP, MT, D, lw.sa.out_w, lw.sa.out_b, D, x, D, film, film_ld, fo + 0 * 2 * D, D, T));
    // ---- audio cross attention
    A2P_TRY(lnp(lw.n2w, lw.n2b, nullptr, hrP));
    if (tc_attn) {
      A2P_TRY(tc_gemm(c, hrP, MT, D, lw.ca.in_w, 0, D, lw.ca.in_b, TC_PLANES, qplanes));
      A2P_TRY(attn_tc(l, 1));
    } else {
      A2P_TRY(tc_gemm(c, hrP, MT, D, lw.ca.in_w, 0, D, lw.ca.in_b, TC_F32, f32out));
      AttnParams a{};
      a.K.base[0] = c0.base + k0.per_layer * l + k0.ka; a.K.stride[0] = c0.Bc == 1 ? 0 : (long long)S * D;
      a.K.base[1] = c1.base ? c1.base + k1.per_layer * l + k1.ka : nullptr; a.K.stride[1] = c1.Bc == 1 ? 0 : (long long)S * D;
      a.K.rows_per_branch = B;
      a.V = a.K;
      a.V.base[0] = c0.base + k0.per_layer * l + k0.va;
      a.V.base[1] = c1.base ? c1.base + k1.per_layer * l + k1.va : nullptr;
      a.kv_ld = D; a.S_main = S;
      a.Kx = ktt + (size_t)l * D; a.Vx = vtt + (size_t)l * D; a.x_ld = (long long)L * D; a.x_sample_stride = 2LL * L * D; a.S_extra = 2;
      A2P_TRY(attn(a, CAT_ATT_CROSS));
    }
    A2P_TRY(tc_film_gemm(c, attP, MT, D, lw.ca.out_w, lw.ca.out_b, D, x, D, film, film_ld, fo + 1 * 2 * D, D, T));
    // ---- keyframe cross attention (pose)
    if (cf.fmt == A2P_FMT_POSE) {
      A2P_TRY(lnp(lw.n2aw, lw.n2ab, nullptr, hrP));
      if (tc_attn) {
        A2P_TRY(tc_gemm(c, hrP, MT, D, lw.c2.in_w, 0, D, lw.c2.in_b, TC_PLANES, qplanes));
        A2P_TRY(attn_tc(l, 2));
      } else {
        A2P_TRY(tc_gemm(c, hrP, MT, D, lw.c2.in_w, 0, D, lw.c2.in_b, TC_F32, f32out));
        AttnParams a{};
        a.K.base[0] = c0.base + k0.per_layer * l + k0.k2; a.K.stride[0] = c0.Bc == 1 ? 0 : (long long)S2 * D;
        a.K.base[1] = c1.base ? c1.base + k1.per_layer * l + k1.k2 : nullptr; a.K.stride[1] = c1.Bc == 1 ? 0 : (long long)S2 * D;
        a.K.rows_per_branch = B;
        a.V = a.K;
        a.V.base[0] = c0.base + k0.per_layer * l + k0.v2;
        a.V.base[1] = c1.base ? c1.base + k1.per_layer * l + k1.v2 : nullptr;
        a.kv_ld = D; a.S_main = S2; a.S_extra = 0;
        A2P_TRY(attn(a, CAT_ATT_CROSS2));
      }
      A2P_TRY(tc_film_gemm(c, attP, MT, D, lw.c2.out_w, lw.c2.out_b, D, x, D, film, film_ld, fo + 2 * 2 * D, D, T));
    }
    // ---- feed forward: GELU output goes straight to split planes
    c.cat = CAT_FFN;
    A2P_TRY(lnp(lw.n3w, lw.n3b, hP, nullptr));
    { TcGemmParams g{}; g.Cp = uP; g.cp_plane_stride = (long long)MT * cf.FF; g.ldcp = cf.FF;
      A2P_TRY(tc_gemm(c, hP, MT, D, lw.l1w, 0, cf.FF, lw.l1b, TC_GELU_PLANES, g)); }
    A2P_TRY(tc_film_gemm(c, uP, MT, cf.FF, lw.l2w, lw.l2b, D, x, D, film, film_ld, fo + (nf - 1) * 2 * D, D, T));
  }
  for (int l = 0; P == 0 && l < L; ++l) {
    const LayerW& lw = h->lw[l];
    const int fo = l * nf * 2 * D;
    // ---- self attention: q = k = rot(LN1 x), v = LN1 x  (transformer_modules.py:237-247)
    c.cat = CAT_PROJ;
    { int _c = c.cat; c.cat = CAT_LN; c.begin(); A2P_TRY(launch_ln_rope(D, x, D, lw.n1w, lw.n1b, hh, hr, D, h->rope_tab, T, 0, R * T, st)); c.end(); c.cat = _c; }
    h->launches++;
    A2P_TRY(gemm(c, hr, D, R * T, lw.sa.in_w, D, lw.sa.in_b, 2 * D, D, qkv, 3 * D));
    A2P_TRY(gemm(c, hh, D, R * T, lw.sa.in_w + (size_t)2 * D * D, D, lw.sa.in_b + 2 * D, D, D, qkv + 2 * D, 3 * D));
    {
      AttnParams a{};
      a.Q = qkv; a.q_ld = 3 * D; a.q_sample_stride = 3 * sT;
      a.K.base[0] = qkv + D; a.K.stride[0] = 3 * sT; a.K.base[1] = nullptr; a.K.stride[1] = 0; a.K.rows_per_branch = R;
      a.V = a.K; a.V.base[0] = qkv + 2 * D;
      a.kv_ld = 3 * D; a.S_main = T; a.S_extra = 0;
      a.O = att; a.o_ld = D; a.o_sample_stride = sT; a.T = T; a.H = H; a.R = R; a.scale_log2e = scale_log2e;
      c.cat = CAT_ATT_SELF;
      c.begin(); A2P_TRY(launch_attn_simt(a, dh, st)); c.end();
      c.cat = CAT_PROJ;
      h->launches++;
    }
    A2P_TRY(film_gemm(c, att, D, R * T, lw.sa.out_w, lw.sa.out_b, D, D, x, D, film, film_ld, fo + 0 * 2 * D, D, T));
    // ---- cross attention over [audio tokens | 2 time tokens]  (transformer_modules.py:251-262)
    { int _c = c.cat; c.cat = CAT_LN; c.begin(); A2P_TRY(launch_ln_rope(D, x, D, lw.n2w, lw.n2b, nullptr, hr, D, h->rope_tab, T, 0, R * T, st)); c.end(); c.cat = _c; }
    h->launches++;
    A2P_TRY(gemm(c, hr, D, R * T, lw.ca.in_w, D, lw.ca.in_b, D, D, qkv, 3 * D));
    {
      AttnParams a{};
      a.Q = qkv; a.q_ld = 3 * D; a.q_sample_stride = 3 * sT;
      a.K.base[0] = c0.base + k0.per_layer * l + k0.ka; a.K.stride[0] = c0.Bc == 1 ? 0 : (long long)S * D;
      a.K.base[1] = c1.base ? c1.base + k1.per_layer * l + k1.ka : nullptr; a.K.stride[1] = c1.Bc == 1 ? 0 : (long long)S * D;
      a.K.rows_per_branch = B;
      a.V = a.K;
      a.V.base[0] = c0.base + k0.per_layer * l + k0.va;
      a.V.base[1] = c1.base ? c1.base + k1.per_layer * l + k1.va : nullptr;
      a.kv_ld = D; a.S_main = S;
      a.Kx = ktt + (size_t)l * D; a.Vx = vtt + (size_t)l * D; a.x_ld = (long long)L * D; a.x_sample_stride = 2LL * L * D; a.S_extra = 2;
      a.O = att; a.o_ld = D; a.o_sample_stride = sT; a.T = T; a.H = H; a.R = R; a.scale_log2e = scale_log2e;
      c.cat = CAT_ATT_CROSS;
      c.begin(); A2P_TRY(launch_attn_simt(a, dh, st)); c.end();
      c.cat = CAT_PROJ;
      h->launches++;
    }
    A2P_TRY(film_gemm(c, att, D, R * T, lw.ca.out_w, lw.ca.out_b, D, D, x, D, film, film_ld, fo + 1 * 2 * D, D, T));
    // ---- cross attention over keyframe tokens (pose; transformer_modules.py:204-214)
    if (cf.fmt == A2P_FMT_POSE) {
      { int _c = c.cat; c.cat = CAT_LN; c.begin(); A2P_TRY(launch_ln_rope(D, x, D, lw.n2aw, lw.n2ab, nullptr, hr, D, h->rope_tab, T, 0, R * T, st)); c.end(); c.cat = _c; }
      h->launches++;
      A2P_TRY(gemm(c, hr, D, R * T, lw.c2.in_w, D, lw.c2.in_b, D, D, qkv, 3 * D));
      AttnParams a{};
      a.Q = qkv; a.q_ld = 3 * D; a.q_sample_stride = 3 * sT;
      a.K.base[0] = c0.base + k0.per_layer * l + k0.k2; a.K.stride[0] = c0.Bc == 1 ? 0 : (long long)S2 * D;
      a.K.base[1] = c1.base ? c1.base + k1.per_layer * l + k1.k2 : nullptr; a.K.stride[1] = c1.Bc == 1 ? 0 : (long long)S2 * D;
      a.K.rows_per_branch = B;
      a.V = a.K;
      a.V.base[0] = c0.base + k0.per_layer * l + k0.v2;
      a.V.base[1] = c1.base ? c1.base + k1.per_layer * l + k1.v2 : nullptr;
      a.kv_ld = D; a.S_main = S2; a.S_extra = 0;
      a.O = att; a.o_ld = D; a.o_sample_stride = sT; a.T = T; a.H = H; a.R = R; a.scale_log2e = scale_log2e;
      c.cat = CAT_ATT_CROSS2;
      c.begin(); A2P_TRY(launch_attn_simt(a, dh, st)); c.end();
      c.cat = CAT_PROJ;
      h->launches++;
      A2P_TRY(film_gemm(c, att, D, R * T, lw.c2.out_w, lw.c2.out_b, D, D, x, D, film, film_ld, fo + 2 * 2 * D, D, T));
    }
    // ---- feed forward (transformer_modules.py:265-267)
    c.cat = CAT_FFN;
    { int _c = c.cat; c.cat = CAT_LN; c.begin(); A2P_TRY(launch_ln_rope(D, x, D, lw.n3w, lw.n3b, hh, nullptr, D, h->rope_tab, T, 0, R * T, st)); c.end(); c.cat = _c; }
    h->launches++;
    A2P_TRY(gemm(c, hh, D, R * T, lw.l1w, D, lw.l1b, cf.FF, D, u, cf.FF, EPI_GELU));
    A2P_TRY(film_gemm(c, u, cf.FF, R * T, lw.l2w, lw.l2b, D, cf.FF, x, D, film, film_ld, fo + (nf - 1) * 2 * D, D, T));
  }
  // ---- final projection (+ causal TCN for pose; model/diffusion.py:397-402)
  c.cat = CAT_IO_TCN;
  if (P > 0 && cf.fmt == A2P_FMT_POSE) {
    // tensor-core arm of final_layer + causal TCN: activations travel as split planes in the left-padded layout
    // [R][T+24][C]; conv tap j reads rows shifted by (2-j)*dil (TMA zero-fills rows < 0); fp32 copies feed the skips.
    const int cm = C > 256 ? C : 256;
    const int Pl = T + TCN_PAD, Mp = R * Pl;
    __nv_bfloat16* U = reinterpret_cast<__nv_bfloat16*>(wsb + w.tcnUP);
    __nv_bfloat16* V = reinterpret_cast<__nv_bfloat16*>(wsb + w.tcnVP);
    float *X = F(w.tcnA), *Y = F(w.tcnC);
    c.cat = CAT_IO_TCN;
    if (!chain) {
      A2P_TRY(launch_split_planes(P, x, D, hP, pstrideD, MT, D, 1.f, st));
      A2P_CUDA(cudaMemset2DAsync(U, (size_t)Pl * C * 2, 0, (size_t)TCN_PAD * C * 2, (size_t)P * R, st));
      h->launches += 2;
      TcGemmParams g{};
      g.Cp = U; g.cp_plane_stride = (long long)Mp * C; g.ldcp = C; g.remap_rps = T; g.remap_pad = TCN_PAD;
      A2P_TRY(tc_gemm(c, hP, MT, D, h->fin_w, 0, C, h->fin_b, TC_PLANES, g));
    }
    struct St { __nv_bfloat16* in; int cin; __nv_bfloat16* outp; int cout; float* f32; const float* skip; };
    St stg[6] = {{U, C, V, cm, nullptr, nullptr}, {V, cm, U, C, X, nullptr}, {U, C, V, C, Y, X},
                 {V, C, U, C, X, Y},              {U, C, V, C, Y, X},        {V, C, U, C, nullptr, Y}};
    for (int i = 0; i < 6; ++i) {
      TcGemmParams g{};
      g.taps = 3; g.dil = TCN_DIL[i]; g.slope = 0.2f;
      g.Cp = stg[i].outp; g.cp_plane_stride = (long long)Mp * stg[i].cout; g.ldcp = stg[i].cout;
      g.C = stg[i].f32; g.ldc = stg[i].cout;
      g.skip = stg[i].skip; g.ldskip = stg[i].cout;
      A2P_TRY(tc_gemm(c, stg[i].in, Mp, stg[i].cin, h->conv_t[i], 0, stg[i].cout, h->conv_b[i], TC_LRELU_PLANES, g));
    }
    {
      TcGemmParams g{};
      g.C = Y; g.ldc = C;
      A2P_TRY(tc_gemm(c, U, Mp, C, h->fconv_w, 0, C, h->fconv_b, TC_F32, g));
    }
    *x0_sample_stride = (long long)Pl * C;
    const float* o0 = Y + (size_t)TCN_PAD * C;
    *x0_cond = (mask & A2P_MASK_COND) ? o0 : nullptr;
    *x0_uncond = (mask == A2P_MASK_BOTH) ? o0 + (size_t)B * Pl * C : (mask == A2P_MASK_UNCOND ? o0 : nullptr);
    return 0;
  }
  if (P > 0) {
    A2P_TRY(launch_split_planes(P, x, D, hP, pstrideD, MT, D, 1.f, st));
    h->launches++;
    TcGemmParams g{};
    g.C = out; g.ldc = C;
    A2P_TRY(tc_gemm(c, hP, MT, D, h->fin_w, 0, C, h->fin_b, TC_F32, g));
  } else {
    A2P_TRY(gemm(c, x, D, R * T, h->fin_w, D, h->fin_b, C, D, out, C));
  }
  if (cf.fmt == A2P_FMT_POSE) {
    const int cm = C > 256 ? C : 256;
    const int P = T + TCN_PAD;
    float *A = F(w.tcnA), *Bf = F(w.tcnB), *Cc = F(w.tcnC);
    long long total = (long long)R * P * (C / 4);
    pad_copy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(out, A, T, TCN_PAD, C / 4, total);
    h->launches++;
    struct Stage { const float* in; int cin; float* outp; int cout; bool skip; };
    Stage stg[6] = {{A, C, Bf, cm, false}, {Bf, cm, A, C, false}, {A, C, Cc, C, true},
                    {Cc, C, A, C, true},   {A, C, Cc, C, true},   {Cc, C, A, C, true}};
    for (int i = 0; i < 6; ++i) {
      GemmParams ex{};
      ex.taps = 3; ex.dil = TCN_DIL[i]; ex.Kc = stg[i].cin; ex.slope = 0.2f;
      ex.skip = stg[i].in; ex.ldskip = stg[i].cin;
      const bool skip = stg[i].skip && stg[i].cin == stg[i].cout;
      A2P_TRY(gemm(c, stg[i].in, stg[i].cin, R * P, h->conv_w[i], 3LL * stg[i].cin, h->conv_b[i], stg[i].cout, 3 * stg[i].cin,
                   stg[i].outp, stg[i].cout, skip ? EPI_LRELU_SKIPAVG : EPI_LRELU, &ex));
    }
    A2P_TRY(gemm(c, A, C, R * P, h->fconv_w, C, h->fconv_b, C, C, Cc, C));
    *x0_sample_stride = (long long)P * C;
    const float* o0 = Cc + (size_t)TCN_PAD * C;
    *x0_cond = (mask & A2P_MASK_COND) ? o0 : nullptr;
    *x0_uncond = (mask == A2P_MASK_BOTH) ? o0 + (size_t)B * P * C : (mask == A2P_MASK_UNCOND ? o0 : nullptr);
  } else {
    *x0_sample_stride = (long long)T * C;
    *x0_cond = (mask & A2P_MASK_COND) ? out : nullptr;
    *x0_uncond = (mask == A2P_MASK_BOTH) ? out + (size_t)B * T * C : (mask == A2P_MASK_UNCOND ? out : nullptr);
  }
  return 0;
}

int launch_k3(Ctx& c, K3Params p) {
  dim3 grid(ceil_div(p.T, 32), ceil_div(p.C, 32), p.B), block(32, 8);
  k3_sampler_kernel<<<grid, block, 0, c.st>>>(p);
  if (c.h) c.h->launches++;
  A2P_CUDA(cudaGetLastError());
  return 0;
}

int transpose_in(Ctx& c, const float* x, float* xin, int B, int C, int T) {
  dim3 grid(ceil_div(T, 32), ceil_div(C, 32), B), block(32, 8);
  bct_to_btc_kernel<<<grid, block, 0, c.st>>>(x, xin, C, T);
  c.h->launches++;
  A2P_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// =================================================================== C-ABI
extern "C" {

int a2p_abi_version(void) { return A2P_ABI_VERSION; }
const char* a2p_last_error(void) { return a2p::last_error().c_str(); }
int a2p_has_tcgen05(void) { return 1; }

int a2p_denoiser_create(a2p_denoiser_t** out, const a2p_model_cfg* cfg) {
  if (!out) A2P_FAIL("null out");
  A2P_TRY(check_cfg(cfg));
  int ndev = 0;
  A2P_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev <= 0) A2P_FAIL("no CUDA device: a2p_b200 has no CPU fallback");
  A2P_TRY(init_attn_simt());
  A2P_TRY(init_skinny_gemm());
  a2p_denoiser* h = new a2p_denoiser();
  h->cfg = *cfg;
  h->dh = cfg->D / cfg->H;
  h->nf = cfg->fmt == A2P_FMT_POSE ? 4 : 3;
  if (getenv("A2P_ATTN_SKEW_NS")) h->attn_skew_ns = atoi(getenv("A2P_ATTN_SKEW_NS"));
  *out = h;
  return 0;
}

void a2p_denoiser_destroy(a2p_denoiser_t* h) {
  if (!h) return;
  for (int i = 0; i < a2p_denoiser::MAXG; ++i) {
    if (h->gexec[i]) cudaGraphExecDestroy(h->gexec[i]);
    if (h->group_stream[i]) cudaStreamDestroy(h->group_stream[i]);
    if (h->ev_gdone[i]) cudaEventDestroy(h->ev_gdone[i]);
    if (h->ev_gfork[i]) cudaEventDestroy(h->ev_gfork[i]);
  }
  if (h->ev_gstart) cudaEventDestroy(h->ev_gstart);
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  for (int i = 0; i < a2p_denoiser::MAXU; ++i) {
    if (h->cond_stream[i]) cudaStreamDestroy(h->cond_stream[i]);
    if (h->ev_fork[i]) cudaEventDestroy(h->ev_fork[i]);
    if (h->ev_join[i]) cudaEventDestroy(h->ev_join[i]);
    if (h->unit_stream[i]) cudaStreamDestroy(h->unit_stream[i]);
    if (h->ev_ujoin[i]) cudaEventDestroy(h->ev_ujoin[i]);
  }
  if (h->ev_bfork) cudaEventDestroy(h->ev_bfork);
  delete h;
}

size_t a2p_packed_weight_bytes(const a2p_model_cfg* cfg) {
  if (check_cfg(cfg)) return 0;
  return packed_layout(*cfg).total;
}

int a2p_denoiser_bind_weights(a2p_denoiser_t* h, const a2p_weight_t* table, int n, void* packed, size_t packed_bytes,
                              void* stream) {
  if (!h || !table || !packed) A2P_FAIL("bind_weights: null argument");
  const a2p_model_cfg& cf = h->cfg;
  const PackedLayout pl = packed_layout(cf);
  if (packed_bytes < pl.total) A2P_FAIL("bind_weights: packed arena too small (%zu < %zu)", packed_bytes, pl.total);
  cudaStream_t st = (cudaStream_t)stream;
  std::map<std::string, std::pair<const float*, int64_t>> m;
  for (int i = 0; i < n; ++i) m[table[i].name] = {table[i].ptr, table[i].numel};
  std::string err;
  const int64_t D = cf.D, C = cf.C, FF = cf.FF;
  auto W = [&](const std::string& k, int64_t numel) { return find(m, k, numel, err); };
  h->time_w1 = W("time_mlp.1.weight", 4 * D * D); h->time_b1 = W("time_mlp.1.bias", 4 * D);
  h->time_w2 = W("to_time_cond.0.weight", 4 * D * D); h->time_b2 = W("to_time_cond.0.bias", D);
  h->time_w3 = W("to_time_tokens.0.weight", 8 * D * D); h->time_b3 = W("to_time_tokens.0.bias", 2 * D);
  h->normc_w = W("norm_cond.weight", D); h->normc_b = W("norm_cond.bias", D);
  h->inp_w = W("input_projection.weight", D * C); h->inp_b = W("input_projection.bias", D);
  h->fin_w = W("final_layer.weight", C * D); h->fin_b = W("final_layer.bias", C);
  h->time_freqs = W("a2p.time_freqs", D / 2);
  const float* freqs = W("rotary.freqs", D / 2);
  {   // conditioning encoders: optional (a table without them can still be driven through set_conditioning with caller-made tokens)
    auto WO = [&](const std::string& k, int64_t numel) -> const float* {
      auto it = m.find(k);
      return (it != m.end() && (numel < 0 || it->second.second == numel)) ? it->second.first : nullptr;
    };
    a2p_denoiser::EncW& ew = h->encw;
    ew = a2p_denoiser::EncW{};
    auto itc = m.find("cond_projection.weight");
    if (itc != m.end() && itc->second.second % D == 0) {
      ew.feat_dim = (int)(itc->second.second / D);
      ew.cp_w = itc->second.first; ew.cp_b = WO("cond_projection.bias", D);
      ew.p0w = WO("non_attn_cond_projection.0.weight", D); ew.p0b = WO("non_attn_cond_projection.0.bias", D);
      ew.p1w = WO("non_attn_cond_projection.1.weight", D * D); ew.p1b = WO("non_attn_cond_projection.1.bias", D);
      ew.p3w = WO("non_attn_cond_projection.3.weight", D * D); ew.p3b = WO("non_attn_cond_projection.3.bias", D);
      bool ok = ew.cp_b && ew.p0w && ew.p0b && ew.p1w && ew.p1b && ew.p3w && ew.p3b;
      if (cf.fmt == A2P_FMT_POSE) {
        ew.fp_w = WO("frame_cond_projection.weight", D * C); ew.fp_b = WO("frame_cond_projection.bias", D);
        ew.fn_w = WO("frame_norm_cond.weight", D); ew.fn_b = WO("frame_norm_cond.bias", D);
        ok = ok && ew.fp_w && ew.fp_b && ew.fn_w && ew.fn_b;
      } else {
        for (int i = 0; i < 2; ++i) {
          const std::string p = "cond_encoder." + std::to_string(i) + ".";
          a2p_denoiser::EncLayerW& e = ew.enc[i];
          e.n1w = WO(p + "norm1.weight", D); e.n1b = WO(p + "norm1.bias", D);
          e.n2w = WO(p + "norm2.weight", D); e.n2b = WO(p + "norm2.bias", D);
          e.in_w = WO(p + "self_attn.in_proj_weight", 3 * D * D); e.in_b = WO(p + "self_attn.in_proj_bias", 3 * D);
          e.out_w = WO(p + "self_attn.out_proj.weight", D * D); e.out_b = WO(p + "self_attn.out_proj.bias", D);
          e.l1w = WO(p + "linear1.weight", FF * D); e.l1b = WO(p + "linear1.bias", FF);
          e.l2w = WO(p + "linear2.weight", D * FF); e.l2b = WO(p + "linear2.bias", D);
          ok = ok && e.n1w && e.n1b && e.n2w && e.n2b && e.in_w && e.in_b && e.out_w && e.out_b && e.l1w && e.l1b && e.l2w && e.l2b;
        }
      }
      ew.ok = ok;
    }
  }
  char* pb = (char*)packed;
  auto P = [&](size_t off) { return reinterpret_cast<float*>(pb + off); };
  h->film_w = P(pl.film_w); h->film_b = P(pl.film_b); h->ttk_w = P(pl.ttk_w); h->ttk_b = P(pl.ttk_b);
  h->ttv_w = P(pl.ttv_w); h->ttv_b = P(pl.ttv_b); h->rope_tab = reinterpret_cast<float2*>(pb + pl.rope);
  const char* film_names_pose[4] = {"film1", "film2", "film2a", "film3"};
  const char* film_names_face[3] = {"film1", "film2", "film3"};
  for (int l = 0; l < cf.L; ++l) {
    const std::string p = "seqTransDecoder.stack." + std::to_string(l) + ".";
    LayerW& lw = h->lw[l];
    auto A = [&](const std::string& name, AttnW& a) {
      a.in_w = W(p + name + ".in_proj_weight", 3 * D * D); a.in_b = W(p + name + ".in_proj_bias", 3 * D);
      a.out_w = W(p + name + ".out_proj.weight", D * D); a.out_b = W(p + name + ".out_proj.bias", D);
    };
    A("self_attn", lw.sa); A("multihead_attn", lw.ca);
    if (cf.fmt == A2P_FMT_POSE) {
      A("multihead_attn2", lw.c2);
      lw.n2aw = W(p + "norm2a.weight", D); lw.n2ab = W(p + "norm2a.bias", D);
    }
    lw.l1w = W(p + "linear1.weight", FF * D); lw.l1b = W(p + "linear1.bias", FF);
    lw.l2w = W(p + "linear2.weight", D * FF); lw.l2b = W(p + "linear2.bias", D);
    lw.n1w = W(p + "norm1.weight", D); lw.n1b = W(p + "norm1.bias", D);
    lw.n2w = W(p + "norm2.weight", D); lw.n2b = W(p + "norm2.bias", D);
    lw.n3w = W(p + "norm3.weight", D); lw.n3b = W(p + "norm3.bias", D);
    if (!err.empty()) A2P_FAIL("bind_weights: %s", err.c_str());
    for (int f = 0; f < h->nf; ++f) {
      const std::string fn = p + (cf.fmt == A2P_FMT_POSE ? film_names_pose[f] : film_names_face[f]) + ".block.1.";
      const float* fw = W(fn + "weight", 2 * D * D);
      const float* fb = W(fn + "bias", 2 * D);
      if (!err.empty()) A2P_FAIL("bind_weights: %s", err.c_str());
      A2P_CUDA(cudaMemcpyAsync(h->film_w + ((size_t)l * h->nf + f) * 2 * D * D, fw, sizeof(float) * 2 * D * D, cudaMemcpyDeviceToDevice, st));
      A2P_CUDA(cudaMemcpyAsync(h->film_b + ((size_t)l * h->nf + f) * 2 * D, fb, sizeof(float) * 2 * D, cudaMemcpyDeviceToDevice, st));
    }
    // stacked K / V projections of the cross-attention (rows [D,2D) and [2D,3D) of in_proj) for the time tokens
    A2P_CUDA(cudaMemcpyAsync(h->ttk_w + (size_t)l * D * D, lw.ca.in_w + D * D, sizeof(float) * D * D, cudaMemcpyDeviceToDevice, st));
    A2P_CUDA(cudaMemcpyAsync(h->ttk_b + (size_t)l * D, lw.ca.in_b + D, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
    A2P_CUDA(cudaMemcpyAsync(h->ttv_w + (size_t)l * D * D, lw.ca.in_w + 2 * D * D, sizeof(float) * D * D, cudaMemcpyDeviceToDevice, st));
    A2P_CUDA(cudaMemcpyAsync(h->ttv_b + (size_t)l * D, lw.ca.in_b + 2 * D, sizeof(float) * D, cudaMemcpyDeviceToDevice, st));
  }
  if (cf.fmt == A2P_FMT_POSE) {
    const int64_t cm = C > 256 ? C : 256;
    const int64_t chans[6][2] = {{cm, C}, {C, cm}, {C, C}, {C, C}, {C, C}, {C, C}};
    for (int i = 0; i < 6; ++i) {
      const std::string p = "post_pose_layers." + std::to_string(i) + ".";
      const float* cw = W(p + "weight", chans[i][0] * chans[i][1] * 3);
      h->conv_b[i] = W(p + "bias", chans[i][0]);
      if (!err.empty()) A2P_FAIL("bind_weights: %s", err.c_str());
      h->conv_w[i] = P(pl.conv_w[i]);
      h->conv_t[i] = P(pl.conv_t[i]);
      int tot = (int)(chans[i][0] * chans[i][1] * 3);
      permute_conv_w_kernel<<<ceil_div(tot, 256), 256, 0, st>>>(cw, h->conv_w[i], (int)chans[i][0], (int)chans[i][1]);
      permute_conv_w_tapmajor_kernel<<<ceil_div(tot, 256), 256, 0, st>>>(cw, h->conv_t[i], (int)chans[i][0], (int)chans[i][1]);
    }
    h->fconv_w = W("final_conv.weight", C * C); h->fconv_b = W("final_conv.bias", C);
  }
  if (!err.empty()) A2P_FAIL("bind_weights: %s", err.c_str());
  h->wplanes.clear(); h->wnumel.clear();
  if (cf.split_terms > 0) {
    __nv_bfloat16* cur = reinterpret_cast<__nv_bfloat16*>(pb + pl.planes);
    const __nv_bfloat16* endp = reinterpret_cast<__nv_bfloat16*>(pb + pl.planes + pl.planes_bytes);
    auto split_w = [&](const float* wsrc, long long rows, long long cols) -> int {
      const long long numel = rows * cols;
      if (cur + numel * cf.split_terms > endp) A2P_FAIL("bind_weights: plane arena overflow");
      A2P_TRY(launch_split_planes(cf.split_terms, wsrc, cols, cur, numel, rows, (int)cols, 1.f, st));
      h->wplanes[wsrc] = cur; h->wnumel[wsrc] = numel;
      cur += align_up((size_t)numel * cf.split_terms, 64);
      return 0;
    };
    for (int l = 0; l < cf.L; ++l) {
      const LayerW& lw = h->lw[l];
      A2P_TRY(split_w(lw.sa.in_w, 3 * D, D)); A2P_TRY(split_w(lw.sa.out_w, D, D));
      A2P_TRY(split_w(lw.ca.in_w, 3 * D, D)); A2P_TRY(split_w(lw.ca.out_w, D, D));
      if (cf.fmt == A2P_FMT_POSE) { A2P_TRY(split_w(lw.c2.in_w, 3 * D, D)); A2P_TRY(split_w(lw.c2.out_w, D, D)); }
      A2P_TRY(split_w(lw.l1w, FF, D)); A2P_TRY(split_w(lw.l2w, D, FF));
    }
    A2P_TRY(split_w(h->inp_w, D, C)); A2P_TRY(split_w(h->fin_w, C, D));
    if (cf.fmt == A2P_FMT_POSE) {
      const int64_t cm = C > 256 ? C : 256;
      const int64_t chans[6][2] = {{cm, C}, {C, cm}, {C, C}, {C, C}, {C, C}, {C, C}};
      for (int i = 0; i < 6; ++i) A2P_TRY(split_w(h->conv_t[i], 3 * chans[i][0], chans[i][1]));   // planes [P][3][Cout][Cin]
      A2P_TRY(split_w(h->fconv_w, C, C));
    }
    A2P_TRY(init_umma_gemm());
    A2P_TRY(init_umma_attn());
    A2P_TRY(init_umma_chain());
    A2P_TRY(init_umma_attn2());
    A2P_TRY(init_umma_attn_short());
  }
  {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  rope_table_kernel<<<ceil_div(cf.max_pos * (int)(D / 2), 256), 256, 0, st>>>(freqs, h->rope_tab, cf.max_pos, (int)(D / 2));
  A2P_CUDA(cudaGetLastError());
  h->bound = true;
  h->gvalid = false;
  return 0;
}

size_t a2p_kv_cache_bytes(const a2p_model_cfg* cfg, int Bc, int S) {
  if (check_cfg(cfg) || Bc <= 0 || S <= 0) return 0;
  return kv_layout(*cfg, Bc, S, cfg->S2).total * sizeof(float);
}

size_t a2p_conditioning_workspace_bytes(const a2p_model_cfg* cfg, int Bc, int S) {
  if (check_cfg(cfg) || Bc <= 0 || S <= 0) return 0;
  const size_t D = cfg->D;
  return (2 * align_up((size_t)Bc * S * D, 64) + align_up((size_t)Bc * (cfg->S2 > 0 ? cfg->S2 : 1) * D, 64)) * sizeof(float) + 512;
}

size_t a2p_workspace_bytes(const a2p_model_cfg* cfg, int B, int T) {
  if (check_cfg(cfg) || B <= 0 || T <= 0) return 0;
  const size_t one = ws_layout(*cfg, B, T).total;
  if (!(cfg->split_terms == 2 && cfg->D == 256)) return one;
  // the fused arm runs a CFG sampling step as 2 G concurrent forwards (sample_loop_impl), each in its own workspace region
  size_t need = 2 * align_up(one, 1024);
  for (int G = 2; G <= a2p_denoiser::MAXU / 2 && G <= B; G *= 2) {
    const size_t region = align_up(ws_layout(*cfg, ceil_div(B, G), T).total, 1024);
    const size_t n = 2 * (size_t)G * region + 256 + align_up((size_t)B * T * cfg->C * 4, 256);
    if (n > need) need = n;
  }
  return need;
}

// ---- step-invariant conditioning encoders (model/diffusion.py:355-381), exact fp32 (FFMA GEMMs, fp32 attention): they run once
// per distinct y and their result feeds every one of the 2000 denoiser evaluations of a loop
__global__ void pad_cols_kernel(const float* __restrict__ src, long long ld_src, int cols, float* __restrict__ dst, long long ld_dst,
                                long long rows) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * ld_dst) return;
  const long long r = idx / ld_dst;
  const int c = (int)(idx - r * ld_dst);
  dst[idx] = c < cols ? src[r * ld_src + c] : 0.f;
}
// out[b][d] = mean over s of tok[b][s][d]  (tokens.mean(dim=-2), model/diffusion.py:380); fp64 accumulation, one thread per column
__global__ void mean_rows_kernel(const float* __restrict__ tok, int S, int D, float* __restrict__ out) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (d >= D) return;
  const float* p = tok + (long long)b * S * D + d;
  double acc = 0.0;
  for (int s_ = 0; s_ < S; ++s_) acc += (double)p[(long long)s_ * D];
  out[(long long)b * D + d] = (float)(acc / (double)S);
}

static size_t encode_ws_floats(const a2p_model_cfg& cf, int Bc, int S, int feat_dim) {
  const size_t D = cf.D, rows = (size_t)Bc * S;
  const size_t Kp = align_up((size_t)feat_dim, 8);
  size_t n = 0;
  if ((size_t)feat_dim != Kp) n += align_up(rows * Kp, 64) + align_up(D * Kp, 64);
  n += 3 * align_up((size_t)Bc * D, 64);                                   // pooled mean, LayerNorm, hidden layer
  if (cf.fmt == A2P_FMT_POSE) n += align_up((size_t)Bc * (cf.S2 > 0 ? cf.S2 : 1) * D, 64);
  else n += 2 * align_up(rows * D, 64) + align_up(rows * 3 * D, 64) + align_up(rows * D, 64) + align_up(rows * cf.FF, 64);
  return n;
}

size_t a2p_encode_workspace_bytes(const a2p_model_cfg* cfg, int Bc, int S, int feat_dim) {
  if (check_cfg(cfg) || Bc <= 0 || S <= 0 || feat_dim <= 0) return 0;
  return encode_ws_floats(*cfg, Bc, S, feat_dim) * sizeof(float) + 512;
}

int a2p_denoiser_encode_conditioning(a2p_denoiser_t* h, int Bc, int S, int S2, int feat_dim, const float* feats,
                                     const float* keyframes, float* cond_tokens, float* cond_hidden, float* pose_tokens,
                                     void* ws, size_t ws_bytes, void* stream) {
  if (!h || !h->bound) A2P_FAIL("encode_conditioning: weights not bound");
  const a2p_model_cfg& cf = h->cfg;
  const a2p_denoiser::EncW& ew = h->encw;
  if (!ew.ok) A2P_FAIL("encode_conditioning: the bound weight table has no (complete) cond_projection / non_attn_cond_projection / %s",
                       cf.fmt == A2P_FMT_POSE ? "frame_cond_projection" : "cond_encoder");
  if (Bc <= 0 || S <= 0 || !feats || !cond_tokens || !cond_hidden || !ws) A2P_FAIL("encode_conditioning: bad argument");
  if (feat_dim != ew.feat_dim) A2P_FAIL("encode_conditioning: feature width %d, cond_projection expects %d", feat_dim, ew.feat_dim);
  if (cf.fmt == A2P_FMT_POSE && (!keyframes || !pose_tokens || S2 <= 0 || S2 > cf.S2))
    A2P_FAIL("encode_conditioning: pose model needs keyframes / pose_tokens with 0 < S2 <= %d", cf.S2);
  if (S > cf.max_pos) A2P_FAIL("encode_conditioning: S=%d exceeds cfg.max_pos=%d", S, cf.max_pos);
  if (ws_bytes < a2p_encode_workspace_bytes(&cf, Bc, S, feat_dim)) A2P_FAIL("encode_conditioning: workspace too small");
  const int D = cf.D, rows = Bc * S;
  Ctx c{h, (cudaStream_t)stream};
  cudaStream_t st = c.st;
  c.cat = CAT_COND;
  float* wp = reinterpret_cast<float*>(ws);
  auto take = [&](size_t n) { float* r = wp; wp += align_up(n, 64); return r; };
  // ---- tokens = cond_projection(feats)   (model/diffusion.py:372; K padded to a multiple of 8 for the GEMM's 16-byte loads)
  const int Kp = (int)align_up((size_t)feat_dim, 8);
  const float *A = feats, *Wc = ew.cp_w;
  if (Kp != feat_dim) {
    float* Ap = take((size_t)rows * Kp);
    float* Wp = take((size_t)D * Kp);
    pad_cols_kernel<<<(unsigned)(((long long)rows * Kp + 255) / 256), 256, 0, st>>>(feats, feat_dim, feat_dim, Ap, Kp, rows);
    pad_cols_kernel<<<(unsigned)(((long long)D * Kp + 255) / 256), 256, 0, st>>>(ew.cp_w, feat_dim, feat_dim, Wp, Kp, D);
    A2P_CUDA(cudaGetLastError());
    h->launches += 2;
    A = Ap; Wc = Wp;
  }
  A2P_TRY(gemm(c, A, Kp, rows, Wc, Kp, ew.cp_b, D, Kp, cond_tokens, D));
  float* pooled = take((size_t)Bc * D);
  float* pooled_n = take((size_t)Bc * D);
  float* hid1 = take((size_t)Bc * D);
  if (cf.fmt == A2P_FMT_FACE) {
    // ---- cond_encoder: two pre-LN encoder layers with rotary self-attention (transformer_modules.py:69-102, model/diffusion.py:158-171)
    float* hn = take((size_t)rows * D);
    float* hr = take((size_t)rows * D);
    float* qkv = take((size_t)rows * 3 * D);
    float* att = take((size_t)rows * D);
    float* u = take((size_t)rows * cf.FF);
    const long long sS = (long long)S * D;
    const float scale_log2e = (1.0f / sqrtf((float)h->dh)) * 1.4426950408889634f;
    for (int i = 0; i < 2; ++i) {
      const a2p_denoiser::EncLayerW& e = ew.enc[i];
      A2P_TRY(launch_ln_rope(D, cond_tokens, D, e.n1w, e.n1b, hn, hr, D, h->rope_tab, S, 0, rows, st));
      h->launches++;
      A2P_TRY(gemm(c, hr, D, rows, e.in_w, D, e.in_b, 2 * D, D, qkv, 3 * D));                                        // q | k from the rotated rows
      A2P_TRY(gemm(c, hn, D, rows, e.in_w + (size_t)2 * D * D, D, e.in_b + 2 * D, D, D, qkv + 2 * D, 3 * D));       // v from the plain rows
      AttnParams a{};
      a.Q = qkv; a.q_ld = 3 * D; a.q_sample_stride = 3 * sS;
      a.K.base[0] = qkv + D; a.K.stride[0] = 3 * sS; a.K.base[1] = nullptr; a.K.stride[1] = 0; a.K.rows_per_branch = Bc;
      a.V = a.K; a.V.base[0] = qkv + 2 * D;
      a.kv_ld = 3 * D; a.S_main = S; a.S_extra = 0;
      a.O = att; a.o_ld = D; a.o_sample_stride = sS; a.T = S; a.H = cf.H; a.R = Bc; a.scale_log2e = scale_log2e;
      A2P_TRY(launch_attn_simt(a, h->dh, st));
      h->launches++;
      A2P_TRY(gemm(c, att, D, rows, e.out_w, D, e.out_b, D, D, cond_tokens, D, EPI_RESID));
      A2P_TRY(launch_ln_rope(D, cond_tokens, D, e.n2w, e.n2b, hn, nullptr, D, h->rope_tab, S, 0, rows, st));
      h->launches++;
      A2P_TRY(gemm(c, hn, D, rows, e.l1w, D, e.l1b, cf.FF, D, u, cf.FF, EPI_GELU));
      A2P_TRY(gemm(c, u, cf.FF, rows, e.l2w, cf.FF, e.l2b, D, cf.FF, cond_tokens, D, EPI_RESID));
    }
  }
  // ---- hidden = non_attn_cond_projection(mean_S tokens)   (model/diffusion.py:380-381)
  mean_rows_kernel<<<dim3(ceil_div(D, 128), Bc), 128, 0, st>>>(cond_tokens, S, D, pooled);
  A2P_CUDA(cudaGetLastError());
  A2P_TRY(launch_ln_rope(D, pooled, D, ew.p0w, ew.p0b, pooled_n, nullptr, D, h->rope_tab, 1, 0, Bc, st));
  h->launches += 2;
  A2P_TRY(gemm(c, pooled_n, D, Bc, ew.p1w, D, ew.p1b, D, D, hid1, D, EPI_SILU));
  A2P_TRY(gemm(c, hid1, D, Bc, ew.p3w, D, ew.p3b, D, D, cond_hidden, D));
  if (cf.fmt == A2P_FMT_POSE) {
    // ---- pose tokens = frame_norm_cond(frame_cond_projection(keyframes))   (model/diffusion.py:316-336; unknown keyframes
    //      are zeroed by the caller; the null-embedding select of the uncond branch is the caller's too)
    float* ph = take((size_t)Bc * S2 * D);
    A2P_TRY(gemm(c, keyframes, cf.C, Bc * S2, ew.fp_w, cf.C, ew.fp_b, D, cf.C, ph, D));
    A2P_TRY(launch_ln_rope(D, ph, D, ew.fn_w, ew.fn_b, pose_tokens, nullptr, D, h->rope_tab, 1, 0, Bc * S2, st));
    h->launches++;
  }
  return 0;
}

int a2p_denoiser_set_conditioning(a2p_denoiser_t* h, int branch, int Bc, int S, int S2, const float* cond_tokens,
                                  const float* cond_hidden, const float* pose_tokens, void* kv_cache, size_t kv_bytes,
                                  void* ws, size_t ws_bytes, void* stream) {
  if (!h || !h->bound) A2P_FAIL("set_conditioning: weights not bound");
  if (branch != 0 && branch != 1) A2P_FAIL("set_conditioning: branch must be 0 or 1");
  const a2p_model_cfg& cf = h->cfg;
  if (Bc <= 0 || S <= 0 || !cond_tokens || !cond_hidden || !kv_cache || !ws) A2P_FAIL("set_conditioning: bad argument");
  if (cf.fmt == A2P_FMT_POSE && (!pose_tokens || S2 <= 0 || S2 > cf.S2)) A2P_FAIL("set_conditioning: pose model needs pose_tokens with 0 < S2 <= %d", cf.S2);
  if (cf.fmt == A2P_FMT_FACE) S2 = 0;
  if (S + 2 > cf.max_pos) A2P_FAIL("set_conditioning: S+2=%d exceeds cfg.max_pos=%d", S + 2, cf.max_pos);
  const KvLayout kl = kv_layout(cf, Bc, S, S2);
  if (kv_bytes < kl.total * sizeof(float)) A2P_FAIL("set_conditioning: kv cache too small");
  const size_t D = cf.D;
  const size_t need = a2p_conditioning_workspace_bytes(&cf, Bc, S);
  if (ws_bytes < need) A2P_FAIL("set_conditioning: workspace too small (%zu < %zu)", ws_bytes, need);
  Ctx c{h, (cudaStream_t)stream};
  float* base = (float*)kv_cache;
  float* mem_n = (float*)ws;
  float* mem_r = mem_n + align_up((size_t)Bc * S * D, 64);
  float* pose_buf = mem_r + align_up((size_t)Bc * S * D, 64);
  // norm_cond on the audio rows + rotation at positions 0..S-1 (model/diffusion.py:392-393; transformer_modules.py:253)
  { int _c = c.cat; c.cat = CAT_LN; c.begin(); A2P_TRY(launch_ln_rope(cf.D, cond_tokens, D, h->normc_w, h->normc_b, mem_n, mem_r, D, h->rope_tab, S, 0, Bc * S, c.st)); c.end(); c.cat = _c; }
  h->launches++;
  float* pose_r = pose_buf;
  for (int l = 0; l < cf.L; ++l) {
    const LayerW& lw = h->lw[l];
    float* lb = base + kl.per_layer * l;
    A2P_TRY(gemm(c, mem_r, D, Bc * S, lw.ca.in_w + D * D, D, lw.ca.in_b + D, cf.D, cf.D, lb + kl.ka, D));
    A2P_TRY(gemm(c, mem_n, D, Bc * S, lw.ca.in_w + 2 * D * D, D, lw.ca.in_b + 2 * D, cf.D, cf.D, lb + kl.va, D));
  }
  if (cf.fmt == A2P_FMT_POSE) {
    long long rows = (long long)Bc * S2;
    rope_only_kernel<<<(unsigned)((rows * (D / 2) + 255) / 256), 256, 0, c.st>>>(pose_tokens, pose_r, h->rope_tab, cf.D, S2, rows);
    h->launches++;
    for (int l = 0; l < cf.L; ++l) {
      const LayerW& lw = h->lw[l];
      float* lb = base + kl.per_layer * l;
      A2P_TRY(gemm(c, pose_r, D, (int)rows, lw.c2.in_w + D * D, D, lw.c2.in_b + D, cf.D, cf.D, lb + kl.k2, D));
      A2P_TRY(gemm(c, pose_tokens, D, (int)rows, lw.c2.in_w + 2 * D * D, D, lw.c2.in_b + 2 * D, cf.D, cf.D, lb + kl.v2, D));
    }
  }
  if (cf.split_terms > 0) {
    // tensor-core attention operands: split-bf16 K planes [P][Bc*S][D] and V^T planes [P][D][Bc*Sp] (pad columns zero)
    const int P = cf.split_terms;
    for (int l = 0; l < cf.L; ++l) {
      float* lb = base + kl.per_layer * l;
      __nv_bfloat16* kaP = reinterpret_cast<__nv_bfloat16*>(lb + kl.kaP);
      __nv_bfloat16* vtaP = reinterpret_cast<__nv_bfloat16*>(lb + kl.vtaP);
      A2P_TRY(launch_split_planes(P, lb + kl.ka, D, kaP, (long long)Bc * S * D, (long long)Bc * S, cf.D, 1.f, c.st));
      A2P_CUDA(cudaMemsetAsync(vtaP, 0, sizeof(__nv_bfloat16) * (size_t)P * D * Bc * kl.Sp, c.st));
      A2P_TRY(launch_transpose_split(P, lb + kl.va, D, vtaP, (long long)D * Bc * kl.Sp, (long long)Bc * kl.Sp, Bc * S, cf.D, S,
                                     (long long)kl.Sp, 1.f, c.st));
      h->launches += 3;
      if (cf.fmt == A2P_FMT_POSE) {
        __nv_bfloat16* k2P = reinterpret_cast<__nv_bfloat16*>(lb + kl.k2P);
        __nv_bfloat16* vt2P = reinterpret_cast<__nv_bfloat16*>(lb + kl.vt2P);
        A2P_TRY(launch_split_planes(P, lb + kl.k2, D, k2P, (long long)Bc * S2 * D, (long long)Bc * S2, cf.D, 1.f, c.st));
        A2P_CUDA(cudaMemsetAsync(vt2P, 0, sizeof(__nv_bfloat16) * (size_t)P * D * Bc * kl.S2p, c.st));
        A2P_TRY(launch_transpose_split(P, lb + kl.v2, D, vt2P, (long long)D * Bc * kl.S2p, (long long)Bc * kl.S2p, Bc * S2, cf.D, S2,
                                       (long long)kl.S2p, 1.f, c.st));
        h->launches += 3;
      }
    }
  }
  A2P_CUDA(cudaMemcpyAsync(base + kl.hidden, cond_hidden, sizeof(float) * Bc * D, cudaMemcpyDeviceToDevice, c.st));
  CondSet& cs = h->cond[branch];
  cs.set = true; cs.Bc = Bc; cs.S = S; cs.S2 = S2; cs.base = base; cs.hidden = base + kl.hidden;
  h->gvalid = false;
  return 0;
}

int a2p_denoiser_forward(a2p_denoiser_t* h, int B, int T, const float* x, int x_layout, const int64_t* timesteps,
                         int branch_mask, float* out_cond, float* out_uncond, void* ws, size_t ws_bytes, void* stream) {
  if (!h || !h->bound) A2P_FAIL("forward: weights not bound");
  if (B <= 0 || T <= 0 || !x || !timesteps || !ws) A2P_FAIL("forward: bad argument");
  if (branch_mask < 1 || branch_mask > 3) A2P_FAIL("forward: branch_mask must be 1, 2 or 3");
  if ((branch_mask & A2P_MASK_COND) && !out_cond) A2P_FAIL("forward: out_cond is null");
  if ((branch_mask & A2P_MASK_UNCOND) && !out_uncond) A2P_FAIL("forward: out_uncond is null");
  const a2p_model_cfg& cf = h->cfg;
  const WsLayout w = ws_layout(cf, B, T);
  if (ws_bytes < w.total) A2P_FAIL("forward: workspace too small (%zu < %zu)", ws_bytes, w.total);
  Ctx c{h, (cudaStream_t)stream};
  char* wsb = (char*)ws;
  const float* xin = x;
  if (x_layout == A2P_LAYOUT_BC1T) {
    float* xt = reinterpret_cast<float*>(wsb + w.xin);
    A2P_TRY(transpose_in(c, x, xt, B, cf.C, T));
    xin = xt;
  }
  const float *x0c = nullptr, *x0u = nullptr;
  long long sstride = 0;
  A2P_TRY(forward_core(c, B, T, xin, (const long long*)timesteps, nullptr, branch_mask, wsb, &x0c, &x0u, &sstride));
  auto emit = [&](const float* src, float* dst) -> int {
    long long total = (long long)B * T * (cf.C / 4);
    copy_rows_kernel<<<(unsigned)((total + 255) / 256), 256, 0, c.st>>>(src, cf.C, sstride, dst, cf.C, (long long)T * cf.C, T, cf.C / 4, B);
    h->launches++;
    A2P_CUDA(cudaGetLastError());
    return 0;
  };
  if (x0c) A2P_TRY(emit(x0c, out_cond));
  if (x0u) A2P_TRY(emit(x0u, out_uncond));
  return 0;
}

int a2p_sampler_step(int kind, int B, int C, int T, const float* x_t, const float* x0_cond, const float* x0_uncond,
                     const float* scale, const float* coeffs, const float* noise, int clip_denoised, float* x_prev,
                     float* pred_xstart, void* stream) {
  if (kind != A2P_SAMPLER_DDIM && kind != A2P_SAMPLER_ANCESTRAL) A2P_FAIL("sampler_step: unknown kind %d", kind);
  if (B <= 0 || C <= 0 || T <= 0 || !x_t || !x0_cond || !coeffs || !x_prev || !pred_xstart) A2P_FAIL("sampler_step: bad argument");
  if (x0_uncond && !scale) A2P_FAIL("sampler_step: guidance needs scale");
  Ctx c{nullptr, (cudaStream_t)stream};
  K3Params p{};
  p.x_t = x_t; p.x0c = x0_cond; p.x0u = x0_uncond; p.scale = scale; p.coeffs = coeffs; p.step_counter = nullptr;
  p.noise = noise; p.noise_step_stride = 0; p.n_steps = 1; p.x_prev = x_prev; p.pred = pred_xstart;
  p.B = B; p.C = C; p.T = T; p.kind = kind; p.clip = clip_denoised; p.x0_sample_stride = (long long)T * C;
  return launch_k3(c, p);
}

static int sample_loop_impl(a2p_denoiser_t* h, int kind, int B, int T, int n_steps, const float* coeffs, const int64_t* timesteps,
                            const float* scale, float* x, float* pred_xstart, const float* noise_tape, int rng,
                            unsigned long long seed, long long row0, int clip_denoised, int branch_mask, int use_graph, void* ws,
                            size_t ws_bytes, void* stream) {
  if (!h || !h->bound) A2P_FAIL("sample_loop: weights not bound");
  if (kind != A2P_SAMPLER_DDIM && kind != A2P_SAMPLER_ANCESTRAL) A2P_FAIL("sample_loop: unknown kind %d", kind);
  if (B <= 0 || T <= 0 || n_steps <= 0 || !coeffs || !timesteps || !x || !pred_xstart || !ws) A2P_FAIL("sample_loop: bad argument");
  if (branch_mask != A2P_MASK_BOTH && branch_mask != A2P_MASK_COND) A2P_FAIL("sample_loop: branch_mask must be BOTH or COND");
  if (branch_mask == A2P_MASK_BOTH && !scale) A2P_FAIL("sample_loop: CFG needs scale");
  if (kind == A2P_SAMPLER_ANCESTRAL && !noise_tape && !rng) A2P_FAIL("sample_loop: ancestral sampling needs a noise tape (or a2p_sample_loop_rng)");
  const a2p_model_cfg& cf = h->cfg;
  const WsLayout w = ws_layout(cf, B, T);
  if (ws_bytes < w.total) A2P_FAIL("sample_loop: workspace too small (%zu < %zu)", ws_bytes, w.total);
  Ctx c{h, (cudaStream_t)stream};
  char* wsb = (char*)ws;
  int* counter = reinterpret_cast<int*>(wsb + w.counter);
  float* xin = reinterpret_cast<float*>(wsb + w.xin);

  // ---- concurrent forwards.  No op mixes batch rows, so a CFG step is cut into units = {cond, uncond} x G groups of batch
  // rows; every unit is an independent forward with its own workspace region and stream (inside the capture: a parallel
  // branch of the graph), forked after the input transpose and joined before the sampler update.  Each kernel then waits
  // only for ITS predecessor, and launches that cannot fill the machine alone (38 chain CTAs, partial attention rounds)
  // run beside the other units' launches.
  const bool multi = branch_mask == A2P_MASK_BOTH && cf.split_terms == 2 && cf.D == 256 && (T % 8 == 0) && T >= 128 &&
                     !chain_disabled() && !branch_streams_disabled();
  int G = multi ? branch_groups(B) : 1;
  if (G > B) G = B;
  const int Bs_max = ceil_div(B, G);
  const size_t region = align_up(ws_layout(cf, Bs_max, T).total, 1024);
  const int units = 2 * G;
  const size_t shared_off = (size_t)units * region;        // G > 1: step counter + transposed input behind the unit regions
  const size_t need = G == 1 ? 2 * region : shared_off + 256 + align_up((size_t)B * T * cf.C * 4, 256);
  const bool use_units = multi && ws_bytes >= need;
  if (use_units && G > 1) {
    counter = reinterpret_cast<int*>(wsb + shared_off);
    xin = reinterpret_cast<float*>(wsb + shared_off + 256);
  }
  // Independent pipelines (default when the step is cut into G > 1 row groups; A2P_GROUP_PIPELINES=0 joins the groups every
  // step): nothing in a step couples the row groups -- the CFG mix and the sampler update are per row -- so every group runs its
  // own chain [transpose -> cond || uncond forward -> K3 -> its own step counter] on its own stream, one graph per group.  The
  // groups drift apart instead of draining the machine together at every step boundary (the tail of a forward is eight small TCN
  // launches), and their heavy phases interleave.
  static const bool pipes_env = !(getenv("A2P_GROUP_PIPELINES") && atoi(getenv("A2P_GROUP_PIPELINES")) == 0);
  const int n_pipes = (use_units && G > 1 && pipes_env) ? G : 1;
  // groups [g_lo, g_hi) of one step on stream c.st (g_lo = 0, g_hi = G: the whole step)
  auto step_body = [&](int g_lo, int g_hi) -> int {
    const int rb0 = use_units ? (int)((long long)B * g_lo / G) : 0, rb1 = use_units ? (int)((long long)B * g_hi / G) : B;
    int* cnt = counter + (n_pipes > 1 ? g_lo : 0);
    {   // [B,C,1,T] -> [B,T,C] for the rows of these groups
      dim3 grid(ceil_div(T, 32), ceil_div(cf.C, 32), rb1 - rb0), block(32, 8);
      bct_to_btc_kernel<<<grid, block, 0, c.st>>>(x + (size_t)rb0 * cf.C * T, xin + (size_t)rb0 * T * cf.C, cf.C, T);
      h->launches++;
      A2P_CUDA(cudaGetLastError());
    }
    const float *x0c[a2p_denoiser::MAXU] = {}, *x0u[a2p_denoiser::MAXU] = {};
    int gb0[a2p_denoiser::MAXU] = {}, gB[a2p_denoiser::MAXU] = {};
    long long sstride = 0;
    if (use_units) {
      cudaEvent_t& evf = h->ev_gfork[g_lo];
      if (!evf) A2P_CUDA(cudaEventCreateWithFlags(&evf, cudaEventDisableTiming));
      const int stag = branch_stagger();
      const int u_first = 2 * g_lo;
      for (int u = 2 * g_lo; u < 2 * g_hi; ++u) {
        const int g = u >> 1, br = u & 1;                   // the first unit of the range runs on the caller's stream
        const int b0 = (int)((long long)B * g / G), b1 = (int)((long long)B * (g + 1) / G);
        gb0[g] = b0; gB[g] = b1 - b0;
        Ctx cu{h, c.st};
        cu.slot = u;
        cu.concurrent = units;
        if (u == u_first) {
          if (stag == 0) A2P_CUDA(cudaEventRecord(evf, c.st));
          else { cu.stagger_ev = evf; cu.stagger_after = stag; }
        } else {
          if (!h->unit_stream[u]) A2P_CUDA(cudaStreamCreateWithFlags(&h->unit_stream[u], cudaStreamNonBlocking));
          if (!h->ev_ujoin[u]) A2P_CUDA(cudaEventCreateWithFlags(&h->ev_ujoin[u], cudaEventDisableTiming));
          A2P_CUDA(cudaStreamWaitEvent(h->unit_stream[u], evf, 0));
          cu.st = h->unit_stream[u];
        }
        const float* dummy = nullptr;
        A2P_TRY(forward_core(cu, b1 - b0, T, xin + (size_t)b0 * T * cf.C, (const long long*)timesteps, cnt,
                             br ? A2P_MASK_UNCOND : A2P_MASK_COND, wsb + (size_t)u * region, br ? &dummy : &x0c[g],
                             br ? &x0u[g] : &dummy, &sstride, b0, B));
        if (u == u_first && cu.stagger_ev) A2P_CUDA(cudaEventRecord(evf, c.st));   // fewer launches than the stagger asked for
        if (u != u_first) A2P_CUDA(cudaEventRecord(h->ev_ujoin[u], h->unit_stream[u]));
      }
      for (int u = u_first + 1; u < 2 * g_hi; ++u) A2P_CUDA(cudaStreamWaitEvent(c.st, h->ev_ujoin[u], 0));
    } else {
      gB[0] = B;
      A2P_TRY(forward_core(c, B, T, xin, (const long long*)timesteps, cnt, branch_mask, wsb, &x0c[0], &x0u[0], &sstride));
    }
    for (int g = g_lo; g < (use_units ? g_hi : 1); ++g) {
      const size_t off = (size_t)gb0[g] * cf.C * T;
      K3Params p{};
      p.x_t = x + off; p.x0c = x0c[g]; p.x0u = x0u[g]; p.scale = scale ? scale + gb0[g] : nullptr; p.coeffs = coeffs; p.step_counter = cnt;
      p.noise = noise_tape ? noise_tape + off : nullptr; p.noise_step_stride = (long long)B * cf.C * T; p.n_steps = n_steps;
      p.x_prev = x + off; p.pred = pred_xstart + off; p.B = gB[g]; p.C = cf.C; p.T = T; p.kind = kind; p.clip = clip_denoised;
      p.x0_sample_stride = sstride;
      p.rng = rng; p.seed = seed; p.rng_row0 = row0 + gb0[g];
      A2P_TRY(launch_k3(c, p));
    }
    step_dec_kernel<<<1, 1, 0, c.st>>>(cnt);
    h->launches++;
    A2P_CUDA(cudaGetLastError());
    return 0;
  };

  for (int pi = 0; pi < n_pipes; ++pi) {
    step_set_kernel<<<1, 1, 0, c.st>>>(counter + pi, n_steps - 1);
    h->launches++;
  }
  // pipelines 1.. start after everything the caller queued so far (conditioning, the counters) and hand back at the end
  cudaStream_t user = c.st;
  auto pipe_stream = [&](int pi) -> cudaStream_t { return pi == 0 ? user : h->group_stream[pi]; };
  if (n_pipes > 1) {
    if (!h->ev_gstart) A2P_CUDA(cudaEventCreateWithFlags(&h->ev_gstart, cudaEventDisableTiming));
    A2P_CUDA(cudaEventRecord(h->ev_gstart, user));
    for (int pi = 1; pi < n_pipes; ++pi) {
      if (!h->group_stream[pi]) A2P_CUDA(cudaStreamCreateWithFlags(&h->group_stream[pi], cudaStreamNonBlocking));
      if (!h->ev_gdone[pi]) A2P_CUDA(cudaEventCreateWithFlags(&h->ev_gdone[pi], cudaEventDisableTiming));
      A2P_CUDA(cudaStreamWaitEvent(h->group_stream[pi], h->ev_gstart, 0));
    }
  }
  auto join_pipes = [&]() -> int {
    for (int pi = 1; pi < n_pipes; ++pi) {
      A2P_CUDA(cudaEventRecord(h->ev_gdone[pi], h->group_stream[pi]));
      A2P_CUDA(cudaStreamWaitEvent(user, h->ev_gdone[pi], 0));
    }
    return 0;
  };
  auto pipe_groups = [&](int pi, int* lo, int* hi) { if (n_pipes > 1) { *lo = pi; *hi = pi + 1; } else { *lo = 0; *hi = G; } };
  if (!use_graph) {
    for (int i = 0; i < n_steps; ++i)
      for (int pi = 0; pi < n_pipes; ++pi) {
        int lo, hi; pipe_groups(pi, &lo, &hi);
        c.st = pipe_stream(pi);
        A2P_TRY(step_body(lo, hi));
      }
    c.st = user;
    return join_pipes();
  }
  GraphKey key{B, T, kind, n_steps, clip_denoised, branch_mask, coeffs, timesteps, scale, x, pred_xstart, noise_tape, ws,
               h->cond[0].base, h->cond[1].base, seed, row0, rng, use_units ? units : 0, n_pipes};
  if (!h->gvalid || !(h->gkey == key)) {
    for (int i = 0; i < a2p_denoiser::MAXG; ++i) if (h->gexec[i]) { cudaGraphExecDestroy(h->gexec[i]); h->gexec[i] = nullptr; }
    h->gvalid = false;
    h->graph_nodes = 0;
    if (!h->cap_stream) A2P_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
    const int64_t before = h->launches;
    for (int pi = 0; pi < n_pipes; ++pi) {
      cudaGraph_t graph = nullptr;
      int lo, hi; pipe_groups(pi, &lo, &hi);
      c.st = h->cap_stream;
      A2P_CUDA(cudaStreamBeginCapture(c.st, cudaStreamCaptureModeRelaxed));
      int rc = step_body(lo, hi);
      cudaError_t ce = cudaStreamEndCapture(c.st, &graph);
      c.st = user;
      if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
      if (ce != cudaSuccess) A2P_FAIL("graph capture failed: %s", cudaGetErrorString(ce));
      size_t nn = 0;
      cudaGraphGetNodes(graph, nullptr, &nn);
      {   // count the KERNEL nodes only (memset / empty join nodes are not launches of ours)
        std::vector<cudaGraphNode_t> nodes(nn);
        if (nn) cudaGraphGetNodes(graph, nodes.data(), &nn);
        for (size_t i = 0; i < nn; ++i) {
          cudaGraphNodeType ty;
          if (cudaGraphNodeGetType(nodes[i], &ty) == cudaSuccess && ty == cudaGraphNodeTypeKernel) h->graph_nodes++;
        }
      }
      ce = cudaGraphInstantiate(&h->gexec[pi], graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) A2P_FAIL("graph instantiate failed: %s", cudaGetErrorString(ce));
    }
    h->launches = before;  // captured launches are counted per replay below
    h->n_gexec = n_pipes;
    h->gkey = key;
    h->gvalid = true;
  }
  for (int i = 0; i < n_steps; ++i)
    for (int pi = 0; pi < n_pipes; ++pi) A2P_CUDA(cudaGraphLaunch(h->gexec[pi], pipe_stream(pi)));
  h->launches += h->graph_nodes * n_steps;
  return join_pipes();
}

int a2p_sample_loop(a2p_denoiser_t* h, int kind, int B, int T, int n_steps, const float* coeffs, const int64_t* timesteps,
                    const float* scale, float* x, float* pred_xstart, const float* noise_tape, int clip_denoised,
                    int branch_mask, int use_graph, void* ws, size_t ws_bytes, void* stream) {
  return sample_loop_impl(h, kind, B, T, n_steps, coeffs, timesteps, scale, x, pred_xstart, noise_tape, 0, 0ull, 0, clip_denoised,
                          branch_mask, use_graph, ws, ws_bytes, stream);
}

int a2p_sample_loop_rng(a2p_denoiser_t* h, int kind, int B, int T, int n_steps, const float* coeffs, const int64_t* timesteps,
                        const float* scale, float* x, float* pred_xstart, uint64_t seed, int64_t row0, int clip_denoised,
                        int branch_mask, int use_graph, void* ws, size_t ws_bytes, void* stream) {
  return sample_loop_impl(h, kind, B, T, n_steps, coeffs, timesteps, scale, x, pred_xstart, nullptr, 1, (unsigned long long)seed,
                          (long long)row0, clip_denoised, branch_mask, use_graph, ws, ws_bytes, stream);
}

int a2p_sampler_step_rng(int kind, int B, int C, int T, const float* x_t, const float* x0_cond, const float* x0_uncond,
                         const float* scale, const float* coeffs, uint64_t seed, int64_t iteration, int64_t row0,
                         int clip_denoised, float* x_prev, float* pred_xstart, void* stream) {
  if (kind != A2P_SAMPLER_DDIM && kind != A2P_SAMPLER_ANCESTRAL) A2P_FAIL("sampler_step_rng: unknown kind %d", kind);
  if (B <= 0 || C <= 0 || T <= 0 || !x_t || !x0_cond || !coeffs || !x_prev || !pred_xstart) A2P_FAIL("sampler_step_rng: bad argument");
  if (x0_uncond && !scale) A2P_FAIL("sampler_step_rng: guidance needs scale");
  Ctx c{nullptr, (cudaStream_t)stream};
  K3Params p{};
  p.x_t = x_t; p.x0c = x0_cond; p.x0u = x0_uncond; p.scale = scale; p.coeffs = coeffs; p.step_counter = nullptr;
  p.noise = nullptr; p.noise_step_stride = (long long)iteration; p.n_steps = 1; p.x_prev = x_prev; p.pred = pred_xstart;
  p.B = B; p.C = C; p.T = T; p.kind = kind; p.clip = clip_denoised; p.x0_sample_stride = (long long)T * C;
  p.rng = 1; p.seed = (unsigned long long)seed; p.rng_row0 = (long long)row0;
  return launch_k3(c, p);
}

int a2p_profile_forward_rows(a2p_denoiser_t* h, int B_total, int b0, int Bs, int T, const float* x_btc, const int64_t* timesteps,
                             int branch_mask, void* ws, size_t ws_bytes, void* stream, float* ms_by_cat,
                             int64_t* launches_by_cat, int ncat) {
  if (!h || !h->bound) A2P_FAIL("profile_forward: weights not bound");
  if (ncat < CAT_N || !ms_by_cat || !launches_by_cat) A2P_FAIL("profile_forward: need %d categories", (int)CAT_N);
  if (b0 < 0 || Bs <= 0 || b0 + Bs > B_total) A2P_FAIL("profile_forward: bad row range");
  const WsLayout w = ws_layout(h->cfg, Bs, T);
  if (ws_bytes < w.total) A2P_FAIL("profile_forward: workspace too small");
  Prof prof;
  Ctx c{h, (cudaStream_t)stream};
  c.prof = &prof;
  if (branch_mask != A2P_MASK_BOTH) {   // one unit of a sampling step: launch shapes (chain N split) as in the loop
    const int G = a2p_loop_row_groups(h, B_total, T);
    c.concurrent = G > 0 ? 2 * G : 1;
  }
  const float *x0c, *x0u; long long ss;
  A2P_TRY(forward_core(c, Bs, T, x_btc, (const long long*)timesteps, nullptr, branch_mask, (char*)ws, &x0c, &x0u, &ss, b0, B_total));
  A2P_CUDA(cudaStreamSynchronize(c.st));
  for (int i = 0; i < ncat; ++i) { ms_by_cat[i] = 0.f; launches_by_cat[i] = 0; }
  const bool dump = getenv("A2P_PROFILE_DUMP") != nullptr;
  for (size_t i = 0; i < prof.cat.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, prof.ev[2 * i], prof.ev[2 * i + 1]);
    ms_by_cat[prof.cat[i]] += ms;
    launches_by_cat[prof.cat[i]]++;
    if (dump) fprintf(stderr, "a2p_prof %3zu cat=%d %8.1f us  %s\n", i, prof.cat[i], ms * 1e3f, prof.info[i].c_str());
  }
  for (auto e : prof.ev) cudaEventDestroy(e);
  return 0;
}

int a2p_profile_forward(a2p_denoiser_t* h, int B, int T, const float* x_btc, const int64_t* timesteps, int branch_mask,
                        void* ws, size_t ws_bytes, void* stream, float* ms_by_cat, int64_t* launches_by_cat, int ncat) {
  return a2p_profile_forward_rows(h, B, 0, B, T, x_btc, timesteps, branch_mask, ws, ws_bytes, stream, ms_by_cat, launches_by_cat, ncat);
}

int a2p_loop_row_groups(const a2p_denoiser_t* h, int B, int T) {
  if (!h) return 1;
  const a2p_model_cfg& cf = h->cfg;
  const bool multi = cf.split_terms == 2 && cf.D == 256 && (T % 8 == 0) && T >= 128 && !chain_disabled() && !branch_streams_disabled();
  if (!multi) return 0;                       // 0: one stacked forward for both branches
  const int G = branch_groups(B);
  return G > B ? B : G;
}

int64_t a2p_launch_count(const a2p_denoiser_t* h) { return h ? h->launches : 0; }

}  // extern "C"
