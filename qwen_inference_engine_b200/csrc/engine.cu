// engine.cu -- forward launch order + engine lifecycle. See engine.h.
#include "engine.h"

#include "launch.h"

#include <math.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace qie {

bool g_use_pdl = [] {
  const char* v = getenv("QIE_PDL");
  return !(v && v[0] == '0');
}();

// launch + count + (optionally) bracket with events for the per-kernel-class timing
#define QIE_TRY(kind, expr)                                    \
  do {                                                         \
    qie_engine::ProfRec _r{kind, nullptr, nullptr};            \
    if (e->prof_on) {                                          \
      cudaEventCreate(&_r.a);                                  \
      cudaEventCreate(&_r.b);                                  \
      cudaEventRecord(_r.a, e->stream);                        \
    }                                                          \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) {                                   \
      if (_r.a) cudaEventDestroy(_r.a);                        \
      if (_r.b) cudaEventDestroy(_r.b);                        \
      return _e;                                               \
    }                                                          \
    if (e->prof_on) {                                          \
      cudaEventRecord(_r.b, e->stream);                        \
      e->prof.push_back(_r);                                   \
    }                                                          \
    ++e->launches;                                             \
  } while (0)

static void capture_copy(qie_engine* e, const char* tag, int layer, const bf16* src, size_t elems) {
  if (!e->capture) return;
  std::string key = std::string(tag) + "#" + std::to_string(layer);
  CaptureBuf& b = e->cap[key];
  if (b.cap < elems) {
    if (b.d) cudaFree(b.d);
    cudaMalloc(&b.d, elems * sizeof(bf16));
    b.cap = elems;
  }
  b.elems = elems;
  cudaMemcpyAsync(b.d, src, elems * sizeof(bf16), cudaMemcpyDeviceToDevice, e->stream);
}

// C[rows, *] = A * W^T in chunks of <= 64 rows (the reference-order kernel keeps one
// fp32 fragment per 16 rows; larger M just re-streams W per chunk).
static cudaError_t gemm_rows(qie_engine* e, int kind, GemmArgs g, int n_rows, size_t a_stride) {
  for (int r0 = 0; r0 < n_rows; r0 += 64) {
    GemmArgs c = g;
    c.A = g.A + (size_t)r0 * a_stride;
    c.M = std::min(64, n_rows - r0);
    for (int s = 0; s < g.nseg; ++s)  // EPI_STORE_F32: `out` is a float*, twice the bf16 stride
      c.seg[s].out = g.seg[s].out + (size_t)r0 * g.seg[s].ld_out * (g.epi == EPI_STORE_F32 ? 2 : 1);
    QIE_TRY(kind, launch_gemm_ref_order(c, e->num_sms, e->stream));
  }
  return cudaSuccess;
}

// tcgen05 GEMM launch (fast numerics, rows > 8). x: [n, K] activations.
static cudaError_t gemm_tc(qie_engine* e, int kind, const bf16* x, int n, int K, int nseg, const TensorMap2D* const* w,
                           const int* rows, int epi, bf16* out, int ld_out) {
  TensorMap2D xmap;
  cudaError_t r = make_tensor_map_2d(&xmap, x, n, K, tc_token_tile(n));
  if (r != cudaSuccess) return r;
  TcGemm t{};
  int cols = 0;
  for (int i = 0; i < nseg; ++i) {
    t.w[i] = w[i];
    t.rows[i] = rows[i];
    cols += rows[i];
  }
  t.x = &xmap;
  t.nseg = nseg;
  t.M = n;
  t.K = K;
  t.epi = epi;
  t.out = out;
  t.ld_out = ld_out;
  t.ws = e->gemm_ws;
  t.ws_bytes = e->gemm_ws_bytes;
  t.max_splits = (int)std::max<size_t>(1, e->gemm_ws_bytes / ((size_t)n * cols * sizeof(float)));
  t.counters = e->gemm_counters;
  t.n_counters = 8192;
  t.w_static = 1;
  int nl = 0;
  qie_engine::ProfRec pr{kind, nullptr, nullptr};
  if (e->prof_on) {
    cudaEventCreate(&pr.a);
    cudaEventCreate(&pr.b);
    cudaEventRecord(pr.a, e->stream);
  }
  r = launch_gemm_tcgen05(t, e->num_sms, e->stream, &nl);
  if (r != cudaSuccess) return r;
  if (e->prof_on) {
    cudaEventRecord(pr.b, e->stream);
    e->prof.push_back(pr);
  }
  e->launches += nl;
  return cudaSuccess;
}

// Launch order = llm()'s (qwen_main.cu:77-222 prefill / :271-365 decode): per layer
// rms -> q,k,v -> q/k-norm -> RoPE -> KV store -> attention -> o (+residual) -> rms ->
// up,gate -> SiLU*up -> down (+residual); then final norm -> lm_head -> sampling.
// 18 kernels + 2-4 memcpys per layer in the reference become 8 launches here.
cudaError_t forward_rows(qie_engine* e, int n, int max_kv_len, int out_row0, int n_out, float temperature,
                         bool advance) {
  if (e->tp.size > 1) return forward_rows_tp(e, n, max_kv_len, out_row0, n_out, temperature, advance);
  const qie_config& c = e->cfg;
  const int H = c.hidden, hd = c.head_dim, Dq = c.n_q * hd, Dkv = c.n_kv * hd, I = c.inter;
  cudaStream_t st = e->stream;
  const bool fast = e->opts.numerics == QIE_NUMERICS_FAST;
  const bool tc = fast && n > 8;                  // tcgen05 GEMMs
  const bool fast_attn = fast && advance;         // decode rows: one token per sequence
  const bool fast_prefill = fast && !advance;     // prefill rows: consecutive positions of one sequence
  auto rms = [&](const bf16* x, const bf16* w, bf16* y, int rows) {
    return fast ? launch_rmsnorm_fast(x, w, y, H, rows, H, st, e->eps) : launch_rmsnorm_ref(x, w, y, H, rows, H, st, e->eps);
  };

  if (e->hist_d) QIE_TRY(KK_EMBED, launch_history_append(e->hist_d, (size_t)c.context, e->ids_d, e->pos_d, e->slot_d, n, st));
  if (!e->inject_x) QIE_TRY(KK_EMBED, launch_embedding(e->x, e->embed, e->ids_d, H, n, st));
  const int l_begin = e->layer_count > 0 ? std::min(e->layer_first, c.layers) : 0;
  const int l_end = e->layer_count > 0 ? std::min(l_begin + e->layer_count, c.layers) : c.layers;
  for (int l = l_begin; l < l_end; ++l) {
    const LayerWeights& w = e->L[l];
    QIE_TRY(KK_RMSNORM, rms(e->x, w.in_ln, e->xn, n));
    capture_copy(e, "input_norm", l, e->xn, (size_t)n * H);
    if (tc) {
      const TensorMap2D* wm[3] = {&e->wmaps[l].q, &e->wmaps[l].k, &e->wmaps[l].v};
      const int rows[3] = {Dq, Dkv, Dkv};
      cudaError_t r = gemm_tc(e, KK_GEMM_QKV, e->xn, n, H, 3, wm, rows, EPI_STORE, e->qkv, Dq + 2 * Dkv);
      if (r != cudaSuccess) return r;
    } else {
      GemmArgs g{};
      g.A = e->xn;
      g.lda = H;
      g.K = H;
      g.nseg = 3;
      g.epi = EPI_STORE;
      g.seg[0] = GemmSeg{w.q, nullptr, e->q, Dq, Dq};
      g.seg[1] = GemmSeg{w.k, nullptr, e->k, Dkv, Dkv};
      g.seg[2] = GemmSeg{w.v, nullptr, e->v, Dkv, Dkv};
      cudaError_t r = gemm_rows(e, KK_GEMM_QKV, g, n, H);
      if (r != cudaSuccess) return r;
    }
    {
      QkvPostArgs a{};
      a.q = e->q;
      if (tc) {
        a.q_in = e->qkv;
        a.k = e->qkv + Dq;
        a.v = e->qkv + Dq + Dkv;
        a.q_in_stride = a.kv_stride = Dq + 2 * Dkv;
      } else {
        a.q_in = e->q;
        a.k = e->k;
        a.v = e->v;
        a.q_in_stride = Dq;
        a.kv_stride = Dkv;
      }
      a.q_norm_w = w.q_norm;
      a.k_norm_w = w.k_norm;
      a.cos_t = e->cos_d;
      a.sin_t = e->sin_d;
      a.pos = e->pos_d;
      a.slot = e->slot_d;
      a.block_table = e->block_table_d;
      a.max_pages = e->max_pages_per_seq;
      a.n_tok = n;
      a.n_q = c.n_q;
      a.layer = l;
      a.kv = e->kv;
      a.q_bias = w.q_bias;
      a.k_bias = w.k_bias;
      a.v_bias = w.v_bias;
      a.eps = e->eps;
      a.rope_half = e->rope_half ? 1 : 0;
      QIE_TRY(KK_QKV_POST, launch_qkv_post(a, st));
    }
    capture_copy(e, "q", l, e->q, (size_t)n * Dq);
    if (!tc) capture_copy(e, "v", l, e->v, (size_t)n * Dkv);
    if (fast_attn) {
      FastAttnArgs a{};
      a.q = e->q;
      a.out = e->att;
      a.pos = e->pos_d;
      a.slot = e->slot_d;
      a.block_table = e->block_table_d;
      a.max_pages = e->max_pages_per_seq;
      a.n_tok = n;
      a.n_q = c.n_q;
      a.layer = l;
      const int tiles = (max_kv_len + 63) / 64;
      int ns = (4 * e->num_sms) / std::max(1, n * c.n_kv);
      static const int ns_env = [] { const char* v = getenv("QIE_ATTN_SPLITS"); return v ? atoi(v) : 0; }();  // tuning knob
      if (ns_env > 0) ns = ns_env;
      ns = std::max(1, std::min(ns, std::min(e->attn_max_splits, tiles)));
      a.n_splits = ns;
      a.scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
      a.ws_o = e->attn_ws_o;
      a.ws_ml = e->attn_ws_ml;
      a.kv = e->kv;
      QIE_TRY(KK_ATTN, launch_attention_decode_fast(a, st));
      if (ns > 1) ++e->launches;  // + combine kernel
    } else if (fast_prefill) {
      FastAttnArgs a{};
      a.q = e->q;
      a.out = e->att;
      a.pos = e->pos_d;
      a.slot = e->slot_d;
      a.block_table = e->block_table_d;
      a.max_pages = e->max_pages_per_seq;
      a.n_tok = n;
      a.n_q = c.n_q;
      a.layer = l;
      a.n_splits = 1;
      a.scale_log2 = 1.4426950408889634f / sqrtf((float)hd);
      a.kv = e->kv;
      // head_dim 128 and at least one full 128-row query tile: tcgen05 / TMEM kernel (QIE_ATTN_TC=0 keeps mma.sync)
      static const bool tc_on = [] { const char* v = getenv("QIE_ATTN_TC"); return !(v && v[0] == '0'); }();
      if (tc_on && hd == 128 && n >= 128)
        QIE_TRY(KK_ATTN, launch_attention_prefill_tc(a, 0, st));
      else
        QIE_TRY(KK_ATTN, launch_attention_prefill_fast(a, st));
    } else {
      AttnArgs a{};
      a.q = e->q;
      a.out = e->att;
      a.pos = e->pos_d;
      a.slot = e->slot_d;
      a.block_table = e->block_table_d;
      a.max_pages = e->max_pages_per_seq;
      a.n_tok = n;
      a.n_q = c.n_q;
      a.layer = l;
      a.max_kv_len = max_kv_len;
      a.kv = e->kv;
      QIE_TRY(KK_ATTN, launch_attention_ref(a, st));
    }
    capture_copy(e, "attn", l, e->att, (size_t)n * Dq);
    if (tc) {
      const TensorMap2D* wm[1] = {&e->wmaps[l].o};
      const int rows[1] = {H};
      cudaError_t r = gemm_tc(e, KK_GEMM_O, e->att, n, Dq, 1, wm, rows, EPI_RESIDUAL, e->x, H);
      if (r != cudaSuccess) return r;
    } else {
      GemmArgs g{};
      g.A = e->att;
      g.lda = Dq;
      g.K = Dq;
      g.nseg = 1;
      g.epi = EPI_RESIDUAL;
      g.seg[0] = GemmSeg{w.o, nullptr, e->x, H, H};
      cudaError_t r = gemm_rows(e, KK_GEMM_O, g, n, Dq);
      if (r != cudaSuccess) return r;
    }
    capture_copy(e, "x_attn", l, e->x, (size_t)n * H);
    QIE_TRY(KK_RMSNORM, rms(e->x, w.post_ln, e->xn, n));
    if (tc) {
      const TensorMap2D* wm[2] = {&e->wmaps[l].gate, &e->wmaps[l].up};
      const int rows[2] = {I, I};
      cudaError_t r = gemm_tc(e, KK_GEMM_GATEUP, e->xn, n, H, 2, wm, rows, EPI_SILU_MUL, e->h, I);
      if (r != cudaSuccess) return r;
    } else {
      GemmArgs g{};
      g.A = e->xn;
      g.lda = H;
      g.K = H;
      g.nseg = 1;
      g.epi = EPI_SILU_MUL;
      g.seg[0] = GemmSeg{w.gate, w.up, e->h, I, I};
      cudaError_t r = gemm_rows(e, KK_GEMM_GATEUP, g, n, H);
      if (r != cudaSuccess) return r;
    }
    capture_copy(e, "mlp_h", l, e->h, (size_t)n * I);
    if (tc) {
      const TensorMap2D* wm[1] = {&e->wmaps[l].down};
      const int rows[1] = {H};
      cudaError_t r = gemm_tc(e, KK_GEMM_DOWN, e->h, n, I, 1, wm, rows, EPI_RESIDUAL, e->x, H);
      if (r != cudaSuccess) return r;
    } else {
      GemmArgs g{};
      g.A = e->h;
      g.lda = I;
      g.K = I;
      g.nseg = 1;
      g.epi = EPI_RESIDUAL;
      g.seg[0] = GemmSeg{w.down, nullptr, e->x, H, H};
      cudaError_t r = gemm_rows(e, KK_GEMM_DOWN, g, n, I);
      if (r != cudaSuccess) return r;
    }
    capture_copy(e, "x_out", l, e->x, (size_t)n * H);
  }
  if (n_out == 0 || e->layer_count > 0) return cudaSuccess;  // non-final prefill chunk / layer-isolation hook: no logits
  // final norm only on the rows that feed lm_head (qwen_main.cu:227-236, :367-372)
  QIE_TRY(KK_RMSNORM, rms(e->x + (size_t)out_row0 * H, e->final_norm, e->xn, n_out));
  if (fast && n_out > 8) {
    const TensorMap2D* wm[1] = {&e->lm_head_map};
    const int rows[1] = {c.vocab};
    cudaError_t r = gemm_tc(e, KK_LM_HEAD, e->xn, n_out, H, 1, wm, rows, EPI_STORE, e->logits, c.vocab);
    if (r != cudaSuccess) return r;
  } else {
    GemmArgs g{};
    g.A = e->xn;
    g.lda = H;
    g.K = H;
    g.nseg = 1;
    g.epi = EPI_STORE;
    g.seg[0] = GemmSeg{e->lm_head, nullptr, e->logits, c.vocab, c.vocab};
    cudaError_t r = gemm_rows(e, KK_LM_HEAD, g, n_out, H);
    if (r != cudaSuccess) return r;
  }
  capture_copy(e, "logits", -1, e->logits, (size_t)n_out * c.vocab);
  if (e->hist_d && e->rep_penalty != 1.0f)  // the context of output row r: positions 0 .. pos[out_row0 + r] of its sequence
    QIE_TRY(KK_SAMPLE, launch_repetition_penalty(e->logits, e->hist_d, e->slot_d + out_row0, e->pos_d + out_row0, 1, 0, (size_t)max_kv_len,
                                                 (size_t)c.context, n_out, c.vocab, e->rep_penalty, st));
  QIE_TRY(KK_SAMPLE, launch_sample_topk(e->logits, e->sampled_d, n_out, c.vocab, temperature, e->topk, e->seed, 0,
                             e->add_step ? e->rowstep_d : nullptr, st));
  if (advance) QIE_TRY(KK_ADVANCE, launch_advance(e->pos_d, e->ids_d, e->sampled_d, n_out, e->rowstep_d, st));
  return cudaSuccess;
}


// weight-tile set: larger tiles for small batches when the ring still holds enough of them
static int mega_tile_set(const qie_engine* e, int n, int max_kv_len) {
  const qie_config& c = e->cfg;
  const bool tp = e->tp.size > 1;
  const int fast = !tp && e->opts.numerics == QIE_NUMERICS_FAST;
  const int I = tp ? e->plan.inter : c.inter, nq = tp ? e->plan.n_q : c.n_q, nkv = tp ? e->plan.n_kv : c.n_kv;
  if (n <= 8 && e->mega_kc[1] > e->mega_kc[0] &&
      decode_mega_supports(c.hidden, I, c.layers, nq, nkv, c.head_dim, n, max_kv_len, e->num_sms, e->mega_kc[1], fast, e->kv.page_size))
    return 1;
  return 0;
}

// Tensor-parallel forward (BASELINE configs[4]: 7B-arch over 2/4 GPUs).  Every rank holds the full
// weight blob and works on its shard by pointer arithmetic: q/k/v/gate/up/lm_head by output rows,
// o_proj/down_proj by input columns (row stride = the full width).  The two row-parallel GEMMs
// produce fp32 partial sums (the HMMA accumulators, not yet rounded) that are summed over ranks with
// ncclAllReduce (NVLink) and rounded to bf16 ONCE, like the unsharded projection output, before the
// residual add; the greedy token is the best of the ranks' local arg-max candidates in the
// reference's tie-break order.  Norms, residual stream and sampling state are replicated.
cudaError_t forward_rows_tp(qie_engine* e, int n, int max_kv_len, int out_row0, int n_out, float temperature,
                            bool advance) {
  const qie_config& c = e->cfg;
  const TpPlan& pl = e->plan;
  const int H = c.hidden, hd = c.head_dim, Dq = c.n_q * hd, I = c.inter;
  const int Dq_l = pl.n_q * hd, Dkv_l = pl.n_kv * hd, I_l = pl.inter, V_l = pl.vocab;
  cudaStream_t st = e->stream;
  if (!e->tp.comm) return cudaErrorNotReady;
  if (e->topk != 1) return cudaErrorNotSupported;  // top-k sampling over a sharded vocabulary is not built
  (void)temperature;

  QIE_TRY(KK_EMBED, launch_embedding(e->x, e->embed, e->ids_d, H, n, st));
  for (int l = 0; l < c.layers; ++l) {
    const LayerWeights& w = e->L[l];
    QIE_TRY(KK_RMSNORM, launch_rmsnorm_ref(e->x, w.in_ln, e->xn, H, n, H, st));
    {
      GemmArgs g{};
      g.A = e->xn;
      g.lda = H;
      g.K = H;
      g.nseg = 3;
      g.epi = EPI_STORE;
      g.seg[0] = GemmSeg{w.q + (size_t)pl.q_row0 * H, nullptr, e->q, Dq_l, Dq_l, 0};
      g.seg[1] = GemmSeg{w.k + (size_t)pl.kv_row0 * H, nullptr, e->k, Dkv_l, Dkv_l, 0};
      g.seg[2] = GemmSeg{w.v + (size_t)pl.kv_row0 * H, nullptr, e->v, Dkv_l, Dkv_l, 0};
      cudaError_t r = gemm_rows(e, KK_GEMM_QKV, g, n, H);
      if (r != cudaSuccess) return r;
    }
    {
      QkvPostArgs a{};
      a.q = e->q;
      a.q_in = e->q;
      a.k = e->k;
      a.v = e->v;
      a.q_in_stride = Dq_l;
      a.kv_stride = Dkv_l;
      a.q_norm_w = w.q_norm;
      a.k_norm_w = w.k_norm;
      a.cos_t = e->cos_d;
      a.sin_t = e->sin_d;
      a.pos = e->pos_d;
      a.slot = e->slot_d;
      a.block_table = e->block_table_d;
      a.max_pages = e->max_pages_per_seq;
      a.n_tok = n;
      a.n_q = pl.n_q;
      a.layer = l;
      a.kv = e->kv;  // local kv heads
      QIE_TRY(KK_QKV_POST, launch_qkv_post(a, st));
    }
    {
      AttnArgs a{};
      a.q = e->q;
      a.out = e->att;
      a.pos = e->pos_d;
      a.slot = e->slot_d;
      a.block_table = e->block_table_d;
      a.max_pages = e->max_pages_per_seq;
      a.n_tok = n;
      a.n_q = pl.n_q;
      a.layer = l;
      a.max_kv_len = max_kv_len;
      a.kv = e->kv;
      QIE_TRY(KK_ATTN, launch_attention_ref(a, st));
    }
    {  // o_proj over the local head columns -> partial sums -> all-reduce -> residual
      GemmArgs g{};
      g.A = e->att;
      g.lda = Dq_l;
      g.K = Dq_l;
      g.nseg = 1;
      g.epi = EPI_STORE_F32;
      g.seg[0] = GemmSeg{w.o + pl.q_row0, nullptr, reinterpret_cast<bf16*>(e->tp_buf), H, H, Dq};
      cudaError_t r = gemm_rows(e, KK_GEMM_O, g, n, Dq_l);
      if (r != cudaSuccess) return r;
      r = tp_allreduce_f32(&e->tp, e->tp_buf, (size_t)n * H, st);
      if (r != cudaSuccess) return r;
      QIE_TRY(KK_GEMM_O, launch_residual_add_f32(e->x, e->tp_buf, (size_t)n * H, st));
    }
    QIE_TRY(KK_RMSNORM, launch_rmsnorm_ref(e->x, w.post_ln, e->xn, H, n, H, st));
    {
      GemmArgs g{};
      g.A = e->xn;
      g.lda = H;
      g.K = H;
      g.nseg = 1;
      g.epi = EPI_SILU_MUL;
      g.seg[0] = GemmSeg{w.gate + (size_t)pl.inter0 * H, w.up + (size_t)pl.inter0 * H, e->h, I_l, I_l, 0};
      cudaError_t r = gemm_rows(e, KK_GEMM_GATEUP, g, n, H);
      if (r != cudaSuccess) return r;
    }
    {  // down_proj over the local intermediate columns -> all-reduce -> residual
      GemmArgs g{};
      g.A = e->h;
      g.lda = I_l;
      g.K = I_l;
      g.nseg = 1;
      g.epi = EPI_STORE_F32;
      g.seg[0] = GemmSeg{w.down + pl.inter0, nullptr, reinterpret_cast<bf16*>(e->tp_buf), H, H, I};
      cudaError_t r = gemm_rows(e, KK_GEMM_DOWN, g, n, I_l);
      if (r != cudaSuccess) return r;
      r = tp_allreduce_f32(&e->tp, e->tp_buf, (size_t)n * H, st);
      if (r != cudaSuccess) return r;
      QIE_TRY(KK_GEMM_DOWN, launch_residual_add_f32(e->x, e->tp_buf, (size_t)n * H, st));
    }
  }
  if (n_out == 0) return cudaSuccess;
  QIE_TRY(KK_RMSNORM, launch_rmsnorm_ref(e->x + (size_t)out_row0 * H, e->final_norm, e->xn, H, n_out, H, st));
  {
    GemmArgs g{};
    g.A = e->xn;
    g.lda = H;
    g.K = H;
    g.nseg = 1;
    g.epi = EPI_STORE;
    g.seg[0] = GemmSeg{e->lm_head + (size_t)pl.vocab0 * H, nullptr, e->logits, V_l, V_l, 0};
    cudaError_t r = gemm_rows(e, KK_LM_HEAD, g, n_out, H);
    if (r != cudaSuccess) return r;
  }
  // local arg-max (reference tie-break; vocab/tp is a multiple of 256, so the low byte of the index is
  // the same locally and globally), then the best candidate over the ranks
  QIE_TRY(KK_SAMPLE, launch_sample_topk(e->logits, e->sampled_d, n_out, V_l, 1.0f, 1, e->seed, 0, nullptr, st));
  const int R = e->opts.max_batch_tokens;
  QIE_TRY(KK_SAMPLE, launch_tp_cand_make(e->logits, e->sampled_d, e->tp_cand, n_out, V_l, pl.vocab0, st));
  {
    cudaError_t r = tp_allgather(&e->tp, e->tp_cand, e->tp_cand + R, (size_t)n_out * sizeof(TpCand), st);
    if (r != cudaSuccess) return r;
  }
  QIE_TRY(KK_SAMPLE, launch_tp_cand_merge(e->tp_cand + R, e->tp.size, n_out, e->sampled_d, st));
  if (advance) QIE_TRY(KK_ADVANCE, launch_advance(e->pos_d, e->ids_d, e->sampled_d, n_out, e->rowstep_d, st));
  return cudaSuccess;
}

bool decode_uses_mega(const qie_engine* e, int n, int max_kv_len) {
  if (!e->use_mega || e->capture || !e->mega_layers_d || e->layer_count > 0 || e->inject_x) return false;
  if (e->has_bias || e->rope_half || e->eps != 1e-04f) return false;  // the persistent kernel implements the reference's semantics only
  const qie_config& c = e->cfg;
  if (e->tp.size > 1) {
    // tensor parallel: the persistent kernel runs this rank's shard and exchanges partial sums over NVLink peer
    // mappings (<= MEGA_TP_ROWS rows, greedy sampling, reference-order numerics)
    if (!e->tp_mega_ready || n > MEGA_TP_ROWS || e->topk != 1) return false;
    const TpPlan& pl = e->plan;
    return decode_mega_supports(c.hidden, pl.inter, c.layers, pl.n_q, pl.n_kv, c.head_dim, n, max_kv_len, e->num_sms,
                                e->mega_kc[mega_tile_set(e, n, max_kv_len)], 0, e->kv.page_size);
  }
  const int fast = e->opts.numerics == QIE_NUMERICS_FAST;  // fast numerics: persistent kernel for <= 8 rows
  // more rows than one launch takes (BASELINE configs[3]: 256 sequences on one GPU) run as consecutive launches of
  // <= MEGA_ROWS_PER_LAUNCH rows: both block sizes must fit
  const int nb = std::min(n, MEGA_ROWS_PER_LAUNCH), tail = n % MEGA_ROWS_PER_LAUNCH;
  if (n > MEGA_ROWS_PER_LAUNCH && (fast || e->mega_layers_run > 0)) return false;
  if (tail && n > MEGA_ROWS_PER_LAUNCH &&
      !decode_mega_supports(c.hidden, c.inter, c.layers, c.n_q, c.n_kv, c.head_dim, tail, max_kv_len, e->num_sms,
                            e->mega_kc[mega_tile_set(e, tail, max_kv_len)], fast, e->kv.page_size))
    return false;
  return decode_mega_supports(c.hidden, c.inter, c.layers, c.n_q, c.n_kv, c.head_dim, nb, max_kv_len, e->num_sms,
                              e->mega_kc[mega_tile_set(e, nb, max_kv_len)], fast, e->kv.page_size);
}

static cudaError_t forward_decode_mega_rows(qie_engine* e, int row0, int n, int max_kv_len, float temperature);

cudaError_t forward_decode_mega(qie_engine* e, int n_total, int max_kv_len, float temperature) {
  // row blocks: every block is one cooperative launch over its own slice of the row metadata / logits / samples;
  // the activation buffers are reused (the launches are stream-ordered)
  for (int row0 = 0; row0 < n_total; row0 += MEGA_ROWS_PER_LAUNCH) {
    cudaError_t r = forward_decode_mega_rows(e, row0, std::min(MEGA_ROWS_PER_LAUNCH, n_total - row0), max_kv_len, temperature);
    if (r != cudaSuccess) return r;
  }
  const qie_config& c = e->cfg;
  const bool penal = e->hist_d && e->rep_penalty != 1.0f;
  if (e->tp.size <= 1 && (e->topk != 1 || penal) && e->mega_layers_run <= 0) {
    // top-k > 1 / repetition penalty: the kernel left the logits; penalty, then the sampler (radix-select top-k +
    // the reference's XORWOW draw, logit_decode.cu:149-274), then the step bookkeeping
    if (penal)
      QIE_TRY(KK_SAMPLE, launch_repetition_penalty(e->logits, e->hist_d, e->slot_d, e->pos_d, 1, 0, (size_t)max_kv_len, (size_t)c.context,
                                                   n_total, c.vocab, e->rep_penalty, e->stream));
    QIE_TRY(KK_SAMPLE, launch_sample_topk(e->logits, e->sampled_d, n_total, c.vocab, temperature, e->topk, e->seed, 0,
                                          e->add_step ? e->rowstep_d : nullptr, e->stream));
    QIE_TRY(KK_ADVANCE, launch_advance(e->pos_d, e->ids_d, e->sampled_d, n_total, e->rowstep_d, e->stream));
  }
  return cudaSuccess;
}

static cudaError_t forward_decode_mega_rows(qie_engine* e, int row0, int n, int max_kv_len, float temperature) {
  const qie_config& c = e->cfg;
  if (e->hist_d)
    QIE_TRY(KK_EMBED, launch_history_append(e->hist_d, (size_t)c.context, e->ids_d + row0, e->pos_d + row0, e->slot_d + row0, n, e->stream));
  MegaArgs a{};
  a.H = c.hidden;
  a.I = c.inter;
  a.L = c.layers;
  a.n_q = c.n_q;
  a.n_kv = c.n_kv;
  a.hd = c.head_dim;
  a.V = c.vocab;
  a.layers = e->mega_layers_d;
  const int tset = mega_tile_set(e, n, max_kv_len);
  a.wmaps = e->mega_wmaps_d[tset];
  a.KC = e->mega_kc[tset];
  a.kvmap = e->mega_kvmap_d;
  a.hmap = tset == 0 && !(e->tp.size > 1) ? e->mega_hmap_d : nullptr;  // boxes of KC = mega_kc[0] columns over the unsharded h
  a.embed = e->embed;
  a.final_norm = e->final_norm;
  a.lm_head = e->lm_head;
  a.cos_t = e->cos_d;
  a.sin_t = e->sin_d;
  a.B = n;
  a.ids = e->ids_d + row0;
  a.pos = e->pos_d + row0;
  a.slot = e->slot_d + row0;
  a.block_table = e->block_table_d;
  a.max_pages = e->max_pages_per_seq;
  a.max_kv_len = max_kv_len;
  a.rowstep = e->rowstep_d + row0;
  a.kv = e->kv;
  a.x = e->x;
  a.xn = e->xn;
  a.qkv = e->qkv;
  a.att = e->att;
  a.h = e->h;
  a.logits = e->logits + (size_t)row0 * c.vocab;
  a.cand = e->mega_cand_d;
  a.sampled = e->sampled_d + row0;
  a.bar = e->mega_bar_d;
  a.prof = e->mega_prof_on ? e->mega_prof_d : nullptr;
  a.prof_stride = 16 * c.layers + 8;
  a.greedy = e->topk == 1 && !(e->hist_d && e->rep_penalty != 1.0f);
  a.fast = e->opts.numerics == QIE_NUMERICS_FAST ? 1 : 0;
  {
    // measured on B200 (batch 64 / ctx 2048): no gain -- the attention phase is bound by shared-memory
    // broadcast loads and FP32 issue, not by DRAM latency -- so the prefetch stays opt-in (QIE_MEGA_KVPF=1)
    static const bool pf = [] { const char* v = getenv("QIE_MEGA_KVPF"); return v && v[0] == '1'; }();
    a.kv_l2_prefetch = pf ? 1 : 0;
  }
  a.advance = 1;
  a.n_layers_run = e->mega_layers_run;
  const bool tp = e->tp.size > 1;
  if (tp) {
    // this rank's shard: local heads / intermediate slice / vocabulary range, weight views of the shard
    const TpPlan& pl = e->plan;
    a.I = pl.inter;
    a.n_q = pl.n_q;
    a.n_kv = pl.n_kv;
    a.V = pl.vocab;
    a.wmaps = e->mega_wmaps_tp_d[tset];
    a.lm_head = e->lm_head + (size_t)pl.vocab0 * c.hidden;
    a.greedy = 1;  // decode_uses_mega admits tensor parallel only with top-k 1
    a.fast = 0;
    a.tp_size = e->tp.size;
    a.tp_rank = e->tp.rank;
    a.tp_vocab0 = pl.vocab0;
    for (int r = 0; r < e->tp.size; ++r) {
      char* base = reinterpret_cast<char*>(e->tp_peer_xbuf[r]);
      a.tp_flag[r] = reinterpret_cast<unsigned*>(base);
      a.tp_cand[r] = reinterpret_cast<MegaCand*>(base + 256);  // [tp][MEGA_TP_ROWS] x 8 bytes <= 2 KiB < MEGA_TP_HEADER
      a.tp_part[r] = reinterpret_cast<float*>(base + MEGA_TP_HEADER);
    }
    a.tp_epoch = reinterpret_cast<unsigned*>(e->tp_xbuf) + 32;  // byte 128 of the header: local generation base
    a.x2 = e->x2;
  }
  // fast numerics, <= DECODE_GEMV_MAX_ROWS rows: the GEMV kernel (decode_gemv.cu) -- contiguous row ranges per CTA,
  // warp per group of weight rows, split-KV flash decoding, phases chained through polled per-layer buffers instead of
  // grid barriers; same step semantics (logits, arg-max, bookkeeping)
  const bool gemv = a.fast && !tp && e->use_gemv && e->gemv_part_d && e->mega_layers_run <= 0 &&
                    decode_gemv_supports(a.H, a.I, a.L, a.n_q, a.n_kv, a.hd, n, max_kv_len, e->num_sms);
  cudaError_t r = gemv ? launch_decode_gemv(a, e->gemv_part_d, e->num_sms, e->stream, e->gemv_dataflow) : launch_decode_mega(a, e->num_sms, e->stream);
  if (tp) {
    // the ranks must stay on the same path: no silent fallback here.  The step is ONE launch per rank: the
    // arg-max over the ranks' vocabulary ranges and the step bookkeeping happen inside the kernel.
    if (r != cudaSuccess) return r;
    ++e->launches;
    return cudaSuccess;
  }
  if (r != cudaSuccess) {
    // e.g. the cooperative launch cannot be co-scheduled (SMs taken by another context): this
    // engine falls back to the per-operator GPU launches for good (still no CPU path).  A shape
    // is always run eagerly before it is captured into a graph, so this never happens mid-capture.
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(e->stream, &cs);
    if (cs != cudaStreamCaptureStatusNone) return r;
    (void)cudaGetLastError();
    if (row0 > 0) return r;  // earlier row blocks of this step have already advanced: no consistent fallback
    e->use_mega = false;
    return cudaErrorNotReady;  // the caller (decode_forward) re-runs the whole step through forward_rows
  }
  ++e->launches;
  (void)temperature;
  return cudaSuccess;
}

}  // namespace qie
