// engine.cu -- forward launch order + engine lifecycle. See engine.h.
#include "engine.h"

#include <math.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace qie {

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
    if (_e != cudaSuccess) return _e;                          \
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
    for (int s = 0; s < g.nseg; ++s) c.seg[s].out = g.seg[s].out + (size_t)r0 * g.seg[s].ld_out;
    QIE_TRY(kind, launch_gemm_ref_order(c, e->num_sms, e->stream));
  }
  return cudaSuccess;
}

// Launch order = llm()'s (qwen_main.cu:77-222 prefill / :271-365 decode): per layer
// rms -> q,k,v -> q/k-norm -> RoPE -> KV store -> attention -> o (+residual) -> rms ->
// up,gate -> SiLU*up -> down (+residual); then final norm -> lm_head -> sampling.
// 18 kernels + 2-4 memcpys per layer in the reference become 8 launches here.
cudaError_t forward_rows(qie_engine* e, int n, int max_kv_len, int out_row0, int n_out, float temperature,
                         bool advance) {
  const qie_config& c = e->cfg;
  const int H = c.hidden, hd = c.head_dim, Dq = c.n_q * hd, Dkv = c.n_kv * hd, I = c.inter;
  cudaStream_t st = e->stream;

  QIE_TRY(KK_EMBED, launch_embedding(e->x, e->embed, e->ids_d, H, n, st));
  for (int l = 0; l < c.layers; ++l) {
    const LayerWeights& w = e->L[l];
    QIE_TRY(KK_RMSNORM, launch_rmsnorm_ref(e->x, w.in_ln, e->xn, H, n, H, st));
    capture_copy(e, "input_norm", l, e->xn, (size_t)n * H);
    {
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
      a.k = e->k;
      a.v = e->v;
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
      QIE_TRY(KK_QKV_POST, launch_qkv_post(a, st));
    }
    capture_copy(e, "q", l, e->q, (size_t)n * Dq);
    capture_copy(e, "v", l, e->v, (size_t)n * Dkv);
    {
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
    {
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
    QIE_TRY(KK_RMSNORM, launch_rmsnorm_ref(e->x, w.post_ln, e->xn, H, n, H, st));
    {
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
    {
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
  if (n_out == 0) return cudaSuccess;  // non-final prefill chunk: only the KV cache is needed
  // final norm only on the rows that feed lm_head (qwen_main.cu:227-236, :367-372)
  QIE_TRY(KK_RMSNORM, launch_rmsnorm_ref(e->x + (size_t)out_row0 * H, e->final_norm, e->xn, H, n_out, H, st));
  {
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
  QIE_TRY(KK_SAMPLE, launch_sample_topk(e->logits, e->sampled_d, n_out, c.vocab, temperature, e->topk, e->seed, 0,
                             e->add_step ? e->rowstep_d : nullptr, st));
  if (advance) QIE_TRY(KK_ADVANCE, launch_advance(e->pos_d, e->ids_d, e->sampled_d, n_out, e->rowstep_d, st));
  return cudaSuccess;
}

}  // namespace qie
