// capi.cu -- the extern "C" boundary declared in include/qie_b200.h.
#include <math.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>

#include "engine.h"

using namespace qie;

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  return fail(QIE_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}
#define CU(expr)                                        \
  do {                                                  \
    cudaError_t _e = (expr);                            \
    if (_e != cudaSuccess) return cuda_fail(_e, #expr); \
  } while (0)

extern "C" {

const char* qie_last_error(void) { return g_err; }
int qie_abi_version(void) { return 5; }

// ---------------------------------------------------------------- operator level
int qie_embedding(qie_bf16* out, const qie_bf16* table, const int* ids, size_t hidden, size_t n_tok, qie_stream st) {
  CU(launch_embedding((bf16*)out, (const bf16*)table, ids, hidden, n_tok, (cudaStream_t)st));
  return QIE_OK;
}
int qie_rmsnorm(const qie_bf16* x, const qie_bf16* w, qie_bf16* y, size_t hidden, size_t n_tok, qie_stream st) {
  if (hidden == 0 || hidden > 48 * 1024 / 4) return fail(QIE_EINVAL, "rmsnorm: hidden %zu unsupported", hidden);
  CU(launch_rmsnorm_ref((const bf16*)x, (const bf16*)w, (bf16*)y, hidden, n_tok, hidden, (cudaStream_t)st));
  return QIE_OK;
}
int qie_matmul(const qie_bf16* A, const qie_bf16* B, qie_bf16* C, int M, int N, int K, qie_stream st) {
  // reference argument meaning (helpers.cuh:81): N = inner dim, K = output columns
  if (M < 0 || N <= 0 || K <= 0 || (N & 7)) return fail(QIE_EINVAL, "matmul: need N %% 8 == 0 (got M=%d N=%d K=%d)", M, N, K);
  int dev = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  for (int r0 = 0; r0 < M; r0 += 64) {
    GemmArgs g{};
    g.A = (const bf16*)A + (size_t)r0 * N;
    g.lda = N;
    g.M = std::min(64, M - r0);
    g.K = N;
    g.nseg = 1;
    g.epi = EPI_STORE;
    g.seg[0] = GemmSeg{(const bf16*)B, nullptr, (bf16*)C + (size_t)r0 * K, K, K};
    CU(launch_gemm_ref_order(g, sms, (cudaStream_t)st));
  }
  return QIE_OK;
}
int qie_qknorm(qie_bf16* x, const qie_bf16* w, int head_dim, int n_tok, int row_dim, int n_heads, qie_stream st) {
  if (head_dim != 64 && head_dim != 128 && head_dim != 256) return fail(QIE_EINVAL, "head_dim %d unsupported", head_dim);
  CU(launch_qknorm_ref((bf16*)x, (const bf16*)w, head_dim, n_tok, row_dim, n_heads, (cudaStream_t)st));
  return QIE_OK;
}
int qie_rope(const float* cos_t, const float* sin_t, qie_bf16* x, int n_tok, int pos0, int head_dim, int row_dim,
             int n_heads, qie_stream st) {
  if (head_dim != 64 && head_dim != 128 && head_dim != 256) return fail(QIE_EINVAL, "head_dim %d unsupported", head_dim);
  CU(launch_rope_ref(cos_t, sin_t, (bf16*)x, n_tok, nullptr, pos0, head_dim, row_dim, n_heads, (cudaStream_t)st));
  return QIE_OK;
}
int qie_precompute_cos_sin(float* h_cos, float* h_sin, int seq_len, int head_dim) {
  // src/include.cpp:5-18 -- float operands throughout, so powf / cosf / sinf
  if (!h_cos || !h_sin || seq_len < 0 || head_dim <= 0) return fail(QIE_EINVAL, "precompute_cos_sin: bad args");
  float base = 1000000;
  int half = head_dim / 2;
  for (int i = 0; i < half; i++) {
    float exponent = 2 * ((float)i / (float)head_dim);
    float theta = powf(base, -exponent);
    for (int pos = 0; pos < seq_len; pos++) {
      h_cos[(size_t)pos * half + i] = cosf(pos * theta);
      h_sin[(size_t)pos * half + i] = sinf(pos * theta);
    }
  }
  return QIE_OK;
}
int qie_silu(qie_bf16* x, size_t n, qie_stream st) {
  CU(launch_silu((bf16*)x, n, (cudaStream_t)st));
  return QIE_OK;
}
int qie_elem_mul(const qie_bf16* a, const qie_bf16* b, qie_bf16* c, size_t n, qie_stream st) {
  CU(launch_elem_mul((const bf16*)a, (const bf16*)b, (bf16*)c, n, (cudaStream_t)st));
  return QIE_OK;
}
int qie_residual_add(qie_bf16* a, const qie_bf16* b, size_t n, qie_stream st) {
  CU(launch_residual_add((bf16*)a, (const bf16*)b, n, (cudaStream_t)st));
  return QIE_OK;
}

static KvGeom geom_of(const qie_kv_view* v) {
  KvGeom g{};
  g.pool = (bf16*)v->pool;
  g.n_pages = v->n_pages;
  g.page_size = v->page_size;
  g.n_layers = v->n_layers;
  g.n_kv = v->n_kv_heads;
  g.hd = v->head_dim;
  return g;
}
int qie_kv_store(const qie_kv_view* kv, int layer, const qie_bf16* K, const qie_bf16* V, const int* pos,
                 const int* slot, const int* block_table, int max_pages, int n_tok, qie_stream st) {
  if (!kv || layer < 0 || layer >= kv->n_layers) return fail(QIE_EINVAL, "kv_store: bad layer");
  CU(launch_kv_store(geom_of(kv), layer, (const bf16*)K, (const bf16*)V, pos, slot, block_table, max_pages, n_tok,
                     (cudaStream_t)st));
  return QIE_OK;
}
int qie_attention(const qie_kv_view* kv, int layer, const qie_bf16* Q, qie_bf16* out, const int* pos, const int* slot,
                  const int* block_table, int max_pages, int n_tok, int n_q_heads, qie_stream st) {
  if (!kv || layer < 0 || layer >= kv->n_layers) return fail(QIE_EINVAL, "attention: bad layer");
  AttnArgs a{};
  a.q = (const bf16*)Q;
  a.out = (bf16*)out;
  a.pos = pos;
  a.slot = slot;
  a.block_table = block_table;
  a.max_pages = max_pages;
  a.n_tok = n_tok;
  a.n_q = n_q_heads;
  a.layer = layer;
  a.max_kv_len = max_pages * kv->page_size;
  a.kv = geom_of(kv);
  CU(launch_attention_ref(a, (cudaStream_t)st));
  return QIE_OK;
}
int qie_attention_pagelist(const qie_bf16* const* d_k_pages, const qie_bf16* const* d_v_pages, int n_pages, int page_size,
                           int n_layers, int layer, int n_kv_heads, int head_dim, const qie_bf16* Q, qie_bf16* out, const int* pos,
                           int n_tok, int n_q_heads, qie_stream st) {
  if (!d_k_pages || !d_v_pages || n_pages <= 0 || page_size <= 0 || layer < 0 || layer >= n_layers || n_kv_heads <= 0 ||
      n_q_heads % n_kv_heads)
    return fail(QIE_EINVAL, "attention_pagelist: bad geometry");
  AttnArgs a{};
  a.q = (const bf16*)Q;
  a.out = (bf16*)out;
  a.pos = pos;
  a.n_tok = n_tok;
  a.n_q = n_q_heads;
  a.layer = layer;
  a.max_kv_len = n_pages * page_size;
  a.kv.page_size = page_size;
  a.kv.n_kv = n_kv_heads;
  a.kv.hd = head_dim;
  a.kv.n_layers = n_layers;
  a.k_pages = reinterpret_cast<const bf16* const*>(d_k_pages);
  a.v_pages = reinterpret_cast<const bf16* const*>(d_v_pages);
  a.pl_layers = n_layers;
  CU(launch_attention_ref(a, (cudaStream_t)st));
  return QIE_OK;
}
int qie_kv_store_pagelist(qie_bf16* const* d_k_pages, qie_bf16* const* d_v_pages, int n_pages, int page_size, int n_layers,
                          int layer, int kv_dim, const qie_bf16* K, const qie_bf16* V, int pos0, int n_tok, qie_stream st) {
  if (!d_k_pages || !d_v_pages || page_size <= 0 || layer < 0 || layer >= n_layers || pos0 < 0 ||
      (size_t)pos0 + n_tok > (size_t)n_pages * page_size)
    return fail(QIE_EINVAL, "kv_store_pagelist: rows [%d, %d) do not fit %d pages of %d", pos0, pos0 + n_tok, n_pages, page_size);
  CU(launch_kv_store_pagelist(reinterpret_cast<bf16* const*>(d_k_pages), reinterpret_cast<bf16* const*>(d_v_pages), page_size,
                              n_layers, layer, kv_dim, (const bf16*)K, (const bf16*)V, pos0, n_tok, (cudaStream_t)st));
  return QIE_OK;
}
int qie_sample_topk(const qie_bf16* logits, int* out_tokens, int n_rows, size_t vocab, float temperature, int k,
                    uint64_t seed, uint64_t seed_stride, qie_stream st) {
  CU(launch_sample_topk((const bf16*)logits, out_tokens, n_rows, vocab, temperature, k, seed, seed_stride, nullptr,
                        (cudaStream_t)st));
  return QIE_OK;
}

int qie_sample_topk_subseq(const qie_bf16* logits, int* out_tokens, int n_rows, size_t vocab, float temperature, int k,
                           uint64_t seed, uint64_t seed_stride, uint64_t subsequence, qie_stream st) {
  CU(launch_sample_topk((const bf16*)logits, out_tokens, n_rows, vocab, temperature, k, seed, seed_stride, nullptr,
                        (cudaStream_t)st, subsequence));
  return QIE_OK;
}

int qie_repetition_penalty(qie_bf16* logits, const int* context_tokens, size_t context_len, int vocab_size, float penalty,
                           qie_stream st) {
  if (!logits || (!context_tokens && context_len) || vocab_size <= 0 || !(penalty > 0.0f)) return fail(QIE_EINVAL, "repetition_penalty: bad argument");
  CU(launch_repetition_penalty((bf16*)logits, context_tokens, nullptr, nullptr, 0, context_len, context_len, 0, 1, vocab_size, penalty,
                               (cudaStream_t)st));
  return QIE_OK;
}

static int scratch(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return QIE_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  CU(cudaMalloc(p, need));
  *cap = need;
  return QIE_OK;
}

int qie_matmul_fast(const qie_bf16* A, const qie_bf16* B, qie_bf16* C, int M, int N, int K, qie_stream st) {
  if (M < 1 || N < 64 || (N % 8) || K < 1) return fail(QIE_EINVAL, "matmul_fast: need M>=1, N>=64, N%%8==0");
  static thread_local void* ws = nullptr;
  static thread_local size_t ws_cap = 0;
  static thread_local int* counters = nullptr;
  if (!counters) {
    CU(cudaMalloc(&counters, 8192 * sizeof(int)));
    CU(cudaMemset(counters, 0, 8192 * sizeof(int)));
  }
  int dev = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int rc = scratch(&ws, &ws_cap, std::max<size_t>((size_t)32 << 20, (size_t)M * K * 4 * 4));
  if (rc) return rc;
  TensorMap2D wmap, xmap;
  CU(make_tensor_map_2d(&wmap, (const bf16*)B, K, N, 128));
  CU(make_tensor_map_2d(&xmap, (const bf16*)A, M, N, tc_token_tile(M)));
  TcGemm t{};
  t.w[0] = &wmap;
  t.x = &xmap;
  t.rows[0] = K;
  t.nseg = 1;
  t.M = M;
  t.K = N;
  t.epi = EPI_STORE;
  t.out = (bf16*)C;
  t.ld_out = K;
  t.ws = (float*)ws;
  t.ws_bytes = ws_cap;
  t.max_splits = (int)std::max<size_t>(1, ws_cap / ((size_t)M * K * sizeof(float)));
  t.counters = counters;
  t.n_counters = 8192;
  t.w_static = 0;  // B may have been written by the caller's previous kernel
  CU(launch_gemm_tcgen05(t, sms, (cudaStream_t)st, nullptr));
  return QIE_OK;
}

int qie_attention_decode_fast(const qie_kv_view* kv, int layer, const qie_bf16* Q, qie_bf16* out, const int* pos,
                              const int* slot, const int* block_table, int max_pages, int n_tok, int n_q_heads,
                              int n_splits, qie_stream st) {
  if (!kv || layer < 0 || layer >= kv->n_layers) return fail(QIE_EINVAL, "attention_decode_fast: bad layer");
  if (n_q_heads % kv->n_kv_heads || n_q_heads / kv->n_kv_heads > 16 || (kv->head_dim != 64 && kv->head_dim != 128))
    return fail(QIE_EINVAL, "attention_decode_fast: need n_q/n_kv <= 16 and head_dim 64 or 128");
  static thread_local void *wo = nullptr, *wml = nullptr;
  static thread_local size_t wo_cap = 0, wml_cap = 0;
  if (n_splits <= 0) n_splits = 4;
  n_splits = std::min(n_splits, 64);
  int rc = scratch(&wo, &wo_cap, (size_t)n_splits * n_tok * n_q_heads * kv->head_dim * 4);
  if (rc) return rc;
  rc = scratch(&wml, &wml_cap, (size_t)n_splits * n_tok * n_q_heads * 2 * 4);
  if (rc) return rc;
  FastAttnArgs a{};
  a.q = (const bf16*)Q;
  a.out = (bf16*)out;
  a.pos = pos;
  a.slot = slot;
  a.block_table = block_table;
  a.max_pages = max_pages;
  a.n_tok = n_tok;
  a.n_q = n_q_heads;
  a.layer = layer;
  a.n_splits = n_splits;
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)kv->head_dim);
  a.ws_o = (float*)wo;
  a.ws_ml = (float*)wml;
  a.kv = geom_of(kv);
  CU(launch_attention_decode_fast(a, (cudaStream_t)st));
  return QIE_OK;
}

int qie_attention_prefill_fast(const qie_kv_view* kv, int layer, const qie_bf16* Q, qie_bf16* out, const int* pos,
                               const int* slot, const int* block_table, int max_pages, int n_tok, int n_q_heads,
                               qie_stream st) {
  if (!kv || layer < 0 || layer >= kv->n_layers) return fail(QIE_EINVAL, "attention_prefill_fast: bad layer");
  if (n_q_heads % kv->n_kv_heads || (kv->head_dim != 64 && kv->head_dim != 128))
    return fail(QIE_EINVAL, "attention_prefill_fast: need n_q %% n_kv == 0 and head_dim 64 or 128");
  FastAttnArgs a{};
  a.q = (const bf16*)Q;
  a.out = (bf16*)out;
  a.pos = pos;
  a.slot = slot;
  a.block_table = block_table;
  a.max_pages = max_pages;
  a.n_tok = n_tok;
  a.n_q = n_q_heads;
  a.layer = layer;
  a.n_splits = 1;
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)kv->head_dim);
  a.kv = geom_of(kv);
  CU(launch_attention_prefill_fast(a, (cudaStream_t)st));
  return QIE_OK;
}

int qie_attention_prefill_tc(const qie_kv_view* kv, int layer, const qie_bf16* Q, qie_bf16* out, const int* pos,
                             const int* slot, const int* block_table, int max_pages, int n_tok, int n_q_heads, int variant,
                             qie_stream st) {
  if (!kv || layer < 0 || layer >= kv->n_layers) return fail(QIE_EINVAL, "attention_prefill_tc: bad layer");
  if (n_q_heads % kv->n_kv_heads || kv->head_dim != 128) return fail(QIE_EINVAL, "attention_prefill_tc: head_dim must be 128");
  FastAttnArgs a{};
  a.q = (const bf16*)Q;
  a.out = (bf16*)out;
  a.pos = pos;
  a.slot = slot;
  a.block_table = block_table;
  a.max_pages = max_pages;
  a.n_tok = n_tok;
  a.n_q = n_q_heads;
  a.layer = layer;
  a.n_splits = 1;
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)kv->head_dim);
  a.kv = geom_of(kv);
  CU(launch_attention_prefill_tc(a, variant, (cudaStream_t)st));
  return QIE_OK;
}

// ---------------------------------------------------------------- driver level
void qie_engine_opts_default(qie_engine_opts* o) {
  memset(o, 0, sizeof(*o));
  o->device = 0;
  o->page_size = 16;
  o->max_pages = 0;
  o->kv_bytes = 0;
  o->max_seqs = 64;
  o->max_batch_tokens = 256;
  o->context = 0;
  o->head_dim_hint = 0;
  o->use_graph = 1;
  o->tp_rank = 0;
  o->tp_size = 1;
  o->numerics = QIE_NUMERICS_REFERENCE_ORDER;
}

int qie_synth_checkpoint_write(const qie_config* cfg, uint64_t seed, const char* meta_path,
                               const char* weights_path) {
  if (!cfg || !meta_path || !weights_path) return fail(QIE_EINVAL, "synth: null argument");
  Checkpoint ck = synth_layout(*cfg);
  FILE* fm = fopen(meta_path, "w");
  if (!fm) return fail(QIE_EIO, "cannot write %s", meta_path);
  write_meta(ck, fm);
  fclose(fm);
  FILE* fw = fopen(weights_path, "wb");
  if (!fw) return fail(QIE_EIO, "cannot write %s", weights_path);
  std::vector<uint16_t> buf(1 << 20);
  for (const TensorInfo& t : ck.tensors) {
    size_t elems = (t.end - t.begin) / 2, g0 = t.begin / 2;
    for (size_t e0 = 0; e0 < elems; e0 += buf.size()) {
      size_t m = std::min(buf.size(), elems - e0);
      for (size_t j = 0; j < m; ++j) buf[j] = synth_value(seed, g0 + e0 + j, t.kind);
      if (fwrite(buf.data(), 2, m, fw) != m) {
        fclose(fw);
        return fail(QIE_EIO, "short write to %s", weights_path);
      }
    }
  }
  fclose(fw);
  return QIE_OK;
}

int qie_checkpoint_inspect(const char* meta_path, int head_dim_hint, qie_config* cfg_out, size_t* total_bytes,
                           int* n_tensors) {
  if (!meta_path || !cfg_out) return fail(QIE_EINVAL, "checkpoint_inspect: null argument");
  Checkpoint ck;
  std::string err;
  if (!parse_meta(meta_path, &ck, &err)) return fail(QIE_EIO, "%s", err.c_str());
  if (!derive_config(ck, head_dim_hint, 32786, cfg_out, &err)) return fail(QIE_EIO, "%s", err.c_str());
  if (total_bytes) *total_bytes = ck.total_bytes;
  if (n_tensors) *n_tensors = (int)ck.tensors.size();
  return QIE_OK;
}

int qie_convert_safetensors(const char* const* shard_paths, int n_shards, const char* meta_path, const char* weights_path,
                            int tie_lm_head, size_t* total_bytes, int* n_tensors) {
  if (!shard_paths || n_shards <= 0 || !meta_path || !weights_path) return fail(QIE_EINVAL, "convert_safetensors: null argument");
  std::vector<std::string> shards;
  for (int i = 0; i < n_shards; ++i) {
    if (!shard_paths[i]) return fail(QIE_EINVAL, "convert_safetensors: null shard path");
    shards.push_back(shard_paths[i]);
  }
  const std::string err = convert_safetensors(shards, meta_path, weights_path, tie_lm_head != 0, total_bytes, n_tensors);
  if (!err.empty()) return fail(QIE_EIO, "%s", err.c_str());
  return QIE_OK;
}

static void engine_free(qie_engine* e) {
  if (e && e->hist_d) cudaFree(e->hist_d);
  if (!e) return;
  cudaSetDevice(e->opts.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (auto& kvp : e->graphs)
    if (kvp.second.exec) cudaGraphExecDestroy(kvp.second.exec);
  for (auto& kvp : e->cap)
    if (kvp.second.d) cudaFree(kvp.second.d);
  if (!e->blob_owned) e->blob = nullptr;
  void* dev[] = {e->blob, e->cos_d, e->sin_d, e->kv.pool, e->block_table_d, e->ids_d, e->pos_d, e->slot_d,
                 e->sampled_d, e->rowstep_d, e->x, e->xn, e->q, e->k, e->v, e->att, e->h, e->logits, e->qkv,
                 e->gemm_ws, e->attn_ws_o, e->attn_ws_ml, e->gemm_counters, e->mega_layers_d, e->mega_cand_d,
                 e->mega_bar_d, e->mega_prof_d, e->mega_wmaps_d[0], e->mega_wmaps_d[1], e->tp_buf, e->tp_cand, e->tp_xbuf, e->x2,
                 e->mega_wmaps_tp_d[0], e->mega_wmaps_tp_d[1], e->mega_kvmap_d, e->mega_hmap_d, e->gemv_part_d};
  for (int r = 0; r < MEGA_MAX_TP; ++r)
    if (e->tp_peer_xbuf[r] && e->tp_peer_xbuf[r] != e->tp_xbuf) cudaIpcCloseMemHandle(e->tp_peer_xbuf[r]);
  tp_comm_destroy(&e->tp);
  for (void* p : dev)
    if (p) cudaFree(p);
  if (e->block_table_h) cudaFreeHost(e->block_table_h);
  if (e->stage_h) cudaFreeHost(e->stage_h);
  if (e->sampled_h) cudaFreeHost(e->sampled_h);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

static int engine_finish_setup(qie_engine* e) {
  const qie_config& c = e->cfg;
  const qie_engine_opts& o = e->opts;
  if (c.head_dim != 64 && c.head_dim != 128 && c.head_dim != 256)
    return fail(QIE_EINVAL, "head_dim %d unsupported (64/128/256)", c.head_dim);
  if ((c.hidden & 7) || (c.inter & 7)) return fail(QIE_EINVAL, "hidden/inter must be multiples of 8");
  // weight pointers (assign_weight_pointer, helpers.cuh:18-29)
  auto ptr = [&](const char* sn, int layer) -> const bf16* {
    const TensorInfo* t = e->ck.find(sn, layer);
    return t ? reinterpret_cast<const bf16*>(reinterpret_cast<const char*>(e->blob) + t->begin) : nullptr;
  };
  e->L.resize(c.layers);
  for (int l = 0; l < c.layers; ++l) {
    LayerWeights& w = e->L[l];
    w.in_ln = ptr("input_layernorm.weight", l);
    w.q = ptr("self_attn.q_proj.weight", l);
    w.k = ptr("self_attn.k_proj.weight", l);
    w.v = ptr("self_attn.v_proj.weight", l);
    w.o = ptr("self_attn.o_proj.weight", l);
    w.q_norm = ptr("self_attn.q_norm.weight", l);  // data-driven: absent in Qwen2.5 checkpoints
    w.k_norm = ptr("self_attn.k_norm.weight", l);
    w.post_ln = ptr("post_attention_layernorm.weight", l);
    w.up = ptr("mlp.up_proj.weight", l);
    w.gate = ptr("mlp.gate_proj.weight", l);
    w.down = ptr("mlp.down_proj.weight", l);
    w.q_bias = ptr("self_attn.q_proj.bias", l);  // data-driven: present in Qwen2.5 checkpoints
    w.k_bias = ptr("self_attn.k_proj.bias", l);
    w.v_bias = ptr("self_attn.v_proj.bias", l);
    if (w.q_bias || w.k_bias || w.v_bias) e->has_bias = true;
    if (!w.in_ln || !w.q || !w.k || !w.v || !w.o || !w.post_ln || !w.up || !w.gate || !w.down)
      return fail(QIE_EIO, "layer %d: missing tensor in checkpoint", l);
  }
  // biases the forward does not apply would give wrong logits without an error: refuse them
  for (const TensorInfo& t : e->ck.tensors) {
    const std::string& sn = t.short_name;
    const bool is_bias = sn.size() > 5 && sn.compare(sn.size() - 5, 5, ".bias") == 0;
    if (is_bias && sn != "self_attn.q_proj.bias" && sn != "self_attn.k_proj.bias" && sn != "self_attn.v_proj.bias")
      return fail(QIE_EIO, "checkpoint tensor %s: this bias is not applied by the forward", t.name.c_str());
  }
  e->rope_half = e->opts.semantics == QIE_SEMANTICS_HF;
  e->eps = e->opts.rms_eps > 0.0f ? e->opts.rms_eps : (e->opts.semantics == QIE_SEMANTICS_HF ? 1e-06f : 1e-04f);
  if (e->opts.tp_size > 1 && (e->has_bias || e->rope_half || e->eps != 1e-04f))
    return fail(QIE_EINVAL, "HF semantics / projection biases are not built for tensor parallelism");
  e->embed = ptr("embed_tokens.weight", -1);
  e->final_norm = ptr("norm.weight", -1);
  e->lm_head = ptr("logits", -1);
  if (!e->embed || !e->final_norm || !e->lm_head) return fail(QIE_EIO, "missing embed/norm/lm_head tensor");

  // RoPE tables on the host with libm, exactly as the reference (include.cpp:5-18)
  {
    size_t nel = (size_t)c.context * (c.head_dim / 2);
    std::vector<float> hc(nel), hs(nel);
    qie_precompute_cos_sin(hc.data(), hs.data(), c.context, c.head_dim);
    CU(cudaMalloc(&e->cos_d, nel * sizeof(float)));
    CU(cudaMalloc(&e->sin_d, nel * sizeof(float)));
    CU(cudaMemcpy(e->cos_d, hc.data(), nel * sizeof(float), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->sin_d, hs.data(), nel * sizeof(float), cudaMemcpyHostToDevice));
  }
  // KV pool
  e->kv.page_size = o.page_size;
  e->kv.n_layers = c.layers;
  e->kv.n_kv = c.n_kv;
  e->kv.hd = c.head_dim;
  if (o.tp_size > 1) {  // tensor parallel: this rank caches only its own kv heads
    if (!tp_plan(c, o.tp_rank, o.tp_size, &e->plan))
      return fail(QIE_EINVAL, "tp_size %d does not divide heads (%d q / %d kv) / intermediate %d", o.tp_size, c.n_q, c.n_kv, c.inter);
    e->tp.rank = o.tp_rank;
    e->tp.size = o.tp_size;  // comm stays null until qie_engine_tp_connect
    e->kv.n_kv = e->plan.n_kv;
  }
  size_t page_bytes = e->kv.page_stride() * sizeof(bf16);
  int n_pages = o.max_pages;
  if (n_pages <= 0) {
    size_t bytes = o.kv_bytes ? o.kv_bytes : ((size_t)1 << 30);
    n_pages = (int)std::max<size_t>(1, bytes / page_bytes);
  }
  e->kv.n_pages = n_pages;
  CU(cudaMalloc(&e->kv.pool, (size_t)n_pages * page_bytes));
  CU(cudaMemsetAsync(e->kv.pool, 0, (size_t)n_pages * page_bytes, e->stream));
  e->free_pages.resize(n_pages);
  for (int i = 0; i < n_pages; ++i) e->free_pages[i] = n_pages - 1 - i;  // pop_back hands out 0,1,2,...
  e->seqs.assign(o.max_seqs, Sequence());
  e->max_pages_per_seq = std::min(n_pages, (c.context + o.page_size - 1) / o.page_size);
  size_t bt = (size_t)o.max_seqs * e->max_pages_per_seq;
  CU(cudaMalloc(&e->block_table_d, bt * sizeof(int)));
  CU(cudaMemsetAsync(e->block_table_d, 0, bt * sizeof(int), e->stream));
  CU(cudaMallocHost(&e->block_table_h, bt * sizeof(int)));
  memset(e->block_table_h, 0, bt * sizeof(int));
  // row metadata + activations
  const size_t R = o.max_batch_tokens;
  const size_t H = c.hidden, Dq = (size_t)c.n_q * c.head_dim, Dkv = (size_t)c.n_kv * c.head_dim, I = c.inter;
  CU(cudaMalloc(&e->ids_d, R * sizeof(int)));
  CU(cudaMalloc(&e->pos_d, R * sizeof(int)));
  CU(cudaMalloc(&e->slot_d, R * sizeof(int)));
  CU(cudaMalloc(&e->sampled_d, R * sizeof(int)));
  CU(cudaMalloc(&e->rowstep_d, R * sizeof(int)));
  CU(cudaMallocHost(&e->stage_h, 4 * R * sizeof(int)));
  e->sampled_h_cap = R;
  CU(cudaMallocHost(&e->sampled_h, e->sampled_h_cap * sizeof(int)));
  CU(cudaMalloc(&e->x, R * H * sizeof(bf16)));
  CU(cudaMalloc(&e->xn, R * std::max(H, Dq) * sizeof(bf16)));
  CU(cudaMalloc(&e->q, R * Dq * sizeof(bf16)));
  CU(cudaMalloc(&e->k, R * Dkv * sizeof(bf16)));
  CU(cudaMalloc(&e->v, R * Dkv * sizeof(bf16)));
  CU(cudaMalloc(&e->att, R * Dq * sizeof(bf16)));
  CU(cudaMalloc(&e->h, R * I * sizeof(bf16)));
  e->logits_rows = std::min<int>((int)R, o.max_seqs);
  CU(cudaMalloc(&e->logits, (size_t)e->logits_rows * c.vocab * sizeof(bf16)));
  CU(cudaMalloc(&e->qkv, R * (Dq + 2 * Dkv) * sizeof(bf16)));
  if (o.tp_size > 1) {
    CU(cudaMalloc(&e->tp_buf, R * H * sizeof(float)));
    CU(cudaMalloc(&e->tp_cand, (size_t)(1 + o.tp_size) * R * sizeof(TpCand)));
    if (o.tp_size <= MEGA_MAX_TP) {
      const size_t xb = MEGA_TP_HEADER + (size_t)2 * o.tp_size * MEGA_TP_ROWS * H * sizeof(float);
      CU(cudaMalloc(&e->tp_xbuf, xb));
      CU(cudaMemsetAsync(e->tp_xbuf, 0, xb, e->stream));
      CU(cudaMalloc(&e->x2, R * H * sizeof(bf16)));
    }
  }
  {
    const char* mv = getenv("QIE_MEGA");
    e->use_mega = !(mv && mv[0] == '0');
    std::vector<MegaLayer> ml(c.layers);
    bool aligned = true;
    for (int l = 0; l < c.layers; ++l) {
      const LayerWeights& w = e->L[l];
      ml[l] = MegaLayer{w.in_ln, w.q, w.k, w.v, w.o, w.q_norm, w.k_norm, w.post_ln, w.gate, w.up, w.down};
      const void* ps[] = {w.in_ln, w.q, w.k, w.v, w.o, w.post_ln, w.gate, w.up, w.down, w.q_norm, w.k_norm};
      for (const void* p : ps) aligned = aligned && (((uintptr_t)p & 15) == 0);
    }
    aligned = aligned && (((uintptr_t)e->lm_head & 15) == 0) && (((uintptr_t)e->embed & 15) == 0) &&
              (((uintptr_t)e->final_norm & 15) == 0) && (c.vocab % 2 == 0);
    const size_t Hs = c.hidden;
    aligned = aligned && (Hs % 64 == 0) && (I % 64 == 0) && (Dq % 64 == 0);
    if (aligned) {  // TMA needs 16-byte aligned weight rows and 64-wide k blocks; otherwise the per-operator path is used
      for (int ts = 0; ts < 2; ++ts) {
        e->mega_kc[ts] = decode_mega_kc(c.hidden, ts);
        std::vector<TensorMap2D> maps((size_t)7 * c.layers + 1);
        bool ok = true;
        const int kc = e->mega_kc[ts];
        for (int l = 0; l < c.layers && ok; ++l) {
          const LayerWeights& w = e->L[l];
          TensorMap2D* m = &maps[(size_t)7 * l];
          ok = ok && make_tensor_map_w3d(m + 0, w.q, (int)Dq, (int)H, kc) == cudaSuccess;
          ok = ok && make_tensor_map_w3d(m + 1, w.k, (int)Dkv, (int)H, kc) == cudaSuccess;
          ok = ok && make_tensor_map_w3d(m + 2, w.v, (int)Dkv, (int)H, kc) == cudaSuccess;
          ok = ok && make_tensor_map_w3d(m + 3, w.o, (int)H, (int)Dq, kc) == cudaSuccess;
          ok = ok && make_tensor_map_w3d(m + 4, w.gate, (int)I, (int)H, kc) == cudaSuccess;
          ok = ok && make_tensor_map_w3d(m + 5, w.up, (int)I, (int)H, kc) == cudaSuccess;
          ok = ok && make_tensor_map_w3d(m + 6, w.down, (int)H, (int)I, kc) == cudaSuccess;
        }
        ok = ok && make_tensor_map_w3d(&maps[(size_t)7 * c.layers], e->lm_head, c.vocab, (int)H, kc) == cudaSuccess;
        if (!ok) return fail(QIE_ECUDA, "cuTensorMapEncodeTiled failed for the weight views");
        CU(cudaMalloc(&e->mega_wmaps_d[ts], maps.size() * sizeof(TensorMap2D)));
        CU(cudaMemcpy(e->mega_wmaps_d[ts], maps.data(), maps.size() * sizeof(TensorMap2D), cudaMemcpyHostToDevice));
      }
      if (o.tp_size > 1 && o.tp_size <= MEGA_MAX_TP && (e->plan.inter % 64 == 0) && ((e->plan.n_q * c.head_dim) % 64 == 0)) {
        // this rank's shard as tensor-map views of the full blob: row slices for q/k/v/gate/up/lm_head, column
        // slices (row stride = full width) for o_proj/down_proj
        const TpPlan& pl = e->plan;
        const int Dq_l = pl.n_q * c.head_dim, Dkv_l = pl.n_kv * c.head_dim;
        for (int ts = 0; ts < 2; ++ts) {
          std::vector<TensorMap2D> maps((size_t)7 * c.layers + 1);
          bool ok = true;
          const int kc = e->mega_kc[ts];
          for (int l = 0; l < c.layers && ok; ++l) {
            const LayerWeights& w = e->L[l];
            TensorMap2D* m = &maps[(size_t)7 * l];
            ok = ok && make_tensor_map_w3d(m + 0, w.q + (size_t)pl.q_row0 * H, Dq_l, (int)H, kc) == cudaSuccess;
            ok = ok && make_tensor_map_w3d(m + 1, w.k + (size_t)pl.kv_row0 * H, Dkv_l, (int)H, kc) == cudaSuccess;
            ok = ok && make_tensor_map_w3d(m + 2, w.v + (size_t)pl.kv_row0 * H, Dkv_l, (int)H, kc) == cudaSuccess;
            ok = ok && make_tensor_map_w3d(m + 3, w.o + pl.q_row0, (int)H, Dq_l, kc, (int)Dq) == cudaSuccess;
            ok = ok && make_tensor_map_w3d(m + 4, w.gate + (size_t)pl.inter0 * H, pl.inter, (int)H, kc) == cudaSuccess;
            ok = ok && make_tensor_map_w3d(m + 5, w.up + (size_t)pl.inter0 * H, pl.inter, (int)H, kc) == cudaSuccess;
            ok = ok && make_tensor_map_w3d(m + 6, w.down + pl.inter0, (int)H, pl.inter, kc, (int)I) == cudaSuccess;
          }
          ok = ok && make_tensor_map_w3d(&maps[(size_t)7 * c.layers], e->lm_head + (size_t)pl.vocab0 * H, pl.vocab, (int)H, kc) == cudaSuccess;
          if (!ok) return fail(QIE_ECUDA, "cuTensorMapEncodeTiled failed for the tensor-parallel weight views");
          CU(cudaMalloc(&e->mega_wmaps_tp_d[ts], maps.size() * sizeof(TensorMap2D)));
          CU(cudaMemcpy(e->mega_wmaps_tp_d[ts], maps.data(), maps.size() * sizeof(TensorMap2D), cudaMemcpyHostToDevice));
        }
      }
      CU(cudaMalloc(&e->mega_layers_d, c.layers * sizeof(MegaLayer)));
      CU(cudaMemcpy(e->mega_layers_d, ml.data(), c.layers * sizeof(MegaLayer), cudaMemcpyHostToDevice));
      CU(cudaMalloc(&e->mega_cand_d, (size_t)e->num_sms * 64 * sizeof(MegaCand)));
      CU(cudaMalloc(&e->mega_bar_d, 4096));  // grid-barrier counters (8 shards on separate lines)
      CU(cudaMalloc(&e->mega_prof_d, (size_t)decode_mega_prof_slots(c.layers) * sizeof(unsigned long long)));
      {
        const size_t gb = decode_gemv_scratch_bytes(c.hidden, (int)I, c.layers, c.n_q, c.n_kv, c.head_dim, e->num_sms);
        CU(cudaMalloc(&e->gemv_part_d, gb));
        CU(cudaMemset(e->gemv_part_d, 0xFF, gb));  // "not stored yet" pattern of the GEMV kernel's per-layer buffers
      }
      {
        const char* gv = getenv("QIE_GEMV");
        e->use_gemv = !(gv && gv[0] == '0');
      }
      if ((I % 64) == 0 && e->h) {  // down_proj operand rows for the tile-split phase (tile set 0: batches > 8)
        TensorMap2D hm;
        if (make_tensor_map_w3d(&hm, e->h, (int)R, (int)I, e->mega_kc[0], 0, 16) == cudaSuccess) {
          CU(cudaMalloc(&e->mega_hmap_d, sizeof(TensorMap2D)));
          CU(cudaMemcpy(e->mega_hmap_d, &hm, sizeof(TensorMap2D), cudaMemcpyHostToDevice));
        }
      }
      {
        // K/V rows for the attention phase's TMA stream (head_dim 64, pages of a power of two >= 8 slots; other
        // geometries keep the cp.async tile loader)
        const int psz = e->kv.page_size;
        const unsigned long long rows = (unsigned long long)e->kv.n_pages * e->kv.n_layers * 2ull * e->kv.n_kv * psz;
        TensorMap2D km;
        if (c.head_dim == 64 && psz >= 8 && (psz & (psz - 1)) == 0 &&
            make_tensor_map_kv(&km, e->kv.pool, rows, c.head_dim, std::min(psz, 64)) == cudaSuccess) {
          CU(cudaMalloc(&e->mega_kvmap_d, sizeof(TensorMap2D)));
          CU(cudaMemcpy(e->mega_kvmap_d, &km, sizeof(TensorMap2D), cudaMemcpyHostToDevice));
        }
      }
    }
  }
  if (o.numerics == QIE_NUMERICS_FAST) {
    if (c.n_q / c.n_kv > 16 || c.head_dim > 128) return fail(QIE_EINVAL, "fast numerics: need n_q/n_kv <= 16 and head_dim <= 128");
    if ((c.hidden % 64) || (c.inter % 64) || ((c.n_q * c.head_dim) % 64)) return fail(QIE_EINVAL, "fast numerics: inner dims must be multiples of 64");
    e->wmaps.resize(c.layers);
    for (int l = 0; l < c.layers; ++l) {
      const LayerWeights& w = e->L[l];
      qie_engine::LayerMaps& m = e->wmaps[l];
      CU(make_tensor_map_2d(&m.q, w.q, (int)Dq, (int)H, 128));
      CU(make_tensor_map_2d(&m.k, w.k, (int)Dkv, (int)H, 128));
      CU(make_tensor_map_2d(&m.v, w.v, (int)Dkv, (int)H, 128));
      CU(make_tensor_map_2d(&m.o, w.o, (int)H, (int)Dq, 128));
      CU(make_tensor_map_2d(&m.gate, w.gate, (int)I, (int)H, 64));  // dual tile: 64 gate + 64 up rows
      CU(make_tensor_map_2d(&m.up, w.up, (int)I, (int)H, 64));
      CU(make_tensor_map_2d(&m.down, w.down, (int)H, (int)I, 128));
    }
    CU(make_tensor_map_2d(&e->lm_head_map, e->lm_head, c.vocab, (int)H, 128));
    e->gemm_ws_bytes = std::max<size_t>((size_t)64 << 20, R * 2 * I * sizeof(float) * 2);
    CU(cudaMalloc(&e->gemm_ws, e->gemm_ws_bytes));
    CU(cudaMalloc(&e->gemm_counters, 8192 * sizeof(int)));
    CU(cudaMemsetAsync(e->gemm_counters, 0, 8192 * sizeof(int), e->stream));
    CU(cudaMalloc(&e->attn_ws_o, (size_t)e->attn_max_splits * R * Dq * sizeof(float)));
    CU(cudaMalloc(&e->attn_ws_ml, (size_t)e->attn_max_splits * R * c.n_q * 2 * sizeof(float)));
  }
  CU(cudaStreamSynchronize(e->stream));
  return QIE_OK;
}

static int engine_begin(const qie_engine_opts* opts, qie_engine** out_e) {
  qie_engine_opts o;
  if (opts) o = *opts;
  else qie_engine_opts_default(&o);
  if (o.page_size <= 0) o.page_size = 16;
  if (o.max_seqs <= 0) o.max_seqs = 64;
  if (o.max_batch_tokens <= 0) o.max_batch_tokens = 256;
  if (o.context <= 0) o.context = 32786;  // sic, utills.cu:14
  if (o.tp_size <= 0) o.tp_size = 1;
  if (o.tp_rank < 0 || o.tp_rank >= o.tp_size) return fail(QIE_EINVAL, "tp_rank %d outside [0, %d)", o.tp_rank, o.tp_size);
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0)
    return fail(QIE_ECUDA, "no CUDA device (%s); libqie_b200 has no CPU fallback", cudaGetErrorString(ce));
  if (o.device < 0 || o.device >= ndev) return fail(QIE_EINVAL, "device %d out of range", o.device);
  CU(cudaSetDevice(o.device));
  qie_engine* e = new qie_engine();
  e->opts = o;
  cudaError_t r = cudaDeviceGetAttribute(&e->num_sms, cudaDevAttrMultiProcessorCount, o.device);
  if (r == cudaSuccess) r = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
  if (r != cudaSuccess) {
    delete e;
    return cuda_fail(r, "engine init");
  }
  *out_e = e;
  return QIE_OK;
}

int qie_engine_create(const char* meta_path, const char* weights_path, const qie_engine_opts* opts,
                      qie_engine** out) {
  if (!meta_path || !weights_path || !out) return fail(QIE_EINVAL, "engine_create: null argument");
  qie_engine* e = nullptr;
  int rc = engine_begin(opts, &e);
  if (rc) return rc;
  std::string err;
  if (!parse_meta(meta_path, &e->ck, &err) ||
      !derive_config(e->ck, e->opts.head_dim_hint, e->opts.context, &e->cfg, &err)) {
    engine_free(e);
    return fail(QIE_EIO, "%s", err.c_str());
  }
  // weights.bin -> one device blob, streamed through two pinned buffers
  // (load_all_weights_to_gpu_chunked, iengine.cu:117-223, used one pageable 2 GiB buffer)
  FILE* f = fopen(weights_path, "rb");
  if (!f) {
    engine_free(e);
    return fail(QIE_EIO, "cannot open %s", weights_path);
  }
  fseek(f, 0, SEEK_END);
  size_t fsize = (size_t)ftell(f);
  fseek(f, 0, SEEK_SET);
  if (fsize < e->ck.total_bytes) {
    fclose(f);
    engine_free(e);
    return fail(QIE_EIO, "%s is %zu bytes, meta_data needs %zu", weights_path, fsize, e->ck.total_bytes);
  }
  cudaError_t ce = cudaMalloc(&e->blob, e->ck.total_bytes);
  if (ce != cudaSuccess) {
    fclose(f);
    engine_free(e);
    return fail(QIE_ENOMEM, "cudaMalloc(%zu) for weights: %s", e->ck.total_bytes, cudaGetErrorString(ce));
  }
  const size_t CH = (size_t)64 << 20;
  char* pin[2] = {nullptr, nullptr};
  cudaEvent_t ev[2];
  for (int i = 0; i < 2; ++i) {
    cudaMallocHost(&pin[i], CH);
    cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming);
  }
  size_t off = 0;
  int b = 0;
  bool ok = pin[0] && pin[1];
  while (ok && off < e->ck.total_bytes) {
    size_t m = std::min(CH, e->ck.total_bytes - off);
    cudaEventSynchronize(ev[b]);
    if (fread(pin[b], 1, m, f) != m) {
      ok = false;
      break;
    }
    cudaMemcpyAsync(reinterpret_cast<char*>(e->blob) + off, pin[b], m, cudaMemcpyHostToDevice, e->stream);
    cudaEventRecord(ev[b], e->stream);
    off += m;
    b ^= 1;
  }
  cudaStreamSynchronize(e->stream);
  for (int i = 0; i < 2; ++i) {
    if (pin[i]) cudaFreeHost(pin[i]);
    cudaEventDestroy(ev[i]);
  }
  fclose(f);
  if (!ok) {
    engine_free(e);
    return fail(QIE_EIO, "read error on %s", weights_path);
  }
  rc = engine_finish_setup(e);
  if (rc) {
    engine_free(e);
    return rc;
  }
  *out = e;
  return QIE_OK;
}

int qie_engine_create_from_blob(const char* meta_path, void* device_blob, const qie_engine_opts* opts,
                                qie_engine** out) {
  if (!meta_path || !device_blob || !out) return fail(QIE_EINVAL, "engine_create_from_blob: null argument");
  qie_engine* e = nullptr;
  int rc = engine_begin(opts, &e);
  if (rc) return rc;
  std::string err;
  if (!parse_meta(meta_path, &e->ck, &err) ||
      !derive_config(e->ck, e->opts.head_dim_hint, e->opts.context, &e->cfg, &err)) {
    engine_free(e);
    return fail(QIE_EIO, "%s", err.c_str());
  }
  e->blob = (bf16*)device_blob;
  e->blob_owned = false;
  rc = engine_finish_setup(e);
  if (rc) {
    engine_free(e);
    return rc;
  }
  *out = e;
  return QIE_OK;
}

int qie_engine_create_synthetic(const qie_config* cfg, uint64_t seed, const qie_engine_opts* opts,
                                qie_engine** out) {
  if (!cfg || !out) return fail(QIE_EINVAL, "engine_create_synthetic: null argument");
  qie_engine* e = nullptr;
  int rc = engine_begin(opts, &e);
  if (rc) return rc;
  e->ck = synth_layout(*cfg);
  std::string err;
  if (!derive_config(e->ck, cfg->head_dim, e->opts.context, &e->cfg, &err)) {
    engine_free(e);
    return fail(QIE_EINVAL, "%s", err.c_str());
  }
  cudaError_t ce = cudaMalloc(&e->blob, e->ck.total_bytes);
  if (ce != cudaSuccess) {
    engine_free(e);
    return fail(QIE_ENOMEM, "cudaMalloc(%zu) for weights: %s", e->ck.total_bytes, cudaGetErrorString(ce));
  }
  for (const TensorInfo& t : e->ck.tensors) {
    ce = launch_synth_fill(e->blob, t.begin / 2, (t.end - t.begin) / 2, seed, t.kind, e->stream);
    if (ce != cudaSuccess) {
      engine_free(e);
      return cuda_fail(ce, "synth_fill");
    }
  }
  rc = engine_finish_setup(e);
  if (rc) {
    engine_free(e);
    return rc;
  }
  *out = e;
  return QIE_OK;
}

void qie_engine_destroy(qie_engine* e) { engine_free(e); }

int qie_tp_plan(const qie_config* cfg, int tp_rank, int tp_size, int* out8) {
  if (!cfg || !out8) return fail(QIE_EINVAL, "null argument");
  TpPlan p;
  if (!tp_plan(*cfg, tp_rank, tp_size, &p))
    return fail(QIE_EINVAL, "tp_size %d does not divide heads (%d q / %d kv) / intermediate %d", tp_size, cfg->n_q, cfg->n_kv, cfg->inter);
  const int v[8] = {p.n_q, p.n_kv, p.inter, p.vocab, p.q_row0, p.kv_row0, p.inter0, p.vocab0};
  memcpy(out8, v, sizeof(v));
  return QIE_OK;
}

int qie_tp_unique_id(void* out128) {
  if (!out128) return fail(QIE_EINVAL, "null argument");
  char err[256];
  if (tp_unique_id(out128, err, sizeof(err))) return fail(QIE_ECUDA, "%s", err);
  return QIE_OK;
}

int qie_engine_tp_connect(qie_engine* e, const void* id128) {
  if (!e || !id128) return fail(QIE_EINVAL, "null argument");
  if (e->opts.tp_size <= 1) return fail(QIE_ESTATE, "engine was created with tp_size 1");
  if (e->tp.comm) return fail(QIE_ESTATE, "tensor-parallel communicator already connected");
  CU(cudaSetDevice(e->opts.device));
  char err[256];
  if (tp_comm_init(&e->tp, id128, e->opts.tp_rank, e->opts.tp_size, err, sizeof(err))) return fail(QIE_ECUDA, "%s", err);
  // Peer mappings of the persistent kernel's exchange buffers: every rank exports its buffer with CUDA IPC, the
  // 64-byte handles are all-gathered over the communicator that was just created, every rank opens its peers'.
  // If any step fails the engine stays on the per-operator tensor-parallel path (NCCL all-reduce per projection).
  const char* off = getenv("QIE_TP_MEGA");
  if (e->tp_xbuf && e->mega_wmaps_tp_d[0] && !(off && off[0] == '0')) {
    const int tp = e->opts.tp_size, me = e->opts.tp_rank;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    const bool ok = cudaIpcGetMemHandle(&mine, e->tp_xbuf) == cudaSuccess;
    if (!ok) (void)cudaGetLastError();
    // the all-gather is collective: every rank takes part even if its own export failed
    struct Rec { cudaIpcMemHandle_t h; int ok; int pad[15]; };
    static_assert(sizeof(Rec) == 128, "handle record");
    std::vector<Rec> recs(tp + 1);
    memset(recs.data(), 0, recs.size() * sizeof(Rec));
    recs[0].h = mine;
    recs[0].ok = ok ? 1 : 0;
    Rec* rd = nullptr;
    CU(cudaMalloc(&rd, (size_t)(tp + 1) * sizeof(Rec)));
    CU(cudaMemcpyAsync(rd, recs.data(), sizeof(Rec), cudaMemcpyHostToDevice, e->stream));
    CU(tp_allgather(&e->tp, rd, rd + 1, sizeof(Rec), e->stream));
    CU(cudaMemcpyAsync(recs.data() + 1, rd + 1, (size_t)tp * sizeof(Rec), cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(rd);
    bool all_ok = true;
    for (int r = 0; r < tp; ++r) all_ok = all_ok && recs[1 + r].ok == 1;
    if (all_ok) {
      for (int r = 0; r < tp && all_ok; ++r) {
        if (r == me) {
          e->tp_peer_xbuf[r] = e->tp_xbuf;
        } else if (cudaIpcOpenMemHandle(&e->tp_peer_xbuf[r], recs[1 + r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          (void)cudaGetLastError();
          e->tp_peer_xbuf[r] = nullptr;
          all_ok = false;
        }
      }
    }
    // agree on the outcome: one rank on the persistent kernel and one on NCCL would dead-lock
    int* flag_d = nullptr;
    CU(cudaMalloc(&flag_d, (size_t)(tp + 1) * 128));
    int mine_ok = all_ok ? 1 : 0;
    CU(cudaMemcpyAsync(flag_d, &mine_ok, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CU(tp_allgather(&e->tp, flag_d, flag_d + 32, 128, e->stream));
    std::vector<int> flags((size_t)tp * 32);
    CU(cudaMemcpyAsync(flags.data(), flag_d + 32, (size_t)tp * 128, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    cudaFree(flag_d);
    bool everyone = true;
    for (int r = 0; r < tp; ++r) everyone = everyone && flags[(size_t)r * 32] == 1;
    e->tp_mega_ready = everyone;
  }
  return QIE_OK;
}

int qie_engine_get_config(const qie_engine* e, qie_config* out) {
  if (!e || !out) return fail(QIE_EINVAL, "null argument");
  *out = e->cfg;
  return QIE_OK;
}

const qie_bf16* qie_engine_weight(const qie_engine* e, const char* short_name, int layer, size_t* n_elems) {
  if (!e || !short_name) return nullptr;
  const TensorInfo* t = e->ck.find(short_name, layer);
  if (!t) return nullptr;
  if (n_elems) *n_elems = (t->end - t->begin) / 2;
  return reinterpret_cast<const qie_bf16*>(reinterpret_cast<const char*>(e->blob) + t->begin);
}

int qie_engine_kv_view(const qie_engine* e, qie_kv_view* out) {
  if (!e || !out) return fail(QIE_EINVAL, "null argument");
  out->pool = (qie_bf16*)e->kv.pool;
  out->n_pages = e->kv.n_pages;
  out->page_size = e->kv.page_size;
  out->n_layers = e->kv.n_layers;
  out->n_kv_heads = e->kv.n_kv;
  out->head_dim = e->kv.hd;
  return QIE_OK;
}

qie_stream qie_engine_stream(const qie_engine* e) { return e ? (qie_stream)e->stream : nullptr; }

int qie_engine_limits(const qie_engine* e, int* max_decode_rows, int* max_pages_per_seq, int* max_seqs) {
  if (!e) return fail(QIE_EINVAL, "null engine");
  if (max_decode_rows) *max_decode_rows = std::min(e->opts.max_batch_tokens, e->logits_rows);
  if (max_pages_per_seq) *max_pages_per_seq = e->max_pages_per_seq;
  if (max_seqs) *max_seqs = (int)e->seqs.size();
  return QIE_OK;
}

int qie_engine_set_sampling(qie_engine* e, int topk, float temperature_prefill, float temperature_decode,
                            uint64_t seed, int add_step) {
  if (!e || topk < 1 || topk > 256) return fail(QIE_EINVAL, "set_sampling: topk must be in [1,256]");
  cudaStreamSynchronize(e->stream);
  for (auto& kvp : e->graphs)  // sampling parameters are baked into captured graphs
    if (kvp.second.exec) cudaGraphExecDestroy(kvp.second.exec);
  e->graphs.clear();
  e->topk = topk;
  e->temp_prefill = temperature_prefill;
  e->temp_decode = temperature_decode;
  e->seed = seed;
  e->add_step = add_step;
  return QIE_OK;
}

int qie_engine_set_repetition_penalty(qie_engine* e, float penalty) {
  if (!e || !(penalty > 0.0f)) return fail(QIE_EINVAL, "repetition penalty must be > 0");
  if (e->opts.tp_size > 1 && penalty != 1.0f) return fail(QIE_EINVAL, "repetition penalty over a sharded vocabulary is not built");
  CU(cudaSetDevice(e->opts.device));
  CU(cudaStreamSynchronize(e->stream));
  if (penalty != 1.0f && !e->hist_d) {
    const size_t n = e->seqs.size() * (size_t)e->cfg.context;
    CU(cudaMalloc(&e->hist_d, n * sizeof(int)));
    CU(cudaMemset(e->hist_d, 0xff, n * sizeof(int)));  // -1: positions never written are ignored by the kernel
  }
  e->rep_penalty = penalty;
  for (auto& kvp : e->graphs)  // captured graphs bake the launch sequence in
    if (kvp.second.exec) cudaGraphExecDestroy(kvp.second.exec);
  e->graphs.clear();
  return QIE_OK;
}

int qie_seq_new(qie_engine* e, int* seq) {
  if (!e || !seq) return fail(QIE_EINVAL, "null argument");
  for (int i = 0; i < (int)e->seqs.size(); ++i)
    if (!e->seqs[i].live) {
      e->seqs[i] = Sequence();
      e->seqs[i].live = true;
      *seq = i;
      return QIE_OK;
    }
  return fail(QIE_ENOMEM, "all %zu sequence slots are live", e->seqs.size());
}

static int check_seq(const qie_engine* e, int seq) {
  if (!e || seq < 0 || seq >= (int)e->seqs.size() || !e->seqs[seq].live) return fail(QIE_EINVAL, "bad sequence id %d", seq);
  return QIE_OK;
}

int qie_seq_free(qie_engine* e, int seq) {
  int rc = check_seq(e, seq);
  if (rc) return rc;
  cudaStreamSynchronize(e->stream);
  for (int p : e->seqs[seq].pages) e->free_pages.push_back(p);
  if (e->seqs[seq].host_copy) cudaFreeHost(e->seqs[seq].host_copy);
  e->seqs[seq] = Sequence();
  return QIE_OK;
}

int qie_seq_len(const qie_engine* e, int seq) {
  int rc = check_seq(e, seq);
  if (rc) return rc;
  return e->seqs[seq].len;
}

int qie_kv_pages_free(const qie_engine* e) { return e ? (int)e->free_pages.size() : QIE_EINVAL; }

// grow the page list of `seq` to hold new_len positions; returns 1 if the block table
// row changed, 0 if not, negative on error.
static int ensure_pages(qie_engine* e, int seq, int new_len) {
  Sequence& s = e->seqs[seq];
  if (new_len > e->cfg.context) return fail(QIE_EINVAL, "sequence %d would exceed the context (%d)", seq, e->cfg.context);
  int need = (new_len + e->kv.page_size - 1) / e->kv.page_size;
  if (need > e->max_pages_per_seq) return fail(QIE_ENOMEM, "sequence %d needs %d pages > max %d", seq, need, e->max_pages_per_seq);
  int changed = 0;
  while ((int)s.pages.size() < need) {
    if (e->free_pages.empty()) return fail(QIE_ENOMEM, "KV pool exhausted (%d pages)", e->kv.n_pages);
    int p = e->free_pages.back();
    e->free_pages.pop_back();
    e->block_table_h[(size_t)seq * e->max_pages_per_seq + s.pages.size()] = p;
    s.pages.push_back(p);
    changed = 1;
  }
  return changed;
}

static cudaError_t push_block_row(qie_engine* e, int seq) {
  size_t off = (size_t)seq * e->max_pages_per_seq;
  return cudaMemcpyAsync(e->block_table_d + off, e->block_table_h + off, e->max_pages_per_seq * sizeof(int),
                         cudaMemcpyHostToDevice, e->stream);
}

int qie_prefill(qie_engine* e, int seq, const int32_t* h_ids, int n, int32_t* h_token) {
  int rc = check_seq(e, seq);
  if (rc) return rc;
  if (!h_ids || n <= 0 || !h_token) return fail(QIE_EINVAL, "prefill: empty prompt or null buffers");
  for (int i = 0; i < n; ++i)
    if (h_ids[i] < 0 || h_ids[i] >= e->cfg.vocab) return fail(QIE_EINVAL, "prefill: token id %d out of range", h_ids[i]);
  CU(cudaSetDevice(e->opts.device));
  Sequence& s = e->seqs[seq];
  if (s.host_copy) return fail(QIE_ESTATE, "sequence %d is swapped out (qie_seq_swap_in first)", seq);
  rc = ensure_pages(e, seq, s.len + n);
  if (rc < 0) return rc;
  if (rc) CU(push_block_row(e, seq));
  const int R = e->opts.max_batch_tokens;
  e->launches = 0;
  for (int c0 = 0; c0 < n; c0 += R) {
    int m = std::min(R, n - c0);
    bool last = c0 + m == n;
    int* st = e->stage_h;
    CU(cudaStreamSynchronize(e->stream));  // staging buffer reuse
    for (int i = 0; i < m; ++i) {
      st[i] = h_ids[c0 + i];
      st[R + i] = s.len + c0 + i;
      st[2 * R + i] = seq;
      st[3 * R + i] = s.step;
    }
    CU(cudaMemcpyAsync(e->ids_d, st, m * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->pos_d, st + R, m * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CU(cudaMemcpyAsync(e->slot_d, st + 2 * R, m * sizeof(int), cudaMemcpyHostToDevice, e->stream));
    // the sampled row is row 0 of the sampler -> its step lives in rowstep_d[0]
    CU(cudaMemcpyAsync(e->rowstep_d, st + 3 * R, sizeof(int), cudaMemcpyHostToDevice, e->stream));
    CU(forward_rows(e, m, s.len + c0 + m, m - 1, last ? 1 : 0, e->temp_prefill, false));
  }
  CU(cudaMemcpyAsync(e->sampled_h, e->sampled_d, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  *h_token = e->sampled_h[0];
  s.len += n;
  s.step += 1;  // iengine.cu:419
  return QIE_OK;
}

// one decode step for the batch already staged on the device (ids/pos/slot/rowstep)
static cudaError_t decode_forward(qie_engine* e, int n, int bucket) {
  if (decode_uses_mega(e, n, bucket)) {
    cudaError_t r = forward_decode_mega(e, n, bucket, e->temp_decode);
    if (r != cudaErrorNotReady) return r;  // NotReady: the cooperative launch could not be placed, nothing ran yet
  }
  return forward_rows(e, n, bucket, 0, n, e->temp_decode, true);
}

static int decode_launch(qie_engine* e, int n, int max_kv_len) {
  const int bucket = ((max_kv_len + 127) / 128) * 128;
  if (!e->opts.use_graph || e->capture) {
    CU(decode_forward(e, n, bucket));
    return QIE_OK;
  }
  auto key = std::make_pair(n, bucket);
  auto it = e->graphs.find(key);
  if (it == e->graphs.end()) {
    e->graphs.emplace(key, qie_engine::GraphEntry());
    CU(decode_forward(e, n, bucket));
    return QIE_OK;
  }
  if (!it->second.exec) {
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeThreadLocal));
    long before = e->launches;
    cudaError_t fe = decode_forward(e, n, bucket);
    cudaError_t ce = cudaStreamEndCapture(e->stream, &graph);
    if (fe != cudaSuccess) return cuda_fail(fe, "forward (capture)");
    if (ce != cudaSuccess) return cuda_fail(ce, "cudaStreamEndCapture");
    it->second.launches = e->launches - before;
    e->launches = before;
    CU(cudaGraphInstantiate(&it->second.exec, graph, 0));
    cudaGraphDestroy(graph);
  }
  CU(cudaGraphLaunch(it->second.exec, e->stream));
  e->launches += it->second.launches;  // the graph replays exactly these kernels
  return QIE_OK;
}

static int decode_prepare(qie_engine* e, const int* h_seqs, int n, int* max_kv_len) {
  if (!e || !h_seqs || n <= 0) return fail(QIE_EINVAL, "decode: empty batch");
  if (n > e->opts.max_batch_tokens || n > e->logits_rows) return fail(QIE_EINVAL, "decode: batch %d exceeds max %d", n, e->logits_rows);
  CU(cudaSetDevice(e->opts.device));
  int mk = 0;
  for (int i = 0; i < n; ++i) {
    int rc = check_seq(e, h_seqs[i]);
    if (rc) return rc;
    Sequence& s = e->seqs[h_seqs[i]];
    if (s.len == 0) return fail(QIE_ESTATE, "sequence %d has not been prefilled", h_seqs[i]);
    if (s.host_copy) return fail(QIE_ESTATE, "sequence %d is swapped out (qie_seq_swap_in first)", h_seqs[i]);
    rc = ensure_pages(e, h_seqs[i], s.len + 1);
    if (rc < 0) return rc;
    if (rc) CU(push_block_row(e, h_seqs[i]));
    mk = std::max(mk, s.len + 1);
    for (int j = 0; j < i; ++j)  // two rows of one sequence would write the same KV slot
      if (h_seqs[j] == h_seqs[i]) return fail(QIE_EINVAL, "decode: sequence %d appears twice in the batch", h_seqs[i]);
  }
  *max_kv_len = mk;
  return QIE_OK;
}

static int decode_stage_inputs(qie_engine* e, const int* h_seqs, const int32_t* h_tokens_in, int n) {
  const int R = e->opts.max_batch_tokens;
  int* st = e->stage_h;
  CU(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < n; ++i) {
    if (h_tokens_in[i] < 0 || h_tokens_in[i] >= e->cfg.vocab) return fail(QIE_EINVAL, "decode: token id %d out of range", h_tokens_in[i]);
    const Sequence& s = e->seqs[h_seqs[i]];
    st[i] = h_tokens_in[i];
    st[R + i] = s.len;  // position of the new token (sequence_len-1 after ++, qwen_main.cu:265,300)
    st[2 * R + i] = h_seqs[i];
    st[3 * R + i] = s.step;
  }
  CU(cudaMemcpyAsync(e->ids_d, st, n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->pos_d, st + R, n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->slot_d, st + 2 * R, n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  CU(cudaMemcpyAsync(e->rowstep_d, st + 3 * R, n * sizeof(int), cudaMemcpyHostToDevice, e->stream));
  return QIE_OK;
}

int qie_decode_step(qie_engine* e, const int* h_seqs, const int32_t* h_tokens_in, int n, int32_t* h_tokens_out) {
  if (!h_tokens_in || !h_tokens_out) return fail(QIE_EINVAL, "decode: null buffers");
  int mk = 0;
  int rc = decode_prepare(e, h_seqs, n, &mk);
  if (rc) return rc;
  rc = decode_stage_inputs(e, h_seqs, h_tokens_in, n);
  if (rc) return rc;
  e->launches = 0;
  rc = decode_launch(e, n, mk);
  if (rc) return rc;
  CU(cudaMemcpyAsync(e->sampled_h, e->sampled_d, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < n; ++i) {
    h_tokens_out[i] = e->sampled_h[i];
    e->seqs[h_seqs[i]].len += 1;
    e->seqs[h_seqs[i]].step += 1;
  }
  return QIE_OK;
}

int qie_decode_run(qie_engine* e, const int* h_seqs, const int32_t* h_tokens_in, int n, int steps,
                   int32_t* h_tokens_out) {
  if (!h_tokens_in || !h_tokens_out || steps <= 0) return fail(QIE_EINVAL, "decode_run: bad arguments");
  if (!e) return fail(QIE_EINVAL, "null engine");
  if ((size_t)n * steps > e->sampled_h_cap) {
    CU(cudaStreamSynchronize(e->stream));
    if (e->sampled_h) cudaFreeHost(e->sampled_h);
    e->sampled_h = nullptr;
    e->sampled_h_cap = (size_t)n * steps;
    CU(cudaMallocHost(&e->sampled_h, e->sampled_h_cap * sizeof(int)));
  }
  e->launches = 0;
  for (int s = 0; s < steps; ++s) {
    int mk = 0;
    int rc = decode_prepare(e, h_seqs, n, &mk);
    if (rc) return rc;
    if (s == 0) {
      rc = decode_stage_inputs(e, h_seqs, h_tokens_in, n);
      if (rc) return rc;
    }
    rc = decode_launch(e, n, mk);
    if (rc) return rc;
    CU(cudaMemcpyAsync(e->sampled_h + (size_t)s * n, e->sampled_d, n * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    for (int i = 0; i < n; ++i) {
      e->seqs[h_seqs[i]].len += 1;
      e->seqs[h_seqs[i]].step += 1;
    }
  }
  CU(cudaStreamSynchronize(e->stream));
  memcpy(h_tokens_out, e->sampled_h, (size_t)n * steps * sizeof(int));
  return QIE_OK;
}

int qie_decode_step_device(qie_engine* e, const int* h_seqs, int n) {
  int mk = 0;
  int rc = decode_prepare(e, h_seqs, n, &mk);
  if (rc) return rc;
  e->launches = 0;
  rc = decode_launch(e, n, mk);
  if (rc) return rc;
  for (int i = 0; i < n; ++i) {
    e->seqs[h_seqs[i]].len += 1;
    e->seqs[h_seqs[i]].step += 1;
  }
  return QIE_OK;
}

int qie_sync(qie_engine* e) {
  if (!e) return fail(QIE_EINVAL, "null engine");
  CU(cudaStreamSynchronize(e->stream));
  return QIE_OK;
}

int qie_capture_enable(qie_engine* e, int on) {
  if (!e) return fail(QIE_EINVAL, "null engine");
  e->capture = on != 0;
  return QIE_OK;
}

long qie_capture_read(qie_engine* e, const char* tag, int layer, qie_bf16* h_out, size_t max_elems) {
  if (!e || !tag || !h_out) return fail(QIE_EINVAL, "null argument");
  std::string key = std::string(tag) + "#" + std::to_string(layer);
  auto it = e->cap.find(key);
  if (it == e->cap.end() || !it->second.d) return fail(QIE_EINVAL, "no capture for %s", key.c_str());
  size_t n = std::min(max_elems, it->second.elems);
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaMemcpy(h_out, it->second.d, n * sizeof(bf16), cudaMemcpyDeviceToHost));
  return (long)it->second.elems;
}

long qie_launch_count(const qie_engine* e) { return e ? e->launches : 0; }

// KV offload (the experiment commented out in the reference's main(), iengine.cu:376-429): the pages of a
// sequence are copied to pinned host memory and handed back to the pool; swap-in takes fresh pages (any ids) and
// copies the bytes back.  A page holds all layers, so one copy per page moves page_stride() elements.
int qie_seq_swap_out(qie_engine* e, int seq) {
  int rc = check_seq(e, seq);
  if (rc) return rc;
  Sequence& s = e->seqs[seq];
  if (s.host_copy) return fail(QIE_ESTATE, "sequence %d is already swapped out", seq);
  if (s.pages.empty()) return QIE_OK;
  CU(cudaSetDevice(e->opts.device));
  const size_t pb = e->kv.page_stride() * sizeof(bf16);
  CU(cudaMallocHost(&s.host_copy, s.pages.size() * pb));
  for (size_t i = 0; i < s.pages.size(); ++i)
    CU(cudaMemcpyAsync(static_cast<char*>(s.host_copy) + i * pb, reinterpret_cast<const char*>(e->kv.pool) + (size_t)s.pages[i] * pb, pb,
                       cudaMemcpyDeviceToHost, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  s.host_pages = (int)s.pages.size();
  for (int p : s.pages) e->free_pages.push_back(p);
  s.pages.clear();
  return QIE_OK;
}

int qie_seq_swap_in(qie_engine* e, int seq) {
  int rc = check_seq(e, seq);
  if (rc) return rc;
  Sequence& s = e->seqs[seq];
  if (!s.host_copy) return fail(QIE_ESTATE, "sequence %d is not swapped out", seq);
  if ((int)e->free_pages.size() < s.host_pages) return fail(QIE_ENOMEM, "swap_in: %d pages needed, %zu free", s.host_pages, e->free_pages.size());
  CU(cudaSetDevice(e->opts.device));
  const size_t pb = e->kv.page_stride() * sizeof(bf16);
  for (int i = 0; i < s.host_pages; ++i) {
    const int p = e->free_pages.back();
    e->free_pages.pop_back();
    e->block_table_h[(size_t)seq * e->max_pages_per_seq + i] = p;
    s.pages.push_back(p);
    CU(cudaMemcpyAsync(reinterpret_cast<char*>(e->kv.pool) + (size_t)p * pb, static_cast<const char*>(s.host_copy) + (size_t)i * pb, pb,
                       cudaMemcpyHostToDevice, e->stream));
  }
  CU(push_block_row(e, seq));
  CU(cudaStreamSynchronize(e->stream));
  cudaFreeHost(s.host_copy);
  s.host_copy = nullptr;
  s.host_pages = 0;
  return QIE_OK;
}

// Cache rows of one sequence <-> HOST memory in the REFERENCE's page layout [position][layer][kv_dim]
// (iengine.cu:352, include_cuda.cu:165-279: element offset ((pos % page_size) * L + layer) * Dkv inside a page), so a
// reference page of page_size positions starting at pos0 is exactly the buffer of a read with n = page_size.
// Pages travel whole (page_stride elements, one copy each); the re-layout happens on the host.
static int seq_kv_copy(qie_engine* e, int seq, int pos0, int n, qie_bf16* h_K, qie_bf16* h_V, bool write) {
  int rc = check_seq(e, seq);
  if (rc) return rc;
  if (n <= 0 || pos0 < 0 || !h_K || !h_V) return fail(QIE_EINVAL, "seq_kv: bad range or null buffers");
  Sequence& s = e->seqs[seq];
  if (s.host_copy) return fail(QIE_ESTATE, "sequence %d is swapped out", seq);
  CU(cudaSetDevice(e->opts.device));
  if (write) {
    rc = ensure_pages(e, seq, pos0 + n);
    if (rc < 0) return rc;
    if (rc) CU(push_block_row(e, seq));
  } else if (pos0 + n > s.len) {
    return fail(QIE_EINVAL, "seq_kv_read: rows [%d, %d) beyond the sequence length %d", pos0, pos0 + n, s.len);
  }
  const KvGeom& g = e->kv;
  const int ps = g.page_size, L = g.n_layers, nkv = g.n_kv, hd = g.hd, Dkv = nkv * hd;
  const size_t pe = g.page_stride();
  std::vector<uint16_t> page(pe);
  CU(cudaStreamSynchronize(e->stream));
  for (int pi = pos0 / ps; pi <= (pos0 + n - 1) / ps; ++pi) {
    uint16_t* dev = reinterpret_cast<uint16_t*>(g.pool) + (size_t)s.pages[pi] * pe;
    const int lo = std::max(pos0, pi * ps), hi = std::min(pos0 + n, (pi + 1) * ps);
    if (!write || lo != pi * ps || hi != (pi + 1) * ps) CU(cudaMemcpy(page.data(), dev, pe * 2, cudaMemcpyDeviceToHost));
    for (int pos = lo; pos < hi; ++pos)
      for (int l = 0; l < L; ++l)
        for (int kv = 0; kv < 2; ++kv)
          for (int h = 0; h < nkv; ++h) {
            uint16_t* pool_row = page.data() + (size_t)l * g.layer_stride() + (size_t)kv * g.kv_stride() + (size_t)h * g.head_stride() +
                                 (size_t)(pos - pi * ps) * hd;
            uint16_t* host_row = (kv ? h_V : h_K) + ((size_t)(pos - pos0) * L + l) * Dkv + (size_t)h * hd;
            if (write) memcpy(pool_row, host_row, hd * 2);
            else memcpy(host_row, pool_row, hd * 2);
          }
    if (write) CU(cudaMemcpy(dev, page.data(), pe * 2, cudaMemcpyHostToDevice));
  }
  if (write && pos0 + n > s.len) {
    s.len = pos0 + n;
    if (s.step == 0) s.step = 1;
  }
  return QIE_OK;
}

int qie_seq_kv_read(qie_engine* e, int seq, int pos0, int n, qie_bf16* h_K, qie_bf16* h_V) {
  return seq_kv_copy(e, seq, pos0, n, h_K, h_V, false);
}

int qie_seq_kv_write(qie_engine* e, int seq, int pos0, int n, const qie_bf16* h_K, const qie_bf16* h_V) {
  return seq_kv_copy(e, seq, pos0, n, const_cast<qie_bf16*>(h_K), const_cast<qie_bf16*>(h_V), true);
}

int qie_seq_fill_synthetic(qie_engine* e, int seq, int n_pos, uint64_t seed) {
  int rc = check_seq(e, seq);
  if (rc) return rc;
  if (n_pos <= 0) return fail(QIE_EINVAL, "fill_synthetic: n_pos must be > 0");
  CU(cudaSetDevice(e->opts.device));
  Sequence& s = e->seqs[seq];
  rc = ensure_pages(e, seq, s.len + n_pos);
  if (rc < 0) return rc;
  if (rc) CU(push_block_row(e, seq));
  CU(launch_kv_fill(e->kv, e->block_table_d + (size_t)seq * e->max_pages_per_seq, s.len, n_pos,
                    seed + 0x51ull * (uint64_t)seq, e->stream));
  CU(cudaStreamSynchronize(e->stream));
  s.len += n_pos;
  if (s.step == 0) s.step = 1;
  return QIE_OK;
}

int qie_engine_set_int(qie_engine* e, const char* key, long value) {
  if (!e || !key) return fail(QIE_EINVAL, "null argument");
  std::string k(key);
  if (k == "mega") e->use_mega = value != 0;
  else if (k == "gemv") e->use_gemv = value != 0;
  else if (k == "gemv_dataflow") e->gemv_dataflow = value;
  else if (k == "mega_layers_run") e->mega_layers_run = (int)value;
  else if (k == "mega_prof") e->mega_prof_on = value != 0;
  else if (k == "layer_first") e->layer_first = (int)std::max(0l, value);
  else if (k == "layer_count") e->layer_count = (int)std::max(0l, value);
  else if (k == "inject_x") e->inject_x = value != 0;
  else return fail(QIE_EINVAL, "unknown option %s", key);
  // captured graphs bake these choices in
  CU(cudaStreamSynchronize(e->stream));
  for (auto& kvp : e->graphs)
    if (kvp.second.exec) cudaGraphExecDestroy(kvp.second.exec);
  e->graphs.clear();
  return QIE_OK;
}

int qie_decode_uses_mega(const qie_engine* e, int n_rows, int kv_len) {
  if (!e) return 0;
  return decode_uses_mega(e, n_rows, ((kv_len + 127) / 128) * 128) ? 1 : 0;
}

long qie_mega_prof_read(qie_engine* e, uint64_t* h_out, size_t max_values) {
  if (!e || !h_out) return fail(QIE_EINVAL, "null argument");
  if (!e->mega_prof_d) return fail(QIE_ESTATE, "persistent decode kernel not available for this checkpoint");
  // layout: [16L+8] globaltimer ns, then [16L+8] SM clock counter values of the same points
  size_t n = std::min<size_t>(max_values, (size_t)2 * (16 * e->cfg.layers + 8) + 16);
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaMemcpy(h_out, e->mega_prof_d, n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return (long)n;
}

long qie_engine_read_activation(qie_engine* e, const char* name, void* h_out, size_t max_bytes) {
  if (!e || !name || !h_out) return fail(QIE_EINVAL, "null argument");
  const qie_config& c = e->cfg;
  const size_t R = e->opts.max_batch_tokens;
  const size_t Dq = (size_t)c.n_q * c.head_dim, Dkv = (size_t)c.n_kv * c.head_dim;
  std::string k(name);
  const void* src = nullptr;
  size_t bytes = 0;
  if (k == "x") src = e->x, bytes = R * c.hidden * 2;
  else if (k == "qkv") src = e->qkv, bytes = R * (Dq + 2 * Dkv) * 2;
  else if (k == "att") src = e->att, bytes = R * Dq * 2;
  else if (k == "h") src = e->h, bytes = R * c.inter * 2;
  else if (k == "logits") src = e->logits, bytes = (size_t)e->logits_rows * c.vocab * 2;
  else if (k == "sampled") src = e->sampled_d, bytes = R * sizeof(int);
  else if (k == "kv") src = e->kv.pool, bytes = (size_t)e->kv.n_pages * e->kv.page_stride() * 2;
  else return fail(QIE_EINVAL, "unknown activation %s", name);
  bytes = std::min(bytes, max_bytes);
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaMemcpy(h_out, src, bytes, cudaMemcpyDeviceToHost));
  return (long)bytes;
}

long qie_engine_write_activation(qie_engine* e, const char* name, const void* h_in, size_t bytes) {
  if (!e || !name || !h_in) return fail(QIE_EINVAL, "null argument");
  if (std::string(name) != "x") return fail(QIE_EINVAL, "only the residual stream x can be written");
  const size_t cap = (size_t)e->opts.max_batch_tokens * e->cfg.hidden * 2;
  if (bytes > cap) return fail(QIE_EINVAL, "write_activation: %zu bytes > %zu", bytes, cap);
  CU(cudaSetDevice(e->opts.device));
  CU(cudaStreamSynchronize(e->stream));
  CU(cudaMemcpy(e->x, h_in, bytes, cudaMemcpyHostToDevice));
  return (long)bytes;
}

const char* qie_kernel_kind_name(int kind) {
  static const char* names[KK_COUNT] = {"embedding", "rmsnorm", "gemm_qkv", "qkv_post", "attention", "gemm_o",
                                        "gemm_gateup", "gemm_down", "lm_head", "sample", "advance"};
  return kind >= 0 && kind < KK_COUNT ? names[kind] : nullptr;
}

int qie_decode_step_profile(qie_engine* e, const int* h_seqs, const int32_t* h_tokens_in, int n, float* ms_by_kind,
                            int* launches_by_kind, int n_kinds) {
  if (!h_tokens_in || !ms_by_kind || !launches_by_kind || n_kinds < KK_COUNT)
    return fail(QIE_EINVAL, "decode_step_profile: need %d kinds", (int)KK_COUNT);
  int mk = 0;
  int rc = decode_prepare(e, h_seqs, n, &mk);
  if (rc) return rc;
  rc = decode_stage_inputs(e, h_seqs, h_tokens_in, n);
  if (rc) return rc;
  e->launches = 0;
  e->prof.clear();
  e->prof_on = true;
  cudaError_t fe = forward_rows(e, n, ((mk + 127) / 128) * 128, 0, n, e->temp_decode, true);
  e->prof_on = false;
  cudaError_t se = cudaStreamSynchronize(e->stream);
  for (int i = 0; i < n_kinds; ++i) {
    ms_by_kind[i] = 0.f;
    launches_by_kind[i] = 0;
  }
  for (auto& r : e->prof) {
    float ms = 0.f;
    if (fe == cudaSuccess && se == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      ms_by_kind[r.kind] += ms;
      launches_by_kind[r.kind] += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  e->prof.clear();
  if (fe != cudaSuccess) return cuda_fail(fe, "forward (profile)");
  if (se != cudaSuccess) return cuda_fail(se, "sync (profile)");
  for (int i = 0; i < n; ++i) {
    e->seqs[h_seqs[i]].len += 1;
    e->seqs[h_seqs[i]].step += 1;
  }
  return QIE_OK;
}

}  // extern "C"
