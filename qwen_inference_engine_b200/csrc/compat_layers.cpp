// compat_layers.cpp -- implementation of include/layers/iengine_compat.hh: the reference's
// host API names on top of the C ABI.  Host code only; every computation goes through
// libqie_b200's kernels.
#include "../../include/layers/iengine_compat.hh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <mutex>

#include "checkpoint.h"

static const char* meta_path() {
  const char* p = getenv("QIE_META");
  return p ? p : "../model_files/meta_data.txt";  // tensor_parser.cpp:34
}

std::ostream& operator<<(std::ostream& os, const tensor& t) {  // tensor_parser.cpp:19-28
  os << "Tensor: " << t.tensor_name << "\n";
  os << "  layer: " << t.layer_index << "\n";
  os << "  short_name: " << t.short_name << "\n";
  os << "  shape: [ ";
  for (auto s : t.shape) os << s << " ";
  os << "]\n";
  os << "  offsets: [ " << t.data_offsets[0] << ", " << t.data_offsets[1] << " ]\n";
  return os;
}

std::vector<tensor> parsed_tensors() {
  qie::Checkpoint ck;
  std::string err;
  std::vector<tensor> out;
  if (!qie::parse_meta(meta_path(), &ck, &err)) {
    std::cerr << "parsed_tensors: " << err << "\n";
    return out;
  }
  for (const qie::TensorInfo& ti : ck.tensors) {
    tensor t;
    t.tensor_name = ti.name;
    t.short_name = ti.short_name;
    t.layer_index = ti.layer;
    t.shape = ti.shape;
    t.data_offsets = {ti.begin, ti.end};
    out.push_back(t);
  }
  return out;
}

TensorTable build_indexed_tensors() {  // tensor_parser.cpp:132-165
  TensorTable indexed;
  for (auto& t : parsed_tensors()) {
    size_t slot = t.layer_index >= 0 ? (size_t)t.layer_index : 0;
    auto& v = indexed[t.short_name];
    if (v.size() <= slot) v.resize(slot + 1);
    v[slot] = t;
  }
  return indexed;
}

void precompute_cos_sin(float* c, float* s, int seq_len, int head_dim) { qie_precompute_cos_sin(c, s, seq_len, head_dim); }

// ---- page list (iengine.cu:72-109): host nodes, one device K / V buffer pair per page -----------------
static std::mutex g_pg_mu;
static std::map<page_table*, size_t> g_page_elems;  // node -> elems_per_page it was allocated with

page_table* create_page_list(int pages_required) {
  page_table* head = nullptr;
  page_table** cur = &head;
  for (int i = 0; i < pages_required; ++i) {
    *cur = new page_table();
    (*cur)->k_page_ptr = nullptr;
    (*cur)->v_page_ptr = nullptr;
    (*cur)->page_allocated = 0;
    (*cur)->ptr_to_next_page = nullptr;
    cur = &((*cur)->ptr_to_next_page);
  }
  return head;
}
void allocate_page_buffers(page_table* node, size_t elems_per_page) {
  if (!node || node->k_page_ptr) return;
  if (cudaMalloc(&node->k_page_ptr, elems_per_page * sizeof(__nv_bfloat16)) != cudaSuccess ||
      cudaMalloc(&node->v_page_ptr, elems_per_page * sizeof(__nv_bfloat16)) != cudaSuccess) {
    fprintf(stderr, "allocate_page_buffers: out of device memory\n");
    abort();  // iengine.cu:51-70: the reference's allocator macros abort too
  }
  node->page_allocated = 1;
  std::lock_guard<std::mutex> lk(g_pg_mu);
  g_page_elems[node] = elems_per_page;
}
void free_page_list(page_table* head) {
  cudaDeviceSynchronize();
  std::lock_guard<std::mutex> lk(g_pg_mu);
  while (head) {
    page_table* next = head->ptr_to_next_page;
    if (head->k_page_ptr) cudaFree(head->k_page_ptr);
    if (head->v_page_ptr) cudaFree(head->v_page_ptr);
    g_page_elems.erase(head);
    delete head;
    head = next;
  }
}

// device arrays of the list's K / V page pointers, in list order (what the reference's kernel finds by walking
// ptr_to_next_page); rebuilt per call: the list may have grown (kv_copy_layer_to_cache_decode allocates on demand)
struct PageArrays {
  qie_bf16** d_k = nullptr;
  qie_bf16** d_v = nullptr;
  int n = 0;
  size_t elems = 0;
};
static int gather_pages(page_table* head, PageArrays* out) {
  std::vector<qie_bf16*> k, v;
  size_t elems = 0;
  {
    std::lock_guard<std::mutex> lk(g_pg_mu);
    for (page_table* p = head; p && p->k_page_ptr; p = p->ptr_to_next_page) {
      k.push_back(reinterpret_cast<qie_bf16*>(p->k_page_ptr));
      v.push_back(reinterpret_cast<qie_bf16*>(p->v_page_ptr));
      auto it = g_page_elems.find(p);
      if (it != g_page_elems.end()) elems = it->second;
    }
  }
  if (k.empty()) return QIE_EINVAL;
  static thread_local qie_bf16** d_buf = nullptr;
  static thread_local size_t d_cap = 0;
  if (d_cap < 2 * k.size()) {
    if (d_buf) cudaFree(d_buf);
    d_cap = 2 * k.size() + 64;
    if (cudaMalloc(&d_buf, d_cap * sizeof(void*)) != cudaSuccess) return QIE_ENOMEM;
  }
  cudaMemcpy(d_buf, k.data(), k.size() * sizeof(void*), cudaMemcpyHostToDevice);
  cudaMemcpy(d_buf + k.size(), v.data(), v.size() * sizeof(void*), cudaMemcpyHostToDevice);
  out->d_k = d_buf;
  out->d_v = d_buf + k.size();
  out->n = (int)k.size();
  out->elems = elems;
  return QIE_OK;
}

void kv_copy_layer_to_cache_prefill(ModelBuffers* b, int i, page_table* head, int page_size) {
  PageArrays pa;
  if (!b || gather_pages(head, &pa) != QIE_OK) return;
  if (qie_kv_store_pagelist(pa.d_k, pa.d_v, pa.n, page_size, (int)b->number_of_layers, i, (int)b->hidden_dim_kv,
                            reinterpret_cast<qie_bf16*>(b->K), reinterpret_cast<qie_bf16*>(b->V), 0, (int)b->sequence_len, nullptr) != QIE_OK)
    fprintf(stderr, "kv_copy_layer_to_cache_prefill: %s\n", qie_last_error());
  cudaStreamSynchronize(nullptr);  // the page-pointer scratch is reused by the next call
}
void kv_copy_layer_to_cache_decode(ModelBuffers* b, int i, page_table* head, int page_size) {
  if (!b || !head) return;
  const size_t pos = b->sequence_len - 1;
  const int page_idx = (int)(pos / page_size);
  // include_cuda.cu:248-261: a missing page is appended and allocated on the spot
  page_table* p = head;
  for (int n = 0; n < page_idx; ++n) {
    if (!p->ptr_to_next_page) {
      p->ptr_to_next_page = create_page_list(1);
      allocate_page_buffers(p->ptr_to_next_page, (size_t)page_size * b->number_of_layers * b->hidden_dim_kv);
    }
    p = p->ptr_to_next_page;
  }
  PageArrays pa;
  if (gather_pages(head, &pa) != QIE_OK) return;
  if (qie_kv_store_pagelist(pa.d_k, pa.d_v, pa.n, page_size, (int)b->number_of_layers, i, (int)b->hidden_dim_kv,
                            reinterpret_cast<qie_bf16*>(b->K), reinterpret_cast<qie_bf16*>(b->V), (int)pos, 1, nullptr) != QIE_OK)
    fprintf(stderr, "kv_copy_layer_to_cache_decode: %s\n", qie_last_error());
  cudaStreamSynchronize(nullptr);
}

void launch_attn(__nv_bfloat16* Q, __nv_bfloat16* out, size_t mq, size_t mkv, size_t head_dim, size_t hidden, size_t hidden_kv,
                 int causal, size_t q_abs_base, int layer_id, page_table* head, int page_size) {
  PageArrays pa;
  if (gather_pages(head, &pa) != QIE_OK || !pa.elems || mq == 0 || mkv == 0) return;
  const int n_layers = (int)(pa.elems / ((size_t)page_size * hidden_kv));  // the reference hard-codes 40 (self_attension.cu:35)
  // row t sees positions 0 .. min(mkv - 1, q_abs_base + t) when causal, else 0 .. mkv - 1 (self_attension.cu:84-89)
  std::vector<int> pos(mq);
  for (size_t t = 0; t < mq; ++t) pos[t] = (int)(causal ? std::min(mkv - 1, q_abs_base + t) : mkv - 1);
  static thread_local int* d_pos = nullptr;
  static thread_local size_t d_pos_cap = 0;
  if (d_pos_cap < mq) {
    if (d_pos) cudaFree(d_pos);
    d_pos_cap = mq + 256;
    if (cudaMalloc(&d_pos, d_pos_cap * sizeof(int)) != cudaSuccess) return;
  }
  cudaMemcpy(d_pos, pos.data(), mq * sizeof(int), cudaMemcpyHostToDevice);
  if (qie_attention_pagelist(pa.d_k, pa.d_v, pa.n, page_size, n_layers, layer_id, (int)(hidden_kv / head_dim), (int)head_dim,
                             reinterpret_cast<qie_bf16*>(Q), reinterpret_cast<qie_bf16*>(out), d_pos, (int)mq, (int)(hidden / head_dim),
                             nullptr) != QIE_OK)
    fprintf(stderr, "launch_attn: %s\n", qie_last_error());
  cudaStreamSynchronize(nullptr);  // helpers.cuh:128: the reference wrapper synchronises (d_pos / page arrays are reused)
}

void launch_embedding(__nv_bfloat16* out, __nv_bfloat16* table, int* d_token_ids, size_t hidden, size_t n_tok) {
  qie_embedding(reinterpret_cast<qie_bf16*>(out), reinterpret_cast<qie_bf16*>(table), d_token_ids, hidden, n_tok, nullptr);
}

// ---- engine binding ---------------------------------------------------------------------
static std::mutex g_mu;
static std::map<void*, qie_engine*> g_engines;

extern "C" int qie_engine_create_from_blob(const char* meta_path, void* device_blob, const qie_engine_opts* opts,
                                           qie_engine** out);

qie_engine* qie_compat_engine(__nv_bfloat16* blob) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_engines.find(blob);
  if (it != g_engines.end()) return it->second;
  qie_engine_opts o;
  qie_engine_opts_default(&o);
  o.page_size = 16;
  qie_engine* e = nullptr;
  if (qie_engine_create_from_blob(meta_path(), blob, &o, &e) != QIE_OK) {
    fprintf(stderr, "qie compat: %s\n", qie_last_error());
    return nullptr;
  }
  // the reference hard-codes top-k 50, T = 1.0 (prefill) / 0.7 (decode), seed 1234 + step (qwen_main.cu:241,381-388);
  // QIE_COMPAT_TOPK overrides k (1 = the greedy mode BASELINE configs[0] asks for)
  int topk = 50;
  if (const char* v = getenv("QIE_COMPAT_TOPK")) topk = std::max(1, atoi(v));
  qie_engine_set_sampling(e, topk, 1.0f, 0.7f, 1234, 1);
  g_engines[blob] = e;
  return e;
}

// driver-level state that the reference keeps implicitly (the prompt lives in d_token_ids, the sequence in its page
// list): the prompt ids on the host and the engine sequence slot, keyed by the ModelBuffers the caller owns
struct CompatSeq {
  std::vector<int> ids;
  int seq = -1;
};
static std::map<ModelBuffers*, CompatSeq> g_seqs;

void initialize_model_buffers(ModelBuffers& buf, int* h_token_ids, TensorTable& tensors, std::ifstream&, size_t n) {
  memset(&buf, 0, sizeof(buf));
  buf.sequence_len = n;
  auto dim = [&](const char* name, int i) -> size_t {
    auto it = tensors.find(name);
    return it == tensors.end() || it->second.empty() || (int)it->second[0].shape.size() <= i ? 0 : it->second[0].shape[i];
  };
  buf.hidden_dim = dim("embed_tokens.weight", 1);
  buf.vocab_size = dim("logits", 0);
  buf.up_dim = dim("mlp.up_proj.weight", 0);
  buf.head_dim = dim("self_attn.q_norm.weight", 0);
  buf.hidden_dim_kv = dim("self_attn.k_proj.weight", 0);
  buf.num_of_qheads = buf.head_dim ? dim("self_attn.q_proj.weight", 0) / buf.head_dim : 0;
  buf.num_of_kvheads = buf.head_dim ? buf.hidden_dim_kv / buf.head_dim : 0;
  buf.number_of_layers = tensors.count("input_layernorm.weight") ? tensors["input_layernorm.weight"].size() : 0;
  buf.context_size = CONTEXT_SIZE;
  std::lock_guard<std::mutex> lk(g_mu);
  g_seqs[&buf].ids.assign(h_token_ids, h_token_ids + n);
}
void destroy_model_buffers(ModelBuffers& buf) {
  std::lock_guard<std::mutex> lk(g_mu);
  g_seqs.erase(&buf);
}

int llm(batch_metadata* seq, TensorTable, std::ifstream&, page_table*, int, __nv_bfloat16* blob) {
  qie_engine* e = qie_compat_engine(blob);
  if (!e || !seq || !seq->buffer) return QIE_EINVAL;
  ModelBuffers* b = seq->buffer;
  CompatSeq* cs;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    cs = &g_seqs[b];
  }
  if (cs->seq < 0 && qie_seq_new(e, &cs->seq) != QIE_OK) return QIE_ENOMEM;
  int32_t tok = 0;
  int rc;
  if (seq->state == prefill) {
    rc = qie_prefill(e, cs->seq, cs->ids.data(), (int)cs->ids.size(), &tok);
  } else {
    int32_t in = seq->generated_token;
    rc = qie_decode_step(e, &cs->seq, &in, 1, &tok);
    b->sequence_len += 1;  // qwen_main.cu:265
  }
  if (rc != QIE_OK) {
    fprintf(stderr, "llm: %s\n", qie_last_error());
    return rc;  // negative: never a token id
  }
  return tok;
}

// ---- operator wrappers (default stream) --------------------------------------------------
#define QB(p) reinterpret_cast<qie_bf16*>(p)
void launch_rms(__nv_bfloat16* x, __nv_bfloat16* w, __nv_bfloat16* y, size_t hidden, size_t seqlen) {
  qie_rmsnorm(QB(x), QB(w), QB(y), hidden, seqlen, nullptr);
}
void launch_rope(float* c, float* s, __nv_bfloat16* x, size_t seqlen, size_t hd, size_t hidden, size_t nheads) {
  qie_rope(c, s, QB(x), (int)seqlen, 0, (int)hd, (int)hidden, (int)nheads, nullptr);
}
void launch_rope_single(float* c, float* s, __nv_bfloat16* x, size_t pos, size_t hd, int hidden, int nheads) {
  qie_rope(c, s, QB(x), 1, (int)pos, (int)hd, hidden, nheads, nullptr);
}
void launch_matmul(__nv_bfloat16* A, __nv_bfloat16* B, __nv_bfloat16* C, int M, int N, int K) {
  qie_matmul(QB(A), QB(B), QB(C), M, N, K, nullptr);
}
void launch_elem(__nv_bfloat16* a, __nv_bfloat16* b, __nv_bfloat16* out, int n) { qie_elem_mul(QB(a), QB(b), QB(out), n, nullptr); }
void launch_act(__nv_bfloat16* x, size_t n) { qie_silu(QB(x), n, nullptr); }
void launch_resadd(__nv_bfloat16* x, __nv_bfloat16* y, size_t n) { qie_residual_add(QB(x), QB(y), n, nullptr); }
void launch_qknorm(__nv_bfloat16* X, __nv_bfloat16* w, int hd, int seqlen, int hidden, int nheads) {
  qie_qknorm(QB(X), QB(w), hd, seqlen, hidden, nheads, nullptr);
}
void proj(const tensor& t, std::ifstream&, __nv_bfloat16*, __nv_bfloat16* w_d, size_t, __nv_bfloat16* x, __nv_bfloat16* y,
          int m, int n, int k, __nv_bfloat16* blob) {
  assign_weight_pointer(t, w_d, blob);
  launch_matmul(x, w_d, y, m, n, k);
}
void apply_repetition_penalty(__nv_bfloat16* logits, const int* context_tokens, size_t context_len, int vocab_size, float penalty) {
  qie_repetition_penalty(QB(logits), context_tokens, context_len, vocab_size, penalty, nullptr);
}
int sample_topk_bf16(__nv_bfloat16* logits_d, int vocab, float temperature, int topk, unsigned long long seed, int step) {
  int* d_tok = nullptr;
  int h_tok = -1;
  if (cudaMalloc(&d_tok, sizeof(int)) != cudaSuccess) return QIE_ECUDA;
  // `step` is the cuRAND subsequence in the reference (helpers.cuh:157-166 -> logit_decode.cu:256-257), not a seed offset
  int rc = qie_sample_topk_subseq(QB(logits_d), d_tok, 1, (size_t)vocab, temperature, topk, seed, 0, (unsigned long long)step, nullptr);
  cudaMemcpy(&h_tok, d_tok, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_tok);
  return rc == QIE_OK ? h_tok : rc;
}
