// include/layers/iengine_compat.hh -- the reference's host-side types and entry points
// (/root/reference/layers/include/{iengine.cuh,utils.hh,tensor_parser.hh,helpers.cuh,
// layers_include.cuh}) re-declared on top of libqie_b200, so that iengine's main()
// (src/iengine.cu:226-482) and llm()'s call sites (src/qwen_main.cu) compile against the
// B200 library.  Same names, argument order and meaning; what changes is behind them:
//
//   tensor / TensorTable / build_indexed_tensors / parsed_tensors   tensor_parser.hh:35-50
//       read model_files/meta_data.txt (the file the reference WRITES) instead of
//       re-parsing safetensors; set QIE_META to point somewhere else.
//   page_table / create_page_list / allocate_page_buffers / free_page_list   iengine.cuh:39-55
//       nodes are host bookkeeping; KV memory lives in the engine's page pool
//       (pool[page][layer][k|v][head][slot][hd]) -- k_page_ptr/v_page_ptr point into it.
//   batch_metadata / ModelBuffers / initialize_model_buffers / destroy_model_buffers
//       iengine.cuh:23-37, utils.hh:14-105.  ModelBuffers keeps the fields main() reads.
//   llm()                                                           iengine.cuh:51
//       prefill or one decode token of one sequence; returns the sampled token id, or a
//       NEGATIVE qie error code (never token 0) on failure.
//   launch_* / proj / sample_topk_bf16 / copy_*                     helpers.cuh:18-166
//       operator-level wrappers over the C ABI (default stream, asynchronous).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

#include <fstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "../qie_b200.h"

#define TILE_SIZE 16
#define WARP_SIZE 32
#define CONTEXT_SIZE 32786UL

struct tensor {
  std::string tensor_name;
  std::vector<size_t> shape;
  std::vector<size_t> data_offsets;
  int layer_index = -1;
  std::string short_name;
};
using TensorTable = std::unordered_map<std::string, std::vector<tensor>>;

std::ostream& operator<<(std::ostream& os, const tensor& t);
std::vector<tensor> parsed_tensors();
TensorTable build_indexed_tensors();
void precompute_cos_sin(float* cos_values, float* sin_values, int seq_len, int head_dim);

typedef enum { prefill, decode } State;

struct ModelBuffers {
  int* d_token_ids = nullptr;
  size_t sequence_len = 0;
  size_t number_of_layers = 0, head_dim = 0, hidden_dim = 0, hidden_dim_kv = 0, num_of_qheads = 0, num_of_kvheads = 0,
         context_size = 0, vocab_size = 0, up_dim = 0;
  __nv_bfloat16* embeddings_d = nullptr;  // points into the weight blob
  __nv_bfloat16* k_cache = nullptr;
  __nv_bfloat16* v_cache = nullptr;
  std::vector<int> h_token_ids;  // prompt kept on the host for llm(prefill)
  int qie_seq = -1;              // sequence slot inside the engine
};

typedef struct {
  int sequence_id;
  __nv_bfloat16* k_ptr;
  __nv_bfloat16* v_ptr;
  State state;
  int sequence_len;
  int generated_token;
  int step;
  ModelBuffers* buffer;
} batch_metadata;

typedef struct page_table_struct {
  __nv_bfloat16* k_page_ptr;
  __nv_bfloat16* v_page_ptr;
  int page_allocated = 0;
  struct page_table_struct* ptr_to_next_page;
} page_table;

page_table* create_page_list(int pages_required);
void allocate_page_buffers(page_table* node, size_t elems_per_page);
void free_page_list(page_table* head);

void initialize_model_buffers(ModelBuffers& buf, int* h_token_ids, TensorTable& tensors, std::ifstream& weights,
                              size_t sequence_len);
void destroy_model_buffers(ModelBuffers& buf);

// Bind the compat layer to a weight blob already resident on the device (what main() gets
// from load_all_weights_to_gpu_chunked, iengine.cu:117-223). Called lazily by llm().
qie_engine* qie_compat_engine(__nv_bfloat16* g_gpu_weights_buffer);

int llm(batch_metadata* new_seq, TensorTable tensors, std::ifstream& weights, page_table* kv_cache_seq1, int page_size,
        __nv_bfloat16* g_gpu_weights_buffer);

// ---- helpers.cuh:18-166 ---------------------------------------------------------------
template <class T>
void assign_weight_pointer(const tensor& t, T*& d, __nv_bfloat16* g_gpu_weights_buffer) {
  d = reinterpret_cast<T*>(reinterpret_cast<char*>(g_gpu_weights_buffer) + t.data_offsets[0]);
}
template <class T>
void load_weight(const tensor& t, std::ifstream&, T*, T*& d, size_t, __nv_bfloat16* g_gpu_weights_buffer) {
  assign_weight_pointer(t, d, g_gpu_weights_buffer);
}
void launch_rms(__nv_bfloat16* x, __nv_bfloat16* w, __nv_bfloat16* y, size_t hidden, size_t seqlen);
void launch_rope(float* cos_d, float* sin_d, __nv_bfloat16* x, size_t seqlen, size_t head_dim, size_t hidden_dim,
                 size_t nheads);
void launch_rope_single(float* cos_d, float* sin_d, __nv_bfloat16* x, size_t pos, size_t head_dim, int hidden_dim,
                        int nheads);
void launch_matmul(__nv_bfloat16* A, __nv_bfloat16* B, __nv_bfloat16* C, int M, int N, int K);
void launch_elem(__nv_bfloat16* a, __nv_bfloat16* b, __nv_bfloat16* out, int n);
void launch_act(__nv_bfloat16* x, size_t n);
void launch_resadd(__nv_bfloat16* x, __nv_bfloat16* y, size_t n);
void launch_qknorm(__nv_bfloat16* X, __nv_bfloat16* w, int head_dim, int seqlen, int hidden, int nheads);
void proj(const tensor& t, std::ifstream& f, __nv_bfloat16* w_h, __nv_bfloat16* w_d, size_t w_elems, __nv_bfloat16* x,
          __nv_bfloat16* y, int m, int n, int k, __nv_bfloat16* g_gpu_weights_buffer);
int sample_topk_bf16(__nv_bfloat16* logits_d, int vocab, float temperature, int topk, unsigned long long seed, int step);
static inline void copy_last_vocab_vec(__nv_bfloat16* seq, __nv_bfloat16* dst, int hidden, int seqlen) {
  cudaMemcpyAsync(dst, &seq[(size_t)(seqlen - 1) * hidden], hidden * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice, 0);
}
static inline void copy_first_token(__nv_bfloat16* seq, __nv_bfloat16* dst, int hidden) {
  cudaMemcpyAsync(dst, seq, hidden * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice, 0);
}
