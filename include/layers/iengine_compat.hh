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
//       host nodes with a device K and V buffer each, in the reference's [slot][layer][kv_dim] layout: the
//       operator-level launch_attn / kv_copy_layer_to_cache_* work on them directly.  The driver-level llm()
//       keeps its cache in the engine's page pool instead (pool[page][layer][k|v][head][slot][hd]).
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

// utils.hh:14-88, field for field (a plain struct of sizes and raw pointers: the contract callers read and write).
// initialize_model_buffers fills the sizes; the operator-level call sites of llm() (qwen_main.cu:83-236) use the
// activation buffers, which the caller allocates as the reference does (utills.cu:4-129).  The *_h / *_weights_*
// staging pointers of the reference stay null (weights are addressed in the blob, helpers.cuh:18-29).
struct ModelBuffers {
  int* d_token_ids;
  size_t sequence_len;
  size_t number_of_layers, head_dim, hidden_dim, hidden_dim_kv, num_of_qheads, num_of_kvheads, context_size, vocab_size, up_dim;
  __nv_bfloat16 *embeddings_h, *embeddings_d, *embeddings_out;
  float *cos_values_h, *sin_values_h, *cos_values_d, *sin_values_d;
  __nv_bfloat16 *k_cache, *v_cache;
  __nv_bfloat16 *norm_weights_h, *norm_weights_d, *rms_out;
  __nv_bfloat16 *qk_norm_weights_h, *qk_norm_weights_d;
  __nv_bfloat16 *q_proj_weights_h, *q_proj_weights_d, *Q;
  size_t q_proj_size;
  __nv_bfloat16 *kv_proj_weights_h, *kv_proj_weights_d;
  size_t kv_proj_size;
  __nv_bfloat16 *K, *V;
  __nv_bfloat16* atten_out;
  __nv_bfloat16 *o_proj_weights_h, *o_proj_weights_d, *O, *out_proj;
  size_t o_proj_size;
  __nv_bfloat16 *mlp_up_proj_weights_h, *mlp_up_proj_weights_d, *MLP_UP, *MLP_GATE, *MLP_GATE_OUT, *MLP_DOWN;
  size_t mlp_up_proj_size;
  __nv_bfloat16* test_out;
  __nv_bfloat16* last_x;
  __nv_bfloat16* prefill_output_d;
  __nv_bfloat16 *logits_weights_h, *logits_weights_d;
  size_t logtis_shape;
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
// cudaMalloc pair of elems_per_page bf16 per node (iengine.cu:89-96); elems_per_page = page_size * layers * kv_dim is
// remembered per node so that launch_attn can recover the layer count the reference hard-codes (self_attension.cu:35)
void allocate_page_buffers(page_table* node, size_t elems_per_page);
void free_page_list(page_table* head);  // frees the device buffers and the nodes (the reference std::free()s managed memory)
// include_cuda.cu:165-279: rows of buffer->K / buffer->V of layer i -> the pages, element ((pos % page_size) * layers +
// i) * kv_dim; prefill writes positions 0..sequence_len-1, decode position sequence_len-1 (allocating a page on demand)
void kv_copy_layer_to_cache_prefill(ModelBuffers* buffer, int i, page_table* kv_cache_seq1, int page_size);
void kv_copy_layer_to_cache_decode(ModelBuffers* buffer, int i, page_table* kv_cache_seq1, int page_size);

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
// the two kernels llm() / initialize_model_buffers launch directly with <<<>>> (embedding_matrix_func: utills.cu:51-54,
// qwen_main.cu:259-267; the sampler: qwen_main.cu:388) cannot cross a shared-library boundary as __global__ symbols;
// a maintainer replaces those three launches with launch_embedding / sample_topk_bf16
void launch_embedding(__nv_bfloat16* out, __nv_bfloat16* table, int* d_token_ids, size_t hidden, size_t n_tok);
// helpers.cuh:121-129.  Row t of Q attends to cache positions 0 .. min(mkv - 1, causal ? q_abs_base + t : mkv - 1) of the
// page list (reference layout, see allocate_page_buffers)
void launch_attn(__nv_bfloat16* Q, __nv_bfloat16* out, size_t mq, size_t mkv, size_t head_dim, size_t hidden, size_t hidden_kv,
                 int causal, size_t q_abs_base, int layer_id, page_table* kv_cache_seq1, int page_size);
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
// layers_include.cuh:33 declares the kernel and nothing defines it; host wrapper with the declared argument list
void apply_repetition_penalty(__nv_bfloat16* logits, const int* context_tokens, size_t context_len, int vocab_size, float penalty);
static inline void copy_last_vocab_vec(__nv_bfloat16* seq, __nv_bfloat16* dst, int hidden, int seqlen) {
  cudaMemcpyAsync(dst, &seq[(size_t)(seqlen - 1) * hidden], hidden * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice, 0);
}
static inline void copy_first_token(__nv_bfloat16* seq, __nv_bfloat16* dst, int hidden) {
  cudaMemcpyAsync(dst, seq, hidden * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice, 0);
}
