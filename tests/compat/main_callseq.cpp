// tests/compat/main_callseq.cpp -- TEST PROGRAM.  The call sequence of the reference's host driver, main()
// (/root/reference/layers/src/iengine.cu:226-482), written against include/layers/iengine_compat.hh and linked with
// libqie_b200.so: same names, same argument order, same order of calls --
//   build_indexed_tensors (:245) -> weights.bin into ONE device blob (:117-223) -> create_new_sequence +
//   initialize_model_buffers (:25-47) -> create_page_list + allocate_page_buffers per page (:334-351) ->
//   llm(prefill) -> { generated_token / state = decode / step++ ; llm(decode) } (:384-456)
// -- so that "iengine's main() compiles and runs against the B200 library" is something a test executes
// (tests/test_gpu_compat_replay.py), not a claim in INTEGRATION.md.  What the reference hard-codes (absolute paths,
// the prompt literal, getchar() between tokens) comes from argv / the environment here.
//   usage: QIE_META=<meta_data.txt> main_callseq <weights.bin> <n_new_tokens> <id> <id> ...
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#include "layers/iengine_compat.hh"

static batch_metadata* create_new_sequence(int* h_token_ids, size_t n, TensorTable& tensors, std::ifstream& weights) {  // iengine.cu:25-47
  batch_metadata* s = (batch_metadata*)malloc(sizeof(batch_metadata));
  s->sequence_id = 0;
  s->state = prefill;
  s->sequence_len = (int)n;
  s->generated_token = 0;
  s->step = 0;
  s->buffer = new ModelBuffers();
  initialize_model_buffers(*s->buffer, h_token_ids, tensors, weights, n);
  return s;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: QIE_META=meta_data.txt %s weights.bin n_new id...\n", argv[0]);
    return 2;
  }
  const char* weights_path = argv[1];
  const int n_new = atoi(argv[2]);
  std::vector<int> ids;
  for (int i = 3; i < argc; ++i) ids.push_back(atoi(argv[i]));

  TensorTable tensors = build_indexed_tensors();  // iengine.cu:245
  if (tensors.empty()) return 3;
  size_t max_end = 0;
  for (auto& kv : tensors)
    for (auto& t : kv.second)
      if (t.data_offsets.size() == 2 && t.data_offsets[1] > max_end) max_end = t.data_offsets[1];
  std::ifstream weights(weights_path, std::ios::binary);  // iengine.cu:232
  if (!weights) return 4;
  __nv_bfloat16* g_gpu_weights_buffer = nullptr;  // load_all_weights_to_gpu_chunked, iengine.cu:117-223
  if (cudaMalloc(&g_gpu_weights_buffer, max_end) != cudaSuccess) return 5;
  {
    std::vector<char> chunk(64u << 20);
    size_t done = 0;
    while (done < max_end) {
      size_t n = std::min(chunk.size(), max_end - done);
      weights.read(chunk.data(), (std::streamsize)n);
      if ((size_t)weights.gcount() != n) return 6;
      cudaMemcpy(reinterpret_cast<char*>(g_gpu_weights_buffer) + done, chunk.data(), n, cudaMemcpyHostToDevice);
      done += n;
    }
  }
  batch_metadata* new_seq = create_new_sequence(ids.data(), ids.size(), tensors, weights);
  ModelBuffers* b = new_seq->buffer;
  const int page_size = 4;  // iengine.cu:334
  const int pages_required = (int)((ids.size() + page_size - 1) / page_size) + 1;
  page_table* kv_cache_seq1 = create_page_list(pages_required);
  const size_t elems = (size_t)page_size * b->number_of_layers * b->hidden_dim_kv;  // iengine.cu:352
  for (page_table* p = kv_cache_seq1; p; p = p->ptr_to_next_page) allocate_page_buffers(p, elems);
  b->k_cache = kv_cache_seq1->k_page_ptr;  // iengine.cu:359-360
  b->v_cache = kv_cache_seq1->v_page_ptr;

  for (int i = 0; i < n_new; ++i) {  // iengine.cu:384-456 (the getchar() pause and the prints are the reference's UI)
    int out = llm(new_seq, tensors, weights, kv_cache_seq1, page_size, g_gpu_weights_buffer);
    if (out < 0) {
      fprintf(stderr, "llm failed: %d (%s)\n", out, qie_last_error());
      return 7;
    }
    printf("%d\n", out);
    new_seq->generated_token = out;
    new_seq->state = decode;
    new_seq->step += 1;  // iengine.cu:419
  }
  destroy_model_buffers(*b);
  free_page_list(kv_cache_seq1);
  delete b;
  free(new_seq);
  cudaFree(g_gpu_weights_buffer);
  return 0;
}
