// engine.h -- host side of the forward: weights, KV page pool, sequences, batch
// scheduler state, forward launch order.  Replaces llm() (src/qwen_main.cu:64-417),
// ModelBuffers (include/utils.hh:14-88), batch_metadata / page_table
// (include/iengine.cuh:23-48) and the page-list allocator (src/iengine.cu:72-109) of the
// reference.  One engine per GPU; no global mutable state, so data-parallel replicas are
// independent.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "checkpoint.h"
#include "kernels.h"
#include "tp_nccl.h"

namespace qie {

struct LayerWeights {
  const bf16 *in_ln = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *o = nullptr, *q_norm = nullptr,
             *k_norm = nullptr, *post_ln = nullptr, *up = nullptr, *gate = nullptr, *down = nullptr;
  const bf16 *q_bias = nullptr, *k_bias = nullptr, *v_bias = nullptr;  // Qwen2.5 checkpoints (the reference has none)
};

struct Sequence {
  bool live = false;
  int len = 0;   // positions held in the KV cache (ModelBuffers::sequence_len)
  int step = 0;  // batch_metadata::step
  std::vector<int> pages;
  // swapped out (qie_seq_swap_out): the pages' bytes live in pinned host memory, `pages` is empty
  void* host_copy = nullptr;
  int host_pages = 0;
};

struct CaptureBuf {
  bf16* d = nullptr;
  size_t elems = 0, cap = 0;
};

}  // namespace qie

struct qie_engine {
  qie_config cfg{};
  qie_engine_opts opts{};
  int num_sms = 0;
  cudaStream_t stream = nullptr;

  // weights: ONE device blob, tensors addressed as blob + begin (helpers.cuh:18-29)
  qie::Checkpoint ck;
  qie::bf16* blob = nullptr;
  bool blob_owned = true;
  std::vector<qie::LayerWeights> L;
  const qie::bf16 *embed = nullptr, *final_norm = nullptr, *lm_head = nullptr;
  float *cos_d = nullptr, *sin_d = nullptr;

  // KV pool + page allocator
  qie::KvGeom kv{};
  std::vector<int> free_pages;
  std::vector<qie::Sequence> seqs;
  int max_pages_per_seq = 0;
  int* block_table_d = nullptr;
  int* block_table_h = nullptr;  // pinned mirror

  // per-forward row metadata (device) + pinned staging
  int *ids_d = nullptr, *pos_d = nullptr, *slot_d = nullptr, *sampled_d = nullptr, *rowstep_d = nullptr;
  int *stage_h = nullptr;    // pinned: ids | pos | slot | rowstep
  int* sampled_h = nullptr;  // pinned
  size_t sampled_h_cap = 0;

  // activations [max_batch_tokens, *]
  qie::bf16 *x = nullptr, *xn = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *att = nullptr, *h = nullptr,
            *logits = nullptr;
  int logits_rows = 0;

  // fast numerics (opts.numerics == QIE_NUMERICS_FAST)
  struct LayerMaps {
    qie::TensorMap2D q, k, v, o, gate, up, down;
  };
  std::vector<LayerMaps> wmaps;  // TMA descriptors of the weight matrices (box 128 x 64)
  qie::TensorMap2D lm_head_map;
  qie::bf16* qkv = nullptr;      // fused [rows, Dq + 2*Dkv] projection output
  int* gemm_counters = nullptr;  // fused split-K reduction tickets (self-resetting)
  float* gemm_ws = nullptr;      // split-K partials
  size_t gemm_ws_bytes = 0;
  float *attn_ws_o = nullptr, *attn_ws_ml = nullptr;
  int attn_max_splits = 16;

  // persistent decode kernel (decode_mega.cu): reference-order decode steps of <= 8 rows
  bool use_mega = true;              // QIE_MEGA=0 disables
  qie::MegaLayer* mega_layers_d = nullptr;
  // 7 per layer + lm_head 3-D weight views (box 8 rows x mega_kc[s]); set 1 (larger tiles) serves batches <= 8
  qie::TensorMap2D* mega_wmaps_d[2] = {nullptr, nullptr};
  int mega_kc[2] = {0, 0};
  qie::MegaCand* mega_cand_d = nullptr;
  unsigned* mega_bar_d = nullptr;
  qie::TensorMap2D* mega_hmap_d = nullptr;   // h [rows, I] in boxes of 16 rows x KC (batches > 8): A stream of the tile-split down_proj
  qie::TensorMap2D* mega_kvmap_d = nullptr;  // KV pool rows for the TMA K/V stream of the persistent kernel's attention
  unsigned long long* mega_prof_d = nullptr;  // phase timestamps of the last profiled step
  bool mega_prof_on = false;
  int mega_layers_run = 0;           // debug: run only this many layers (no lm_head)
  // GEMV decode kernel (decode_gemv.cu): fast-numerics steps of <= DECODE_GEMV_MAX_ROWS rows
  int gemv_dataflow = -1;            // set_int("gemv_dataflow", 0/1): grid barriers / polled per-layer buffers in decode_gemv.cu (-1: default)
  bool use_gemv = true;              // QIE_GEMV=0 / set_int("gemv", 0): the split-K variant of decode_mega.cu instead
  float* gemv_part_d = nullptr;      // GEMV decode kernel: per-layer activation buffers + split-KV partials (0xFF = not stored yet)
  // parity hooks of the per-operator forward (tests/test_gpu_layer_isolation.py): run layers [layer_first,
  // layer_first + layer_count) only (0 = all, then final norm + lm_head) and/or take the residual stream x as
  // written by qie_engine_write_activation instead of the embedding rows
  int layer_first = 0, layer_count = 0;
  bool inject_x = false;

  // tensor parallel (opts.tp_size > 1): heads / intermediate / vocabulary sharded, NCCL all-reduce
  qie::TpComm tp;
  qie::TpPlan plan{};
  float* tp_buf = nullptr;            // [max_batch_tokens, hidden] fp32 partial sums of o_proj / down_proj
  qie::TpCand* tp_cand = nullptr;     // [1 + tp_size][max rows] local + gathered arg-max candidates
  // exchange buffer of the persistent kernel: [MEGA_TP_HEADER: flags, generation, candidates][2][tp][MEGA_TP_ROWS][hidden] fp32, exported
  // to the peers through CUDA IPC (handles travel over the NCCL communicator in qie_engine_tp_connect)
  void* tp_xbuf = nullptr;
  void* tp_peer_xbuf[qie::MEGA_MAX_TP] = {};  // [rank] mapped base (own entry = tp_xbuf)
  bool tp_mega_ready = false;
  qie::bf16* x2 = nullptr;            // residual ping-pong partner of x
  qie::TensorMap2D* mega_wmaps_tp_d[2] = {nullptr, nullptr};  // weight views of this rank's shard

  // model semantics (opts.semantics / opts.rms_eps; include/qie_b200.h)
  float eps = 1e-04f;
  bool rope_half = false, has_bias = false;

  // sampling
  float rep_penalty = 1.0f;  // != 1: token history kept per sequence, penalty applied to the logits before sampling
  int* hist_d = nullptr;     // [max_seqs][cfg.context] input tokens by position (allocated when a penalty is first set)
  int topk = 1;
  float temp_prefill = 1.0f, temp_decode = 0.7f;
  uint64_t seed = 1234;
  int add_step = 1;

  // capture (parity hooks)
  bool capture = false;
  std::map<std::string, qie::CaptureBuf> cap;

  // CUDA graphs of the decode step keyed by (rows, kv bucket). A shape is run eagerly
  // the first time it is seen (that also sets kernel attributes) and captured the second.
  struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    long launches = 0;
  };
  std::map<std::pair<int, int>, GraphEntry> graphs;

  long launches = 0;

  // per-kernel-class device timing of one eager step (bench.py roofline leg)
  bool prof_on = false;
  struct ProfRec {
    int kind;
    cudaEvent_t a, b;
  };
  std::vector<ProfRec> prof;
};

namespace qie {
enum KernelKind {
  KK_EMBED = 0, KK_RMSNORM, KK_GEMM_QKV, KK_QKV_POST, KK_ATTN, KK_GEMM_O, KK_GEMM_GATEUP, KK_GEMM_DOWN,
  KK_LM_HEAD, KK_SAMPLE, KK_ADVANCE, KK_COUNT
};
// one forward over n_rows rows already described in ids_d/pos_d/slot_d; logits for rows
// [out_row0, out_row0+n_out); samples into sampled_d[0..n_out). Returns cudaSuccess or the
// first launch error. Counts launches into e->launches.
cudaError_t forward_rows(qie_engine* e, int n_rows, int max_kv_len, int out_row0, int n_out, float temperature,
                         bool advance);
// tensor-parallel forward (per-operator launches, reference-order kernels on the local shard)
cudaError_t forward_rows_tp(qie_engine* e, int n_rows, int max_kv_len, int out_row0, int n_out, float temperature,
                            bool advance);
// true if a decode step of n rows / this kv bucket runs through the persistent kernel
bool decode_uses_mega(const qie_engine* e, int n_rows, int max_kv_len);
// one decode step (rows described in ids_d/pos_d/slot_d) through the persistent kernel
cudaError_t forward_decode_mega(qie_engine* e, int n_rows, int max_kv_len, float temperature);
}  // namespace qie
