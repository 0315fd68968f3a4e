// kernels.h -- host-visible launch interface of the sm_100a kernels (internal header;
// the public boundary is include/qie_b200.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace qie {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------- KV pool geometry
// pool[page][layer][k|v][kv_head][slot][head_dim]
struct KvGeom {
  bf16* pool;
  int n_pages, page_size, n_layers, n_kv, hd;
  __host__ __device__ size_t head_stride() const { return (size_t)page_size * hd; }
  __host__ __device__ size_t kv_stride() const { return (size_t)n_kv * page_size * hd; }
  __host__ __device__ size_t layer_stride() const { return 2 * kv_stride(); }
  __host__ __device__ size_t page_stride() const { return (size_t)n_layers * layer_stride(); }
  __host__ __device__ bf16* chunk(int page, int layer, int kv, int head) const {
    return pool + (size_t)page * page_stride() + (size_t)layer * layer_stride() + (size_t)kv * kv_stride() +
           (size_t)head * head_stride();
  }
};

// ---------------------------------------------------------------- reference-order GEMM
enum { EPI_STORE = 0, EPI_RESIDUAL = 1, EPI_SILU_MUL = 2, EPI_STORE_F32 = 3 };  // F32: `out` is really float* (tensor-parallel partial sums)

struct GemmSeg {
  const bf16* W;   // [rows, K] row-major (HF [out,in])
  const bf16* W2;  // second weight (up_proj) for EPI_SILU_MUL, else unused
  bf16* out;       // [M, ld_out]
  int rows;
  int ld_out;
  int ldw;  // weight row stride in elements; 0 = K (a column slice of a wider matrix has ldw > K)
};

struct GemmArgs {
  const bf16* A;  // [M, lda]
  int lda, M, K;
  int nseg, total_units, epi, dual, stages;
  int unit_begin[4];
  GemmSeg seg[3];
};

// Launch C = A * W^T for up to 3 weight segments sharing A (e.g. q,k,v). M <= 64.
cudaError_t launch_gemm_ref_order(const GemmArgs& args, int num_sms, cudaStream_t st);

// ---------------------------------------------------------------- row-wise ops
cudaError_t launch_embedding(bf16* out, const bf16* table, const int* ids, size_t hidden, size_t n_tok,
                             cudaStream_t st);
cudaError_t launch_rmsnorm_ref(const bf16* x, const bf16* w, bf16* y, size_t hidden, size_t n_tok,
                               size_t x_stride, cudaStream_t st, float eps = 1e-04f);
cudaError_t launch_qknorm_ref(bf16* x, const bf16* w, int hd, int n_tok, int row_dim, int n_heads, cudaStream_t st);
cudaError_t launch_rope_ref(const float* cos_t, const float* sin_t, bf16* x, int n_tok, const int* pos, int pos0,
                            int hd, int row_dim, int n_heads, cudaStream_t st);
cudaError_t launch_silu(bf16* x, size_t n, cudaStream_t st);
cudaError_t launch_elem_mul(const bf16* a, const bf16* b, bf16* c, size_t n, cudaStream_t st);
cudaError_t launch_residual_add(bf16* a, const bf16* b, size_t n, cudaStream_t st);

// fused q/k-norm + RoPE + KV store for one layer:
//   q (in place) and k: per-head RMS norm (if the norm weights are non-null), RoPE at
//   pos[t]; k (post-RoPE) and v (raw) are written to the KV pool.
struct QkvPostArgs {
  bf16* q;          // [n_tok, n_q*hd]  output (contiguous)
  const bf16* q_in; // projection output rows (may alias q), row stride q_in_stride
  const bf16* k;    // row stride kv_stride
  const bf16* v;    // row stride kv_stride
  int q_in_stride, kv_stride;
  const bf16* q_norm_w;
  const bf16* k_norm_w;
  const float* cos_t;
  const float* sin_t;
  const int* pos;          // [n_tok]
  const int* slot;         // [n_tok] block-table row
  const int* block_table;  // [*, max_pages]
  int max_pages, n_tok, n_q, layer;
  KvGeom kv;
  // model semantics beyond the reference's (data / option driven): projection biases (Qwen2.5),
  // RMS epsilon (1e-4 in the reference, normalization.cu:9 / qk_norm.cu:46; 1e-6 in HF), half-rotation RoPE
  const bf16 *q_bias = nullptr, *k_bias = nullptr, *v_bias = nullptr;
  float eps = 1e-04f;
  int rope_half = 0;
};
cudaError_t launch_qkv_post(const QkvPostArgs& a, cudaStream_t st);

cudaError_t launch_kv_store(const KvGeom& kv, int layer, const bf16* K, const bf16* V, const int* pos,
                            const int* slot, const int* block_table, int max_pages, int n_tok, cudaStream_t st);

// reference-order attention: row t attends to positions 0..pos[t] of its sequence.
struct AttnArgs {
  const bf16* q;  // [n_tok, n_q*hd]
  bf16* out;      // [n_tok, n_q*hd]
  const int* pos;
  const int* slot;
  const int* block_table;
  int max_pages, n_tok, n_q, layer;
  int max_kv_len;  // upper bound on pos[t]+1 (sizes shared memory)
  KvGeom kv;
  // page-list mode (the reference's own cache layout, include/iengine.cuh:42-48 + include_cuda.cu:165-279): when
  // k_pages != nullptr, position p of the ONE sequence all rows belong to lives in page p / kv.page_size at element
  // ((p % kv.page_size) * pl_layers + layer) * n_kv*hd of k_pages[..] / v_pages[..]; slot / block_table are unused and
  // kv carries only page_size, n_kv, hd
  const bf16* const* k_pages = nullptr;
  const bf16* const* v_pages = nullptr;
  int pl_layers = 0;
};
cudaError_t launch_attention_ref(const AttnArgs& a, cudaStream_t st);
// K, V rows [n_tok, kv_dim] -> positions pos0 + t of a page list in the reference's layout (see AttnArgs)
cudaError_t launch_kv_store_pagelist(bf16* const* k_pages, bf16* const* v_pages, int page_size, int n_layers, int layer, int kv_dim,
                                     const bf16* K, const bf16* V, int pos0, int n_tok, cudaStream_t st);

// split-KV flash-decoding (fast numerics): one query token per row, row b = sequence slot[b]
struct FastAttnArgs {
  const bf16* q;  // [n_tok, n_q*hd]
  bf16* out;      // [n_tok, n_q*hd]
  const int* pos;
  const int* slot;
  const int* block_table;
  int max_pages, n_tok, n_q, layer, n_splits;
  float scale_log2;  // log2(e) / sqrt(hd)
  float* ws_o;       // [n_splits][n_tok][n_q][hd]
  float* ws_ml;      // [n_splits][n_tok][n_q][2]
  KvGeom kv;
};
cudaError_t launch_attention_decode_fast(const FastAttnArgs& a, cudaStream_t st);
// causal tiled attention for prefill rows (attn_prefill_fast.cu): rows are consecutive positions
// pos[0]+t of one sequence (slot[0]); n_splits / ws_* unused
cudaError_t launch_attention_prefill_fast(const FastAttnArgs& a, cudaStream_t st);
// the same on tcgen05 / TMEM (attn_prefill_tc.cu), head_dim 128; lbo_sbo_swap: descriptor probe, 0 in production
cudaError_t launch_attention_prefill_tc(const FastAttnArgs& a, int lbo_sbo_swap, cudaStream_t st);

// tcgen05 GEMM (fast numerics, M > 8)
struct TensorMap2D {
  alignas(64) unsigned char opaque[128];  // CUtensorMap
};
cudaError_t make_tensor_map_2d(TensorMap2D* out, const bf16* base, int rows, int K, int box_rows);
cudaError_t make_tensor_map_w3d(TensorMap2D* out, const bf16* base, int rows, int K, int kc, int ld = 0, int box_rows = 8);  // ld: row stride in elements (0 = K)
// KV pool as a 2-D tensor {head_dim 64, every (page, layer, K|V, head, slot) row}: box = box_rows slots x 128 bytes, 128-byte swizzle
cudaError_t make_tensor_map_kv(TensorMap2D* out, const bf16* pool, unsigned long long rows, int hd, int box_rows);
int tc_token_tile(int M);
struct TcGemm {
  const TensorMap2D* w[3];  // weight maps, box rows = 128
  const TensorMap2D* x;     // activation map [M, K], box rows = tc_token_tile(M)
  int rows[3];
  int nseg, M, K, epi;
  bf16* out;     // [M, ld_out]; EPI_SILU_MUL: segments are (gate, up), out has rows[0] columns
  int ld_out;
  float* ws;
  size_t ws_bytes;
  int max_splits;
  int* counters;   // zero-initialised ints for the fused split-K reduction (nullptr: separate finalize kernel)
  int n_counters;
  int w_static;    // weights are not produced by a preceding kernel: prefetch them before the PDL wait
};
cudaError_t launch_gemm_tcgen05(const TcGemm& t, int num_sms, cudaStream_t st, int* launches);

cudaError_t launch_rmsnorm_fast(const bf16* x, const bf16* w, bf16* y, size_t hidden, size_t n_tok, size_t x_stride,
                                cudaStream_t st, float eps = 1e-04f);

// sampling (reference tie-break + XORWOW)
cudaError_t launch_sample_topk(const bf16* logits, int* out_tokens, int n_rows, size_t vocab, float temperature,
                               int k, uint64_t seed, uint64_t seed_stride, const int* step_ptr, cudaStream_t st,
                               uint64_t subsequence = 0);  // XORWOW subsequence (the reference's last sample_topk_bf16 argument)

// repetition penalty over each row's context tokens (ops_ref_order.cu) + the engine's token history
cudaError_t launch_repetition_penalty(bf16* logits, const int* ctx, const int* ctx_row, const int* ctx_len, int len_bias, size_t fixed_len,
                                      size_t max_len, size_t ctx_stride, int n_rows, int vocab, float penalty, cudaStream_t st);
cudaError_t launch_history_append(int* hist, size_t stride, const int* ids, const int* pos, const int* slot, int n, cudaStream_t st);

// synthetic weights on the device (twin of the host generator)
cudaError_t launch_synth_fill(bf16* blob, size_t elem_begin, size_t n_elems, uint64_t seed, int kind,
                              cudaStream_t st);

cudaError_t launch_kv_fill(const KvGeom& kv, const int* block_row, int pos0, int n_pos, uint64_t seed,
                           cudaStream_t st);

// tensor-parallel greedy arg-max: cand[i] = (logit of the locally sampled index, global index)
struct TpCand {
  float val;
  int idx;
};
cudaError_t launch_tp_cand_make(const bf16* logits_local, const int* sampled_local, TpCand* cand, int n_rows, size_t vocab_local,
                                int vocab_offset, cudaStream_t st);
// x[i] = bf16(float(x[i]) + float(bf16(y[i]))): the reference's residual_add (residual_add.cu:7) applied to the fp32
// sum of the ranks' partial projections, rounded to bf16 once like the unsharded projection output
cudaError_t launch_residual_add_f32(bf16* x, const float* y, size_t n, cudaStream_t st);
// pick, per row, the best of tp candidates [tp][n_rows] in the reference's tie-break order
cudaError_t launch_tp_cand_merge(const TpCand* all, int tp, int n_rows, int* out_tokens, cudaStream_t st);

// step bookkeeping on the device: pos[i] += 1 ; ids <- sampled tokens
cudaError_t launch_advance(int* pos, int* ids, const int* sampled, int n, int* step_ptr, cudaStream_t st);

// ---------------------------------------------------------------- persistent decode step
// decode_mega.cu: ONE cooperative launch runs a whole decode step (all layers, lm_head,
// greedy arg-max, bookkeeping) for up to 64 sequences in reference-order arithmetic.
constexpr int MEGA_ROWS_PER_LAUNCH = 64;  // rows one persistent launch takes (decode_mega.cu MAX_ROWS); more rows = more launches
constexpr int MEGA_MAX_TP = 8;
constexpr int MEGA_TP_ROWS = 64;
constexpr int MEGA_TP_HEADER = 8192;  // bytes in front of the partial sums: flag words, generation word (byte 128), candidate table (byte 256, [8][64] x 8 bytes)
struct MegaLayer {
  const bf16 *in_ln, *q, *k, *v, *o, *q_norm, *k_norm, *post_ln, *gate, *up, *down;
};
struct MegaCand {
  float val;
  int idx;
};
struct MegaArgs {
  int H, I, L, n_q, n_kv, hd, V;
  const MegaLayer* layers;  // device array [L]
  const TensorMap2D* wmaps; // device array [7*L + 1]: q k v o gate up down per layer, then lm_head (make_tensor_map_w3d)
  const TensorMap2D* kvmap; // device: the KV pool as rows of one head (make_tensor_map_kv), nullptr = no TMA streaming of K/V
  const TensorMap2D* hmap;  // device: h [rows, I] as {64, rows, I/64} with boxes of 16 rows x KC (the A stream of the tile-split down_proj), or nullptr
  const bf16 *embed, *final_norm, *lm_head;
  const float *cos_t, *sin_t;
  int B;  // decode rows, one token per sequence
  int* ids;
  int* pos;
  const int* slot;
  const int* block_table;
  int max_pages, max_kv_len;
  int* rowstep;
  KvGeom kv;
  bf16 *x, *qkv, *att, *h, *logits;  // [B, H] [B, Dq+2Dkv] [B, Dq] [B, I] [B, V]
  bf16* xn;                          // [B, H] normalised rows (batches > 16: one CTA per row)
  int dist_norm;
  MegaCand* cand;                    // [grid, B] per-CTA arg-max candidates
  int* sampled;                      // [B]
  unsigned* bar;                     // grid barrier counter, zeroed by the launcher
  unsigned long long* prof;          // optional: globaltimer at every phase boundary (CTA 0), then clock64 at +prof_stride
  int prof_stride;
  int greedy, advance;
  int fast;                          // 1: fast (tolerance) numerics, <= 8 rows: split-K over the warps, parallel RMSNorm
  int kv_l2_prefetch;                // 1: request the next layer's cached K/V into L2 one phase group ahead of its attention
  int n_layers_run;                  // debug: stop after this many layers (0 = all, then lm_head)
  // tensor parallel inside the kernel (tp_size > 1, <= MEGA_TP_ROWS rows): n_q / n_kv / I / V and the weight maps describe
  // THIS rank's shard; o_proj / down_proj accumulators are written as fp32 partial sums into every rank's
  // exchange buffer over NVLink peer mappings, a cross-GPU flag barrier replaces the grid barrier behind those two
  // phases, and the next phase's row load adds the partial sums in rank order (one bf16 rounding, like tp_size 1)
  int tp_size, tp_rank;
  float* tp_part[MEGA_MAX_TP];       // rank r's exchange buffer [2][tp_size][MEGA_TP_ROWS][H] (peer-mapped for r != tp_rank)
  unsigned* tp_flag[MEGA_MAX_TP];    // rank r's flag words: flag[s] = number of exchanges rank s has completed
  unsigned* tp_epoch;                // local: exchanges completed before this launch (kept on the device: graph replay)
  MegaCand* tp_cand[MEGA_MAX_TP];    // rank r's candidate table [tp_size][MEGA_TP_ROWS]: best (logit, token) of every rank's vocabulary range
  int tp_vocab0;                     // first vocabulary row of this rank
  bf16* x2;                          // ping-pong partner of x for the residual stream
  int KC;  // k elements per weight tile = box depth of wmaps (decode_mega_kc at engine setup)
  // geometry, filled by the launcher (ph_*: per GEMM phase kind qkv/o/gate+up/down/lm_head)
  int ph_nu[5];  // units per round (NW, or NW / token tiles when the tiles of a unit are spread over warps)
  int ph_ts[5], ph_G[5], mtt;  // tile-split phases: CTA c owns token tile c % mtt and the units c / mtt + i * G (gemm_ts)
  int ph_q[5], ph_r[5], ph_nch[5], ph_adv_slot[5][2], ph_adv_par[5][2], ph_round_slot[5], ph_round_par[5];
  int n_slots, slot_bytes, act_bytes, off_act, off_ring, stream_down, attn_kstg, attn_off, off_red, attn_group, attn_hp, attn_tma;
};
// max rows the persistent kernel accepts for this model shape (0 = shape unsupported)
int decode_mega_kc(int H, int big);  // k elements per weight tile (box depth of the weight tensor maps); big: batches <= 8
bool decode_mega_supports(int H, int I, int L, int n_q, int n_kv, int hd, int B, int max_kv_len, int num_sms, int KC, int fast, int page_size);
int decode_mega_prof_slots(int L);
cudaError_t launch_decode_mega(MegaArgs a, int num_sms, cudaStream_t st);

// decode_gemv.cu: the fast-numerics decode step of <= DECODE_GEMV_MAX_ROWS sequences as one persistent kernel of
// memory-bound GEMVs (contiguous row ranges per CTA streamed by 1-D bulk copies, half-warp per weight row,
// split-KV flash decoding, no grid barrier between the phases: consumers poll per-layer activation buffers).  Takes the
// MegaArgs of the step (tensor maps / ring geometry unused) + a scratch buffer of decode_gemv_scratch_bytes() bytes that
// the caller filled with 0xFF bytes ONCE (the kernel leaves it in that state).
constexpr int DECODE_GEMV_MAX_ROWS = 4;
bool decode_gemv_supports(int H, int I, int L, int n_q, int n_kv, int hd, int B, int max_kv_len, int num_sms);
size_t decode_gemv_scratch_bytes(int H, int I, int L, int n_q, int n_kv, int hd, int num_sms);
cudaError_t launch_decode_gemv(MegaArgs a, void* scratch, int num_sms, cudaStream_t st, int dataflow = -1);  // dataflow: -1 default (1, or QIE_GEMV_DATAFLOW), 0 grid barriers, 1 polled buffers

}  // namespace qie
