/*
 * qie_b200.h -- C ABI of libqie_b200.so, the B200-native (sm_100a) replacement for the
 * decode/prefill forward of Rafae1130/qwen_inference_engine.
 *
 * Everything here is `extern "C"`, plain pointers / integers / sizes; no C++ or torch
 * types cross the boundary.  Two levels, mirroring the two levels of the reference's
 * header API; citations are relative to /root/reference/layers/:
 *
 *   (1) operator level  -- one entry point per reference __global__ / launch_* wrapper
 *                          (include/layers_include.cuh:15-35, include/helpers.cuh:45-166).
 *                          Device pointers in, device pointers out, explicit stream.
 *   (2) driver level    -- what iengine's main() drives (include/iengine.cuh:51-55,
 *                          include/utils.hh:90-105, include/tensor_parser.hh:47-52):
 *                          checkpoint load, sequences, KV pages, prefill / decode.
 *
 * Error behaviour: every function returns 0 on success or a negative QIE_E* code;
 * qie_last_error() returns a thread-local message.  Token ids are never used to carry
 * errors (the reference returns token 0 on CUDA failure, src/qwen_main.cu:135,262).
 * There is no CPU fallback: without a CUDA device every compute entry point fails with
 * QIE_ECUDA.
 */
#ifndef QIE_B200_H
#define QIE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QIE_OK 0
#define QIE_EINVAL (-1)   /* bad argument / unsupported shape */
#define QIE_ECUDA (-2)    /* CUDA runtime or launch failure   */
#define QIE_EIO (-3)      /* checkpoint file problem          */
#define QIE_ENOMEM (-4)   /* device memory or KV pages exhausted */
#define QIE_ESTATE (-5)   /* call not valid in the current sequence state */

typedef void* qie_stream;      /* cudaStream_t */
typedef uint16_t qie_bf16;     /* raw __nv_bfloat16 bits */

const char* qie_last_error(void);
/* ABI version; bumped on any signature change. */
int qie_abi_version(void); /* currently 5 */

/* ------------------------------------------------------------------------------------
 * (1) operator level.  All pointers are DEVICE pointers unless named h_*.
 * Argument order follows the reference wrapper each one replaces.
 * ---------------------------------------------------------------------------------- */

/* embedding_matrix_func, src/embedded_matrix.cu:5-17: out[t,:] = table[ids[t],:] */
int qie_embedding(qie_bf16* out, const qie_bf16* table, const int* ids, size_t hidden, size_t n_tok,
                  qie_stream st);

/* launch_rms -> rmsNorm, helpers.cuh:45-49 / src/normalization.cu:5-26.
 * y = bf16((x / sqrt(sum(x^2)/hidden + 1e-4)) * w), sum in the reference's sequential
 * fp32 FFMA order. */
int qie_rmsnorm(const qie_bf16* x, const qie_bf16* w, qie_bf16* y, size_t hidden, size_t n_tok, qie_stream st);

/* launch_matmul / proj -> matrix_mul, helpers.cuh:81-106,132-138 / src/matrix_mul.cu:165.
 * C[M,K] = A[M,N] * B[K,N]^T (B in HF [out,in] layout); N = inner dim, K = out columns,
 * exactly the reference's argument meaning.  fp32 accumulation through m16n8k16 bf16
 * tensor-core MMAs fed in the reference's chunk order, one bf16 rounding at the end. */
int qie_matmul(const qie_bf16* A, const qie_bf16* B, qie_bf16* C, int M, int N, int K, qie_stream st);

/* launch_qknorm -> qkNorm, helpers.cuh:140-142 / src/qk_norm.cu:43-80 (in place). */
int qie_qknorm(qie_bf16* x, const qie_bf16* w, int head_dim, int n_tok, int row_dim, int n_heads, qie_stream st);

/* launch_rope / launch_rope_single -> RoPE, helpers.cuh:51-55,143-147 / src/RoPE.cu:6-22.
 * cos/sin: fp32 tables [context, head_dim/2]; token t uses row pos0 + t. In place. */
int qie_rope(const float* cos_t, const float* sin_t, qie_bf16* x, int n_tok, int pos0, int head_dim, int row_dim,
             int n_heads, qie_stream st);

/* precompute_cos_sin, src/include.cpp:5-18 (host libm, fp32). HOST pointers. */
int qie_precompute_cos_sin(float* h_cos, float* h_sin, int seq_len, int head_dim);

/* launch_act -> activation (SiLU), launch_elem -> element_mul, launch_resadd ->
 * residual_add; helpers.cuh:108-119. */
int qie_silu(qie_bf16* x, size_t n, qie_stream st);
int qie_elem_mul(const qie_bf16* a, const qie_bf16* b, qie_bf16* c, size_t n, qie_stream st);
int qie_residual_add(qie_bf16* a, const qie_bf16* b, size_t n, qie_stream st);

/* KV pool geometry used by the operator-level attention / store entry points.
 * B200 layout (DESIGN.md "KV pool"): one allocation,
 *   pool[page][layer][k|v][kv_head][slot][head_dim]   (bf16)
 * replacing the reference's linked list of per-page cudaMallocs laid out
 * [slot][layer][kv_dim] (include/iengine.cuh:42-48, src/include_cuda.cu:165-279). */
typedef struct {
  qie_bf16* pool;
  int n_pages, page_size, n_layers, n_kv_heads, head_dim;
} qie_kv_view;

/* kv_copy_layer_to_cache_{prefill,decode}, src/include_cuda.cu:165-279: scatter rows of
 * K,V ([n_tok, n_kv*hd]) of `layer` to positions pos[t] of the sequence whose page ids
 * are block_table[slot[t]*max_pages + ...]. */
int qie_kv_store(const qie_kv_view* kv, int layer, const qie_bf16* K, const qie_bf16* V, const int* pos,
                 const int* slot, const int* block_table, int max_pages, int n_tok, qie_stream st);

/* launch_attn -> selfattention, helpers.cuh:121-130 / src/self_attension.cu:10-149.
 * Row t (query token) of Q attends to cache positions 0..pos[t] of its sequence
 * (prefill: causal; decode: the whole cache).  Reference arithmetic order. */
int qie_attention(const qie_kv_view* kv, int layer, const qie_bf16* Q, qie_bf16* out, const int* pos,
                  const int* slot, const int* block_table, int max_pages, int n_tok, int n_q_heads,
                  qie_stream st);

/* The same two operators over the REFERENCE's own cache: a list of pages, each a K and a V device buffer of
 * page_size * n_layers * kv_dim bf16 laid out [slot][layer][kv_dim] (include/iengine.cuh:42-48, src/iengine.cu:89-96,
 * element ((pos % page_size) * n_layers + layer) * kv_dim, src/include_cuda.cu:165-279).  d_k_pages / d_v_pages are
 * DEVICE arrays of n_pages device pointers in list order -- what the reference's kernel finds by walking
 * ptr_to_next_page (src/self_attension.cu:47-53,124-126).  All rows belong to that one sequence; row t attends to
 * positions 0..pos[t].  These are what include/layers/iengine_compat.hh's launch_attn(..., page_table*, page_size) and
 * kv_copy_layer_to_cache_{prefill,decode} call, so a caller that owns reference pages needs no engine. */
int qie_attention_pagelist(const qie_bf16* const* d_k_pages, const qie_bf16* const* d_v_pages, int n_pages, int page_size,
                           int n_layers, int layer, int n_kv_heads, int head_dim, const qie_bf16* Q, qie_bf16* out, const int* pos,
                           int n_tok, int n_q_heads, qie_stream st);
int qie_kv_store_pagelist(qie_bf16* const* d_k_pages, qie_bf16* const* d_v_pages, int n_pages, int page_size, int n_layers,
                          int layer, int kv_dim, const qie_bf16* K, const qie_bf16* V, int pos0, int n_tok, qie_stream st);

/* sample_topk_bf16 -> topk_temperature_softmax_sampling_kernel_bf16,
 * helpers.cuh:157-166 / src/logit_decode.cu:149-274. One row of `vocab` bf16 logits per
 * sequence; same arg-max tie-break and XORWOW stream (subsequence 0) as the reference.
 * out_tokens: device int[n_rows]. seeds: per-row seed = seed + row*seed_stride. */
int qie_sample_topk(const qie_bf16* logits, int* out_tokens, int n_rows, size_t vocab, float temperature, int k,
                    uint64_t seed, uint64_t seed_stride, qie_stream st);

/* The same with the reference's full RNG addressing: sample_topk_bf16's last argument is handed to the kernel as
 * the cuRAND SUBSEQUENCE -- curand_init(seed, subsequence, 0) -- helpers.cuh:157-166, logit_decode.cu:256-257
 * (llm() always passes 0 and folds the step into the seed, qwen_main.cu:241,388). */
int qie_sample_topk_subseq(const qie_bf16* logits, int* out_tokens, int n_rows, size_t vocab, float temperature, int k,
                           uint64_t seed, uint64_t seed_stride, uint64_t subsequence, qie_stream st);

/* apply_repetition_penalty_kernel, include/layers_include.cuh:33 -- declared by the reference with exactly these
 * arguments, never defined or launched.  Conventional semantics (CTRL / HF): every DISTINCT id in context_tokens
 * (DEVICE int[context_len]) has its logit divided by `penalty` when positive and multiplied when not (fp32 IEEE, one
 * bf16 rounding); ids outside [0, vocab_size) are ignored.  One row of logits, in place. */
int qie_repetition_penalty(qie_bf16* logits, const int* context_tokens, size_t context_len, int vocab_size, float penalty,
                           qie_stream st);

/* FAST-numerics operators (same contracts as qie_matmul / qie_attention, results within
 * the 1e-2 bf16 tolerance instead of bit-exact): tcgen05/TMEM/TMA GEMM with split-K, and
 * split-KV flash-decoding (one query token per row; n_splits = 0 picks a default).
 * Scratch space is allocated internally on first use. */
int qie_matmul_fast(const qie_bf16* A, const qie_bf16* B, qie_bf16* C, int M, int N, int K, qie_stream st);
int qie_attention_decode_fast(const qie_kv_view* kv, int layer, const qie_bf16* Q, qie_bf16* out, const int* pos,
                              const int* slot, const int* block_table, int max_pages, int n_tok, int n_q_heads,
                              int n_splits, qie_stream st);

/* FAST-numerics causal attention for PREFILL rows (same contract as qie_attention with the
 * restriction that the n_tok rows are consecutive positions pos[0]+t of ONE sequence
 * slot[0], which is what llm()'s prefill branch produces, qwen_main.cu:160-176):
 * FlashAttention-2 style tiling on mma.sync, K/V streamed from the paged pool. */
int qie_attention_prefill_fast(const qie_kv_view* kv, int layer, const qie_bf16* Q, qie_bf16* out, const int* pos,
                               const int* slot, const int* block_table, int max_pages, int n_tok, int n_q_heads,
                               qie_stream st);

/* The same prefill attention on tcgen05 / TMEM (head_dim 128): S = QK^T and O_t = PV as tcgen05.mma with the
 * accumulators in tensor memory, softmax by one thread per query row over tcgen05.ld.  variant = 0. */
int qie_attention_prefill_tc(const qie_kv_view* kv, int layer, const qie_bf16* Q, qie_bf16* out, const int* pos,
                             const int* slot, const int* block_table, int max_pages, int n_tok, int n_q_heads, int variant,
                             qie_stream st);

/* ------------------------------------------------------------------------------------
 * (2) driver level
 * ---------------------------------------------------------------------------------- */

typedef struct {
  int hidden, inter, layers, n_q, n_kv, head_dim, vocab, context;
} qie_config;

typedef struct {
  int device;            /* CUDA device ordinal */
  int page_size;         /* KV positions per page (reference main() uses 4, iengine.cu:334) */
  int max_pages;         /* KV pool size in pages; 0 = size from kv_bytes */
  size_t kv_bytes;       /* KV pool bytes if max_pages == 0; 0 = 1 GiB default */
  int max_seqs;          /* concurrently live sequences */
  int max_batch_tokens;  /* rows per forward (prefill chunk or decode batch) */
  int context;           /* RoPE table rows; 0 = 32786 as the reference (utills.cu:14) */
  int head_dim_hint;     /* used only if the checkpoint has no q_norm tensor */
  int use_graph;         /* replay decode steps from a CUDA graph */
  int tp_rank, tp_size;  /* tensor-parallel shard (1 = off) */
  int numerics;          /* QIE_NUMERICS_REFERENCE_ORDER (default) or QIE_NUMERICS_FAST */
  int semantics;         /* QIE_SEMANTICS_REFERENCE (default) or QIE_SEMANTICS_HF */
  float rms_eps;         /* RMSNorm / q,k-norm epsilon; 0 = by semantics (1e-4 reference, 1e-6 HF) */
} qie_engine_opts;

/* Model semantics.  REFERENCE: what the reference's kernels compute -- interleaved-pair RoPE (src/RoPE.cu:6-22), eps 1e-4
 * (src/normalization.cu:9, src/qk_norm.cu:46).  HF: what the published Qwen2.5 / Qwen3 checkpoints were trained with --
 * half-rotation RoPE (transformers' rotate_half) and eps 1e-6.  Independently of this switch, and driven by the
 * checkpoint's tensors only: per-head q/k-norm is applied iff self_attn.{q,k}_norm.weight exist (Qwen3: yes, Qwen2.5:
 * no), projection biases are added iff self_attn.{q,k,v}_proj.bias exist (Qwen2.5: yes; y = bf16(bf16(acc) + bias)).
 * HF semantics and biases run on the per-operator launch path (the persistent decode kernel keeps the reference's). */
#define QIE_SEMANTICS_REFERENCE 0
#define QIE_SEMANTICS_HF 1

/* REFERENCE_ORDER: every kernel reproduces the reference's fp32 operation order; results
 * are bit-identical to the reference's kernels (greedy tokens, logits, activations).
 * FAST: rows > 8 use the tcgen05/TMEM GEMM (+ split-K), decode attention is split-KV
 * flash-decoding, RMSNorm reduces in parallel; results agree with the reference within
 * the 1e-2 bf16 tolerance of BASELINE.json, not bit for bit. */
#define QIE_NUMERICS_REFERENCE_ORDER 0
#define QIE_NUMERICS_FAST 1

void qie_engine_opts_default(qie_engine_opts* o);

typedef struct qie_engine qie_engine;

/* Synthetic checkpoint in the reference's on-disk format: weights.bin + the text
 * meta_data.txt that operator<< emits (src/tensor_parser.cpp:19-28). HOST side only. */
int qie_synth_checkpoint_write(const qie_config* cfg, uint64_t seed, const char* meta_path,
                               const char* weights_path);

/* HOST only (no GPU needed): parse meta_data.txt as build_indexed_tensors would index it
 * (tensor_parser.cpp:132-165) and derive the model shape from the tensor shapes.
 * head_dim_hint is used only when the checkpoint has no q_norm tensor. */
int qie_checkpoint_inspect(const char* meta_path, int head_dim_hint, qie_config* cfg_out, size_t* total_bytes,
                           int* n_tensors);

/* HOST only: parsed_tensors(), src/tensor_parser.cpp:31-129 -- HF safetensors shards (in the order given)
 * -> weights.bin + meta_data.txt exactly as the reference lays them out: keys starting "model." or "lm_",
 * byte-lexicographic key order inside a shard, running [begin,end) offsets, every "lm_" tensor named
 * "logits".  The bytes are copied tensor by tensor from each shard's data_offsets.  BF16 only.
 * tie_lm_head != 0: a checkpoint without an lm_head tensor (tied embeddings) gets a second copy of
 * model.embed_tokens.weight appended as "lm_head.weight"; with 0 such a checkpoint is refused (QIE_EIO). */
int qie_convert_safetensors(const char* const* shard_paths, int n_shards, const char* meta_path, const char* weights_path,
                            int tie_lm_head, size_t* total_bytes, int* n_tensors);

/* build_indexed_tensors + load_all_weights_to_gpu_chunked + initialize_model_buffers,
 * tensor_parser.cpp:132-165, iengine.cu:117-223, utills.cu:4-129: parse meta_data.txt,
 * upload weights.bin as ONE device blob, derive the model shape from tensor shapes. */
int qie_engine_create(const char* meta_path, const char* weights_path, const qie_engine_opts* opts,
                      qie_engine** out);
/* Same engine over a weight blob that is ALREADY on the device (the reference's main()
 * uploads weights.bin itself, iengine.cu:117-223, and hands g_gpu_weights_buffer to llm()).
 * The blob is borrowed, never freed by the engine. */
int qie_engine_create_from_blob(const char* meta_path, void* device_blob, const qie_engine_opts* opts,
                                qie_engine** out);
/* Same engine, weights generated on the device by the same counter hash (no file). */
int qie_engine_create_synthetic(const qie_config* cfg, uint64_t seed, const qie_engine_opts* opts,
                                qie_engine** out);
void qie_engine_destroy(qie_engine* e);

/* Tensor parallelism (BASELINE configs[4]; the reference is single-GPU): one
 * process per GPU, every rank creates the engine with the same checkpoint and
 * opts.tp_rank / opts.tp_size, then all ranks call qie_engine_tp_connect with the 128-byte
 * id that rank 0 obtained from qie_tp_unique_id (carried between the processes by the
 * caller: torch.distributed broadcast, a file, MPI ...).  Attention heads, the MLP
 * intermediate dimension and the vocabulary are sharded; o_proj / down_proj partial sums
 * are all-reduced with NCCL over NVLink; greedy sampling merges per-rank candidates in the
 * reference's tie-break order, every rank returns the same tokens.  Prefill/decode calls
 * are collective: all ranks must make the same calls in the same order. */
/* HOST only: what rank tp_rank of tp_size owns. out8 = {n_q, n_kv, inter, vocab} local sizes
 * then {q_row0, kv_row0, inter0, vocab0}: first q_proj / k,v_proj output row (= o_proj input
 * column), first gate/up row (= down_proj input column), first lm_head row.  Vocabulary
 * shards start at multiples of 256 (the sampler's tie-break key is idx mod 256). */
int qie_tp_plan(const qie_config* cfg, int tp_rank, int tp_size, int* out8);
int qie_tp_unique_id(void* out128);
int qie_engine_tp_connect(qie_engine* e, const void* id128);
int qie_engine_get_config(const qie_engine* e, qie_config* out);
/* device pointer into the weight blob for (short_name, layer), as assign_weight_pointer
 * helpers.cuh:18-29; *n_elems receives the element count. NULL on miss. */
const qie_bf16* qie_engine_weight(const qie_engine* e, const char* short_name, int layer, size_t* n_elems);
int qie_engine_kv_view(const qie_engine* e, qie_kv_view* out);
/* what a caller that batches requests has to respect: rows one decode step takes (min of max_batch_tokens and the
 * logits rows), KV pages one sequence can hold, sequence slots */
int qie_engine_limits(const qie_engine* e, int* max_decode_rows, int* max_pages_per_seq, int* max_seqs);
qie_stream qie_engine_stream(const qie_engine* e);

/* sampling parameters applied by prefill/decode (defaults: top-k 1 = greedy).
 * The reference hard-codes k=50, T=1.0 prefill / 0.7 decode, seed 1234(+step)
 * (qwen_main.cu:241,381-388). per-step seed = seed + step when add_step != 0. */
int qie_engine_set_sampling(qie_engine* e, int topk, float temperature_prefill, float temperature_decode,
                            uint64_t seed, int add_step);

/* Driver-level repetition penalty: with penalty != 1 the engine keeps every sequence's tokens (prompt + generated) on
 * the device and applies qie_repetition_penalty over them to the logits of every sampled row before the sampler
 * (greedy or top-k); 1 switches it off.  Positions filled by qie_seq_fill_synthetic / qie_seq_kv_write carry no token
 * and are skipped.  Not available with tensor parallelism. */
int qie_engine_set_repetition_penalty(qie_engine* e, float penalty);

/* create_new_sequence + create_page_list (iengine.cu:25-47,73-87). Returns a slot id. */
int qie_seq_new(qie_engine* e, int* seq);
/* free_page_list + destroy_model_buffers (iengine.cu:98-109, utills.cu:147-205). */
int qie_seq_free(qie_engine* e, int seq);
int qie_seq_len(const qie_engine* e, int seq);
/* KV offload (the experiment commented out in iengine.cu:376-429): copy the sequence's pages to pinned host memory
 * and return them to the pool / take fresh pages and copy the bytes back.  A swapped-out sequence keeps its length
 * and step; prefill/decode on it return QIE_ESTATE until it is swapped in; generation then continues bit for bit. */
int qie_seq_swap_out(qie_engine* e, int seq);
int qie_seq_swap_in(qie_engine* e, int seq);
int qie_kv_pages_free(const qie_engine* e);

/* Cache rows [pos0, pos0+n) of a sequence <-> HOST buffers in the REFERENCE's page layout
 * [position][layer][n_kv*head_dim] (iengine.cu:352; include_cuda.cu:165-279 writes element
 * ((pos % page_size)*L + layer)*Dkv of a page): a read with n = page_size is one reference page.
 * Write grows the sequence (pages are taken from the pool) -- the host-side counterpart of
 * kv_copy_layer_to_cache_prefill for callers that own K/V rows (page-list import, tests). */
int qie_seq_kv_read(qie_engine* e, int seq, int pos0, int n, qie_bf16* h_K, qie_bf16* h_V);
int qie_seq_kv_write(qie_engine* e, int seq, int pos0, int n, const qie_bf16* h_K, const qie_bf16* h_V);

/* llm() with state == prefill, qwen_main.cu:74-247. h_ids: HOST int32[n]. The sampled
 * token is written to *h_token (HOST). */
int qie_prefill(qie_engine* e, int seq, const int32_t* h_ids, int n, int32_t* h_token);
/* llm() with state == decode, qwen_main.cu:250-404, for a batch of sequences:
 * h_seqs[i] consumes h_tokens_in[i] and produces h_tokens_out[i]. HOST buffers; one H2D
 * and one D2H per call. */
int qie_decode_step(qie_engine* e, const int* h_seqs, const int32_t* h_tokens_in, int n, int32_t* h_tokens_out);
/* Greedy/top-k multi-step decode that feeds sampled tokens back on the device:
 * h_tokens_out is HOST int32[steps*n] (step-major). h_tokens_in is the first input. */
int qie_decode_run(qie_engine* e, const int* h_seqs, const int32_t* h_tokens_in, int n, int steps,
                   int32_t* h_tokens_out);

/* Device-resident variant used for kernel-only timing: token ids stay on the device
 * (the previous step's samples), no host copies, no synchronisation. */
int qie_decode_step_device(qie_engine* e, const int* h_seqs, int n);
int qie_sync(qie_engine* e);

/* Continuous batching (what iengine.cuh:23-37 `State`, the commented second sequence of iengine.cu:369-373 and
 * the EOS test of qwen_main.cu:257 sketch).  Requests queue up; every qie_sched_step admits waiting requests
 * FIFO while sequence slots and KV pages allow (pages for prompt + max_new_tokens are reserved at admission, so a
 * decode step never runs out), then advances ALL running requests with one batched decode step; a request that
 * samples `eos_token` (151645 in the reference; -1 = none) or reaches max_new_tokens leaves the batch and its pages
 * are recycled.  In reference-order numerics a request's tokens do not depend on the batch it shares (tested).
 * qie_sched_step returns the number of unfinished requests or a negative error; qie_sched_result copies the tokens
 * generated so far (including the EOS token) and returns their count; *finished is 0 (queued / running), 1 (done) or
 * the negative QIE_E* code that failed THIS request.  qie_sched_submit refuses what could never run (prompt +
 * max_new_tokens beyond the context or the per-sequence page table, token ids outside the vocabulary); a request whose
 * prefill or decode fails later is finished as failed and its sequence and pages are released -- the others go on. */
typedef struct qie_scheduler qie_scheduler;
int qie_sched_create(qie_engine* e, int max_running, int eos_token, qie_scheduler** out);
void qie_sched_destroy(qie_scheduler* s);
int qie_sched_submit(qie_scheduler* s, const int32_t* ids, int n, int max_new_tokens, int* request_id);
int qie_sched_step(qie_scheduler* s);
int qie_sched_result(const qie_scheduler* s, int request_id, int32_t* out, int max_tokens, int* finished);
int qie_sched_stats(const qie_scheduler* s, long* steps, long* decode_rows, long* prefills, int* running, int* waiting);

/* Parity hooks (the reference's dump_device_bf16, qwen_main.cu:42-61). When capture is
 * on, every forward keeps per-layer activations; qie_capture_read copies one of them to
 * HOST memory. Tags: input_norm q k v attn x_attn mlp_h x_out (layer >= 0), logits
 * (layer = -1). Returns the element count or a negative error. */
int qie_capture_enable(qie_engine* e, int on);
long qie_capture_read(qie_engine* e, const char* tag, int layer, qie_bf16* h_out, size_t max_elems);

/* launch statistics of the last forward: kernels launched by this library. */
long qie_launch_count(const qie_engine* e);

/* Measurement helpers (bench.py). qie_seq_fill_synthetic appends n_pos positions of seeded
 * random K/V to a sequence WITHOUT running the model (stand-in for a long prefill when only
 * decode is timed). qie_decode_step_profile runs one decode step eagerly with a CUDA event
 * pair around every launch and returns, per kernel class, the summed device time (ms) and
 * the launch count; class names come from qie_kernel_kind_name (NULL past the last). */
int qie_seq_fill_synthetic(qie_engine* e, int seq, int n_pos, uint64_t seed);
int qie_decode_step_profile(qie_engine* e, const int* h_seqs, const int32_t* h_tokens_in, int n, float* ms_by_kind,
                            int* launches_by_kind, int n_kinds);
const char* qie_kernel_kind_name(int kind);

/* Persistent decode kernel (decode_mega.cu): reference-order decode steps of <= 8 rows run
 * as ONE cooperative launch.  qie_engine_set_int keys: "mega" (0/1, default 1 unless
 * QIE_MEGA=0), "mega_layers_run" (debug: stop after N layers, 0 = whole step), "mega_prof"
 * (0/1: record a device timestamp at every phase boundary).  qie_decode_uses_mega tells
 * whether a step of n rows at this KV length would take that path.
 * qie_mega_prof_read copies the timestamps (ns, globaltimer) of the last profiled step to
 * HOST memory: [0] start, then 16 per layer (QKV: rows loaded, normed, GEMM done, barrier;
 * attention: done, barrier; O: loaded, GEMM, barrier; gate/up: loaded, normed, GEMM, barrier;
 * down: loaded, GEMM, barrier), then lm_head (loaded, normed, GEMM, barrier) and sampling;
 * the buffer holds 16*layers + 8 slots of globaltimer ns followed by as many SM cycle-counter
 * values taken at the same points (16*layers + 6 of each are used), then 10 cycle counters:
 * per GEMM phase kind (qkv, o, gate/up, down, lm_head) the cycles warp 0 of CTA 0 waited for
 * weights and spent in its MMA loop; returns the values copied.
 * qie_engine_read_activation copies an activation buffer of the last forward to HOST
 * memory (names: x qkv att h logits sampled kv); returns the byte count copied. */
int qie_engine_set_int(qie_engine* e, const char* key, long value);
int qie_decode_uses_mega(const qie_engine* e, int n_rows, int kv_len);
long qie_mega_prof_read(qie_engine* e, uint64_t* h_out, size_t max_values);
long qie_engine_read_activation(qie_engine* e, const char* name, void* h_out, size_t max_bytes);
/* Layer-isolation parity hook of the per-operator forward: qie_engine_set_int keys "layer_first" / "layer_count"
 * (run only these layers, no final norm / lm_head; 0 = the whole model) and "inject_x" (1: the residual stream is
 * what qie_engine_write_activation(e, "x", rows, bytes) stored, the embedding gather is skipped).  Lets a test feed
 * layer l of one numerics mode the layer input of another and compare ONE layer's output (north-star 1e-2 bar). */
long qie_engine_write_activation(qie_engine* e, const char* name, const void* h_in, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif
