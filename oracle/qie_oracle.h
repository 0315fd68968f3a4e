/*
 * qie_oracle.h -- CPU restatement (plain C) of the Rafae1130/qwen_inference_engine
 * forward path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (qwen_inference_engine_b200/) never links,
 * imports or executes anything under oracle/.
 *
 * Parity pin: the reference ships no golden vectors or tests, so
 * this restatement is pinned against outputs of the reference's OWN kernels compiled
 * from /root/reference/layers/src into oracle/_ref/libqie_ref.so and run on a B200
 * (oracle/gen_golden.py -> tests/golden/).  Until those fixtures exist the oracle is
 * "parity unpinned".
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/layers/).
 */
#ifndef QIE_ORACLE_H
#define QIE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t orc_bf16; /* raw bfloat16 bits */

typedef struct {
  int hidden;    /* H   */
  int inter;     /* I   (up_dim) */
  int layers;    /* L   */
  int n_q;       /* query heads */
  int n_kv;      /* kv heads */
  int head_dim;  /* hd  */
  int vocab;     /* V   */
  int context;   /* RoPE table rows; the reference uses 32786 (src/utills.cu:14) */
} orc_config;

/* ---- bf16 <-> fp32 (round-to-nearest-even, as __float2bfloat16) ---- */
float orc_bf2f(orc_bf16 v);
orc_bf16 orc_f2bf(float f);

/* ---- operators (one per reference __global__) ---- */

/* src/include.cpp:5-18 precompute_cos_sin */
void orc_precompute_cos_sin(float* cos_values, float* sin_values, int seq_len, int head_dim);

/* src/embedded_matrix.cu:5-17 embedding_matrix_func */
void orc_embedding(orc_bf16* out, const orc_bf16* table, const int* ids, size_t hidden, size_t n_tok);

/* src/normalization.cu:5-26 rmsNorm */
void orc_rmsnorm(const orc_bf16* x, const orc_bf16* w, orc_bf16* y, size_t hidden, size_t n_tok);
void orc_rmsnorm_eps(const orc_bf16* x, const orc_bf16* w, orc_bf16* y, size_t hidden, size_t n_tok, float eps);

/* src/matrix_mul.cu:165-288 matrix_mul (via helpers.cuh:81 launch_matmul):
 * C[M,K] = A[M,N] * B[K,N]^T, fp32 accumulate over N in 16-wide chunks, bf16 store.
 * Argument names follow the reference (N = inner dim, K = output columns). */
void orc_matmul(const orc_bf16* A, const orc_bf16* B, orc_bf16* C, int M, int N, int K);

/* src/qk_norm.cu:43-80 qkNorm (in place; one block per head, tree reduction) */
void orc_qknorm(orc_bf16* qk, const orc_bf16* w, int head_dim, int n_tok, int row_dim, int n_heads);
void orc_qknorm_eps(orc_bf16* qk, const orc_bf16* w, int head_dim, int n_tok, int row_dim, int n_heads, float eps);
/* Qwen2.5 / HF semantics that the reference does not have */
void orc_rope_half(const float* cos_values, const float* sin_values, orc_bf16* x, int n_tok, int head_dim, int row_dim, int n_heads);
void orc_add_bias(orc_bf16* y, const orc_bf16* b, int n_tok, int dim);

/* src/RoPE.cu:6-22 RoPE (in place, interleaved pairs). cos/sin point at the row of
 * the FIRST token (launch_rope_single pre-offsets them, helpers.cuh:143-147). */
void orc_rope(const float* cos_values, const float* sin_values, orc_bf16* x, int n_tok, int head_dim,
              int row_dim, int n_heads);

/* src/SiLU.cu:10-23 activation ; src/element_add.cu:4-13 element_mul ;
 * src/residual_add.cu:7-18 residual_add */
void orc_silu(orc_bf16* x, size_t n);
void orc_elem_mul(const orc_bf16* a, const orc_bf16* b, orc_bf16* c, size_t n);
void orc_residual_add(orc_bf16* a, const orc_bf16* b, size_t n);

/* Paged KV cache in the reference's layout (include/iengine.cuh:42-48,
 * src/include_cuda.cu:165-279): page p holds positions [p*page_size,(p+1)*page_size),
 * element offset inside a page = ((pos%page_size)*L + layer)*Dkv + col. */
typedef struct {
  int page_size, n_layers, kv_dim, n_pages;
  orc_bf16** k_pages;
  orc_bf16** v_pages;
} orc_kv;

orc_kv* orc_kv_new(int page_size, int n_layers, int kv_dim);
void orc_kv_free(orc_kv* kv);
/* src/include_cuda.cu:165-231 / 233-279: copy n_tok rows of K,V (row stride kv_dim)
 * for `layer` into positions [pos0, pos0+n_tok); pages are created on demand. */
void orc_kv_store(orc_kv* kv, int layer, int pos0, int n_tok, const orc_bf16* K, const orc_bf16* V);

/* src/self_attension.cu:10-149 selfattention (grid = q heads, block = head_dim). */
void orc_attention(const orc_bf16* Q, orc_bf16* out, size_t seq_len_q, size_t seq_len_kv, size_t head_dim,
                   size_t q_dim, size_t kv_dim, int n_q_heads, int n_kv_heads, int causal,
                   size_t q_abs_base, int layer, const orc_kv* kv);

/* src/logit_decode.cu:149-274 topk_temperature_softmax_sampling_kernel_bf16
 * (block of 256 threads, k rounds of blockArgMax :19-33, XORWOW draw :255-268). */
int orc_sample_topk(const orc_bf16* logits, float temperature, int k, size_t vocab, uint64_t seed,
                    uint64_t subseq);
/* the k=1 special case, spelled out as the closed-form tie-break rule */
int orc_argmax_ref_tiebreak(const orc_bf16* logits, size_t vocab);
void orc_repetition_penalty(orc_bf16* logits, const int* context_tokens, size_t context_len, int vocab, float penalty);

/* ---- checkpoint (model_files/meta_data.txt + weights.bin) ---- */
typedef struct orc_model orc_model;

/* Parses the text format src/tensor_parser.cpp:19-28 emits and maps weights.bin. */
orc_model* orc_model_load(const char* meta_path, const char* weights_path, int head_dim_hint,
                          int context);
void orc_model_free(orc_model* m);
const orc_config* orc_model_config(const orc_model* m);
/* rope_half != 0: HF rotate_half RoPE; eps: RMSNorm / q,k-norm epsilon (<= 0: the reference's 1e-4).  Projection biases are
 * applied whenever the checkpoint holds self_attn.{q,k,v}_proj.bias. */
void orc_model_set_semantics(orc_model* m, int rope_half, float eps);
/* pointer to a tensor by short name / layer (layer ignored for globals); NULL if absent */
const orc_bf16* orc_model_tensor(const orc_model* m, const char* short_name, int layer, size_t* n_elems);

/* ---- sequences + forward (src/qwen_main.cu:64-417 llm) ---- */
typedef struct orc_seq orc_seq;
orc_seq* orc_seq_new(const orc_model* m, int page_size);
void orc_seq_free(orc_seq* s);
int orc_seq_len(const orc_seq* s);

/* optional per-layer activation dump hook: called with a tag, layer, pointer, count */
typedef void (*orc_dump_fn)(void* user, const char* tag, int layer, const orc_bf16* data, size_t n);
void orc_set_dump(orc_seq* s, orc_dump_fn fn, void* user);

/* prefill branch qwen_main.cu:74-247; returns sampled token (top-k `k`, temperature) and,
 * if logits_out != NULL, the V bf16 logits of the last prompt token. */
int orc_prefill(orc_seq* s, const int* ids, int n_tok, int topk, float temperature, uint64_t seed,
                orc_bf16* logits_out);
/* decode branch qwen_main.cu:250-404 (one token). */
int orc_decode(orc_seq* s, int token, int topk, float temperature, uint64_t seed, orc_bf16* logits_out);

/* threads used by orc_matmul rows (1 = the scalar port; >1 = row-parallel pthreads) */
void orc_set_threads(int n);

/* ---- synthetic checkpoint generator (oracle-side twin of the product's) ---- */
/* value of global element g of a tensor class; kind 0 = matrix N(0,0.02)-like,
 * kind 1 = norm vector 1 + 0.05*n. Pure integer hash + one fp32 multiply (+ one add). */
orc_bf16 orc_synth_value(uint64_t seed, uint64_t g, int kind);
int orc_synth_write(const orc_config* cfg, uint64_t seed, const char* meta_path, const char* weights_path);

#ifdef __cplusplus
}
#endif
#endif
