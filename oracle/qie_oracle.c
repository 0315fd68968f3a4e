/*
 * qie_oracle.c -- CPU restatement (plain C99) of the Rafae1130/qwen_inference_engine
 * decode/prefill forward.  TEST INFRASTRUCTURE ONLY -- see qie_oracle.h.
 *
 * Build: gcc -O2 -ffp-contract=off (no implicit FMA; every FMA the reference's SASS
 * contains is written as fmaf() here, each one noted where it occurs).
 *
 * What cannot be restated bit-for-bit on a CPU: the reference GEMM accumulates through
 * legacy tensor-core HMMA.16816.F32.BF16 (wmma, src/matrix_mul.cu:199-256) whose
 * intra-chunk summation order/rounding is not documented, and device expf differs from
 * glibc expf in the last ulp.  Everything else (sum orders, tree shapes, rounding
 * points, tie-breaks, RNG stream) is restated exactly.
 */
#define _GNU_SOURCE
#include "qie_oracle.h"

#include <fcntl.h>
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

/* ------------------------------------------------------------------ bf16 */
float orc_bf2f(orc_bf16 v) {
  uint32_t u = (uint32_t)v << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

orc_bf16 orc_f2bf(float f) { /* __float2bfloat16 : round-to-nearest-even */
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return 0x7fff; /* NaN */
  u += 0x7fffu + ((u >> 16) & 1u);
  return (orc_bf16)(u >> 16);
}

/* ------------------------------------------------------------------ threads */
static int g_threads = 1;
void orc_set_threads(int n) { g_threads = n < 1 ? 1 : n; }

/* ------------------------------------------------------------------ RoPE tables
 * src/include.cpp:5-18.  All operands are float, so C++ overload resolution in the
 * reference picks powf; pos*theta is int*float -> float; cosf/sinf are glibc's. */
void orc_precompute_cos_sin(float* cos_values, float* sin_values, int seq_len, int head_dim) {
  float base = 1000000;
  int half = head_dim / 2;
  for (int i = 0; i < half; i++) {
    float exponent = 2 * ((float)i / (float)head_dim);
    float theta = powf(base, -exponent);
    for (int pos = 0; pos < seq_len; pos++) {
      cos_values[(size_t)pos * half + i] = cosf(pos * theta);
      sin_values[(size_t)pos * half + i] = sinf(pos * theta);
    }
  }
}

/* ------------------------------------------------------------------ embedding
 * src/embedded_matrix.cu:5-17 */
void orc_embedding(orc_bf16* out, const orc_bf16* table, const int* ids, size_t hidden, size_t n_tok) {
  for (size_t t = 0; t < n_tok; t++)
    memcpy(out + t * hidden, table + (size_t)ids[t] * hidden, hidden * sizeof(orc_bf16));
}

/* ------------------------------------------------------------------ rmsNorm
 * src/normalization.cu:5-26.  One thread per token; sum is a sequential fp32 chain,
 * contracted by nvcc to FFMA (sum = fma(x,x,sum)); eps 1e-4 (:9); (x/rms)*w (:23). */
void orc_rmsnorm(const orc_bf16* x, const orc_bf16* w, orc_bf16* y, size_t hidden, size_t n_tok) {
  orc_rmsnorm_eps(x, w, y, hidden, n_tok, 1e-04f);
}
void orc_rmsnorm_eps(const orc_bf16* x, const orc_bf16* w, orc_bf16* y, size_t hidden, size_t n_tok, float e) {
  for (size_t t = 0; t < n_tok; t++) {
    float sum = 0;
    for (size_t i = 0; i < hidden; i++) {
      float v = orc_bf2f(x[t * hidden + i]);
      sum = fmaf(v, v, sum);
    }
    float rms = sqrtf((sum / (float)hidden) + e);
    for (size_t i = 0; i < hidden; i++) {
      float v = orc_bf2f(x[t * hidden + i]);
      y[t * hidden + i] = orc_f2bf((v / rms) * orc_bf2f(w[i]));
    }
  }
}

/* ------------------------------------------------------------------ matmul
 * src/matrix_mul.cu:165-288.  C[m,k] = sum_n A[m,n]*B[k,n]; the reference walks n in
 * 16-wide chunks (:206) and feeds each chunk to one wmma mma_sync (:256) accumulating
 * in fp32; the result is rounded to bf16 once (:272).  CPU restatement: plain
 * sequential fp32 accumulation in n order (product rounded, then added -- the HMMA
 * datapath keeps products exact, which a CPU float cannot; the difference is below
 * 1 fp32 ulp per step and is absorbed by the bf16 tolerance in the tests). */
typedef struct {
  const orc_bf16 *A, *B;
  orc_bf16* C;
  int M, N, K, k0, k1;
} mm_job;

static void mm_rows(const mm_job* j) {
  int N = j->N;
  float* a = (float*)malloc(sizeof(float) * (size_t)N);
  for (int m = 0; m < j->M; m++) {
    for (int n = 0; n < N; n++) a[n] = orc_bf2f(j->A[(size_t)m * N + n]);
    for (int k = j->k0; k < j->k1; k++) {
      const orc_bf16* b = j->B + (size_t)k * N;
      float acc = 0.f;
      for (int n = 0; n < N; n++) acc += a[n] * orc_bf2f(b[n]);
      j->C[(size_t)m * j->K + k] = orc_f2bf(acc);
    }
  }
  free(a);
}

static void* mm_thread(void* p) {
  mm_rows((const mm_job*)p);
  return NULL;
}

void orc_matmul(const orc_bf16* A, const orc_bf16* B, orc_bf16* C, int M, int N, int K) {
  int nt = g_threads;
  if (nt > K) nt = K;
  if (nt <= 1) {
    mm_job j = {A, B, C, M, N, K, 0, K};
    mm_rows(&j);
    return;
  }
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
  mm_job* jobs = (mm_job*)malloc(sizeof(mm_job) * nt);
  for (int i = 0; i < nt; i++) {
    jobs[i] = (mm_job){A, B, C, M, N, K, (int)((long long)K * i / nt), (int)((long long)K * (i + 1) / nt)};
    pthread_create(&th[i], NULL, mm_thread, &jobs[i]);
  }
  for (int i = 0; i < nt; i++) pthread_join(th[i], NULL);
  free(th);
  free(jobs);
}

/* ------------------------------------------------------------------ qkNorm
 * src/qk_norm.cu:43-80.  Block per head, thread per element; squares go through a
 * shared-memory tree: for stride = hd/2 .. 1: buf[t] += buf[t+stride] (t < stride). */
static float tree_sum(float* buf, int n) {
  for (int stride = n / 2; stride > 0; stride >>= 1)
    for (int t = 0; t < stride; t++) buf[t] += buf[t + stride];
  return buf[0];
}

void orc_qknorm(orc_bf16* qk, const orc_bf16* w, int head_dim, int n_tok, int row_dim, int n_heads) {
  orc_qknorm_eps(qk, w, head_dim, n_tok, row_dim, n_heads, 1e-04f);
}
void orc_qknorm_eps(orc_bf16* qk, const orc_bf16* w, int head_dim, int n_tok, int row_dim, int n_heads, float eps) {
  float buf[256];
  for (int h = 0; h < n_heads; h++)
    for (int tok = 0; tok < n_tok; tok++) {
      orc_bf16* p = qk + (size_t)tok * row_dim + (size_t)h * head_dim;
      for (int t = 0; t < head_dim; t++) {
        float v = orc_bf2f(p[t]);
        buf[t] = v * v;
      }
      float rms = sqrtf((tree_sum(buf, head_dim) / (float)head_dim) + eps);
      for (int t = 0; t < head_dim; t++) p[t] = orc_f2bf((orc_bf2f(p[t]) / rms) * orc_bf2f(w[t]));
    }
}

/* ------------------------------------------------------------------ RoPE
 * src/RoPE.cu:6-22.  Interleaved pairs.  SASS (nvcc 12.9, sm_100a, default -fmad):
 *   v1 = fma(x0, c, -(x1*s));   v2 = fma(c, x1, x0*s). */
void orc_rope(const float* cos_values, const float* sin_values, orc_bf16* x, int n_tok, int head_dim,
              int row_dim, int n_heads) {
  int half = head_dim / 2;
  for (int idx = 0; idx < n_tok; idx++)
    for (int h = 0; h < n_heads; h++)
      for (int i = 0; i < head_dim; i += 2) {
        int j = i / 2;
        size_t base = (size_t)idx * row_dim + (size_t)h * head_dim;
        float x0 = orc_bf2f(x[base + i]), x1 = orc_bf2f(x[base + i + 1]);
        float c = cos_values[(size_t)idx * half + j], s = sin_values[(size_t)idx * half + j];
        float v1 = fmaf(x0, c, -(x1 * s));
        float v2 = fmaf(c, x1, x0 * s);
        x[base + i] = orc_f2bf(v1);
        x[base + i + 1] = orc_f2bf(v2);
      }
}

/* "HF-correct" RoPE (transformers' rotate_half): element j of a head pairs with j + hd/2,
 *   out[j] = x[j]*c_j - x[j+hd/2]*s_j ;  out[j+hd/2] = x[j+hd/2]*c_j + x[j]*s_j,   c_j, s_j the same table rows.
 * Not in the reference (its RoPE is the interleaved form above); fp32 math, one bf16 rounding, same fma placement. */
void orc_rope_half(const float* cos_values, const float* sin_values, orc_bf16* x, int n_tok, int head_dim, int row_dim,
                   int n_heads) {
  int half = head_dim / 2;
  for (int idx = 0; idx < n_tok; idx++)
    for (int h = 0; h < n_heads; h++)
      for (int j = 0; j < half; j++) {
        size_t base = (size_t)idx * row_dim + (size_t)h * head_dim;
        float x0 = orc_bf2f(x[base + j]), x1 = orc_bf2f(x[base + j + half]);
        float c = cos_values[(size_t)idx * half + j], s = sin_values[(size_t)idx * half + j];
        float v1 = fmaf(x0, c, -(x1 * s));
        float v2 = fmaf(c, x1, x0 * s);
        x[base + j] = orc_f2bf(v1);
        x[base + j + half] = orc_f2bf(v2);
      }
}
/* projection bias (Qwen2.5 q/k/v_proj.bias; the reference has none): y = bf16(float(y) + float(b)) on the already
 * rounded projection output, per row */
void orc_add_bias(orc_bf16* y, const orc_bf16* b, int n_tok, int dim) {
  for (int t = 0; t < n_tok; t++)
    for (int i = 0; i < dim; i++) y[(size_t)t * dim + i] = orc_f2bf(orc_bf2f(y[(size_t)t * dim + i]) + orc_bf2f(b[i]));
}

/* ------------------------------------------------------------------ SiLU, mul, add
 * src/SiLU.cu:6-23: y = x * (1/(1+expf(-x))) ; src/element_add.cu:4-13 ;
 * src/residual_add.cu:7-18 */
void orc_silu(orc_bf16* x, size_t n) {
  for (size_t i = 0; i < n; i++) {
    float v = orc_bf2f(x[i]);
    float sg = 1 / (1 + expf(-v));
    x[i] = orc_f2bf(v * sg);
  }
}

void orc_elem_mul(const orc_bf16* a, const orc_bf16* b, orc_bf16* c, size_t n) {
  for (size_t i = 0; i < n; i++) c[i] = orc_f2bf(orc_bf2f(a[i]) * orc_bf2f(b[i]));
}

void orc_residual_add(orc_bf16* a, const orc_bf16* b, size_t n) {
  for (size_t i = 0; i < n; i++) a[i] = orc_f2bf(orc_bf2f(a[i]) + orc_bf2f(b[i]));
}

/* ------------------------------------------------------------------ paged KV */
orc_kv* orc_kv_new(int page_size, int n_layers, int kv_dim) {
  orc_kv* kv = (orc_kv*)calloc(1, sizeof(orc_kv));
  kv->page_size = page_size;
  kv->n_layers = n_layers;
  kv->kv_dim = kv_dim;
  return kv;
}

void orc_kv_free(orc_kv* kv) {
  if (!kv) return;
  for (int i = 0; i < kv->n_pages; i++) {
    free(kv->k_pages[i]);
    free(kv->v_pages[i]);
  }
  free(kv->k_pages);
  free(kv->v_pages);
  free(kv);
}

static void kv_ensure(orc_kv* kv, int page) {
  while (kv->n_pages <= page) {
    size_t elems = (size_t)kv->page_size * kv->n_layers * kv->kv_dim; /* iengine.cu:352 */
    kv->k_pages = (orc_bf16**)realloc(kv->k_pages, sizeof(void*) * (kv->n_pages + 1));
    kv->v_pages = (orc_bf16**)realloc(kv->v_pages, sizeof(void*) * (kv->n_pages + 1));
    kv->k_pages[kv->n_pages] = (orc_bf16*)calloc(elems, sizeof(orc_bf16));
    kv->v_pages[kv->n_pages] = (orc_bf16*)calloc(elems, sizeof(orc_bf16));
    kv->n_pages++;
  }
}

void orc_kv_store(orc_kv* kv, int layer, int pos0, int n_tok, const orc_bf16* K, const orc_bf16* V) {
  for (int t = 0; t < n_tok; t++) {
    int pos = pos0 + t, page = pos / kv->page_size, off = pos % kv->page_size;
    kv_ensure(kv, page);
    size_t o = ((size_t)off * kv->n_layers + layer) * kv->kv_dim; /* include_cuda.cu:263-264 */
    memcpy(kv->k_pages[page] + o, K + (size_t)t * kv->kv_dim, kv->kv_dim * sizeof(orc_bf16));
    memcpy(kv->v_pages[page] + o, V + (size_t)t * kv->kv_dim, kv->kv_dim * sizeof(orc_bf16));
  }
}

/* ------------------------------------------------------------------ attention
 * src/self_attension.cu:10-149.  Per q head, per q token: score[k] = tree(q*k)/sqrtf(hd)
 * (:63-74); causal mask -1e9 (:84-89); thread 0: max, e=expf(s-max), sequential sum,
 * e/=sum (:94-107); out[d] = sequential fma over k of p[k]*v[k][d] (:112-137).
 * GQA: kv head = q head / (n_q/n_kv) (the literal 5 at :33,:116). */
void orc_attention(const orc_bf16* Q, orc_bf16* out, size_t seq_len_q, size_t seq_len_kv, size_t head_dim,
                   size_t q_dim, size_t kv_dim, int n_q_heads, int n_kv_heads, int causal,
                   size_t q_abs_base, int layer, const orc_kv* kv) {
  int group = n_q_heads / n_kv_heads;
  float* score = (float*)malloc(sizeof(float) * (seq_len_kv ? seq_len_kv : 1));
  float buf[256];
  for (int h = 0; h < n_q_heads; h++) {
    int kvh = h / group;
    for (size_t qt = 0; qt < seq_len_q; qt++) {
      const orc_bf16* q = Q + qt * q_dim + (size_t)h * head_dim;
      for (size_t kt = 0; kt < seq_len_kv; kt++) {
        int page = (int)(kt / kv->page_size), off = (int)(kt % kv->page_size);
        const orc_bf16* kp =
            kv->k_pages[page] + ((size_t)off * kv->n_layers + layer) * kv_dim + (size_t)kvh * head_dim;
        for (size_t d = 0; d < head_dim; d++) buf[d] = orc_bf2f(q[d]) * orc_bf2f(kp[d]);
        score[kt] = tree_sum(buf, (int)head_dim) / sqrtf((float)head_dim);
      }
      size_t q_abs = q_abs_base + qt;
      if (causal)
        for (size_t kt = 0; kt < seq_len_kv; kt++)
          if (kt > q_abs) score[kt] = -1e9f;
      float max_val = -1e9f;
      for (size_t kt = 0; kt < seq_len_kv; kt++) max_val = fmaxf(max_val, score[kt]);
      float sum_val = 0.f;
      for (size_t kt = 0; kt < seq_len_kv; kt++) {
        score[kt] = expf(score[kt] - max_val);
        sum_val += score[kt];
      }
      for (size_t kt = 0; kt < seq_len_kv; kt++) score[kt] /= sum_val;
      for (size_t d = 0; d < head_dim; d++) {
        float o = 0.f;
        for (size_t kt = 0; kt < seq_len_kv; kt++) {
          int page = (int)(kt / kv->page_size), off = (int)(kt % kv->page_size);
          const orc_bf16* vp =
              kv->v_pages[page] + ((size_t)off * kv->n_layers + layer) * kv_dim + (size_t)kvh * head_dim;
          o = fmaf(score[kt], orc_bf2f(vp[d]), o); /* out_val += p*v  -> FFMA */
        }
        out[qt * q_dim + (size_t)h * head_dim + d] = orc_f2bf(o);
      }
    }
  }
  free(score);
}

/* ------------------------------------------------------------------ sampling
 * src/logit_decode.cu.  blockArgMax (:19-33): s[tid]=better(s[tid], s[tid+stride]) with
 * better(a,b) = (a.val > b.val) ? a : b  (:15-17) -- ties go to the UPPER partner. */
typedef struct {
  float val;
  int idx;
} pair_t;

static pair_t block_argmax_256(pair_t* s) {
  for (int stride = 128; stride > 0; stride >>= 1)
    for (int tid = 0; tid < stride; tid++) /* tid+stride < 256 always */
      s[tid] = (s[tid].val > s[tid + stride].val) ? s[tid] : s[tid + stride];
  return s[0];
}

/* XORWOW exactly as curand_kernel.h:772-798 (init, subsequence 0, offset 0) and
 * :863-874 (curand) ; curand_uniform.h:69-72. */
typedef struct {
  uint32_t d, v[5];
} xorwow_t;

static void xorwow_init(xorwow_t* st, uint64_t seed) {
  uint32_t s0 = ((uint32_t)seed) ^ 0xaad26b49u;
  uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
  uint32_t t0 = 1099087573u * s0;
  uint32_t t1 = 2591861531u * s1;
  st->d = 6615241u + t1 + t0;
  st->v[0] = 123456789u + t0;
  st->v[1] = 362436069u ^ t0;
  st->v[2] = 521288629u + t1;
  st->v[3] = 88675123u ^ t1;
  st->v[4] = 5783321u + t0;
}

static uint32_t xorwow_next(xorwow_t* st) {
  uint32_t t = (st->v[0] ^ (st->v[0] >> 2));
  st->v[0] = st->v[1];
  st->v[1] = st->v[2];
  st->v[2] = st->v[3];
  st->v[3] = st->v[4];
  st->v[4] = (st->v[4] ^ (st->v[4] << 4)) ^ (t ^ (t << 1));
  st->d += 362437u;
  return st->v[4] + st->d;
}

int orc_sample_topk(const orc_bf16* logits, float temperature, int k, size_t vocab, uint64_t seed,
                    uint64_t subseq) {
  enum { BS = 256 };
  if (k <= 0) return -1;
  if ((size_t)k > vocab) k = (int)vocab;
  if (k > BS) k = BS;
  if (!(temperature > 0.0f)) temperature = 1.0f;
  if (subseq != 0) {
    fprintf(stderr, "orc_sample_topk: only subsequence 0 is restated (the reference never uses another)\n");
    abort();
  }
  float topk_vals[BS];
  int topk_idxs[BS];
  for (int i = 0; i < BS; i++) {
    topk_vals[i] = -INFINITY;
    topk_idxs[i] = -1;
  }
  pair_t s[BS];
  for (int sel = 0; sel < k; sel++) {
    for (int tid = 0; tid < BS; tid++) {
      pair_t local = {-INFINITY, -1};
      for (size_t idx = tid; idx < vocab; idx += BS) {
        float v = orc_bf2f(logits[idx]);
        int chosen = 0;
        for (int t = 0; t < sel; t++)
          if ((int)idx == topk_idxs[t]) {
            chosen = 1;
            break;
          }
        if (!chosen && v > local.val) {
          local.val = v;
          local.idx = (int)idx;
        }
      }
      s[tid] = local;
    }
    pair_t g = block_argmax_256(s);
    if (g.idx == -1) {
      for (int i = sel; i < k; i++) {
        topk_vals[i] = -INFINITY;
        topk_idxs[i] = -1;
      }
    } else {
      topk_vals[sel] = g.val;
      topk_idxs[sel] = g.idx;
    }
    if (topk_idxs[sel] == -1) break;
  }
  int actual_k = 0;
  for (int i = 0; i < k; i++) {
    if (topk_idxs[i] != -1)
      actual_k++;
    else
      break;
  }
  if (actual_k == 0) return -1;
  float max_val = topk_vals[0] / temperature;
  for (int i = 1; i < actual_k; i++) {
    float v = topk_vals[i] / temperature;
    if (v > max_val) max_val = v;
    topk_vals[i] = v;
  }
  topk_vals[0] = topk_vals[0] / temperature;
  float sum = 0.0f;
  for (int i = 0; i < actual_k; i++) {
    topk_vals[i] = expf(topk_vals[i] - max_val);
    sum += topk_vals[i];
  }
  xorwow_t rng;
  xorwow_init(&rng, seed);
  float u = ((float)xorwow_next(&rng) * 2.3283064e-10f + (2.3283064e-10f / 2.0f)) * sum;
  float cum = 0.0f;
  int picked = topk_idxs[actual_k - 1];
  for (int i = 0; i < actual_k; i++) {
    cum += topk_vals[i];
    if (u <= cum) {
      picked = topk_idxs[i];
      break;
    }
  }
  return picked;
}

/* Closed form of the k=1 case: thread tid keeps the LOWEST idx == tid (mod 256) among
 * its maxima (strict >, :198); the reduction prefers the upper partner on ties, and the
 * LAST round (stride 1) decides bit 0 of tid, so bit 0 is the most significant
 * preference bit: winner = max value, then largest bit-reversed (idx & 255), then
 * lowest idx.  -inf / NaN entries are never selected (v > -inf is false). */
static unsigned brev8(unsigned x) {
  x = ((x & 0xF0u) >> 4) | ((x & 0x0Fu) << 4);
  x = ((x & 0xCCu) >> 2) | ((x & 0x33u) << 2);
  x = ((x & 0xAAu) >> 1) | ((x & 0x55u) << 1);
  return x;
}

int orc_argmax_ref_tiebreak(const orc_bf16* logits, size_t vocab) {
  float best = -INFINITY;
  int bi = -1;
  for (size_t i = 0; i < vocab; i++) {
    float v = orc_bf2f(logits[i]);
    if (!(v > -INFINITY)) continue;
    if (bi < 0 || v > best) {
      best = v;
      bi = (int)i;
    } else if (v == best) {
      unsigned a = brev8((unsigned)i & 255u), b = brev8((unsigned)bi & 255u);
      if (a > b) bi = (int)i; /* same class keeps the lower idx (we scan ascending) */
    }
  }
  return bi;
}

/* apply_repetition_penalty_kernel: DECLARED in the reference (include/layers_include.cuh:33: logits, context_tokens,
 * context_len, vocab_size, penalty) but never defined or launched -- there is no reference behaviour to match, so this
 * states the conventional one (CTRL / HF RepetitionPenaltyLogitsProcessor): every DISTINCT token id of the context has
 * its logit divided by the penalty when positive, multiplied when not; fp32 IEEE arithmetic on the bf16 logit, one
 * bf16 rounding.  Ids outside [0, vocab) are ignored. */
void orc_repetition_penalty(orc_bf16* logits, const int* context_tokens, size_t context_len, int vocab, float penalty) {
  for (size_t i = 0; i < context_len; i++) {
    int t = context_tokens[i];
    if (t < 0 || t >= vocab) continue;
    int seen = 0;
    for (size_t j = 0; j < i && !seen; j++) seen = context_tokens[j] == t;
    if (seen) continue;
    float v = orc_bf2f(logits[t]);
    v = v > 0.0f ? v / penalty : v * penalty;
    logits[t] = orc_f2bf(v);
  }
}

/* ------------------------------------------------------------------ checkpoint */
typedef struct {
  char name[160];
  char short_name[96];
  int layer;
  size_t shape[4];
  int ndim;
  size_t begin, end;
} orc_tensor;

struct orc_model {
  orc_config cfg;
  orc_tensor* t;
  int nt;
  const uint8_t* blob;
  size_t blob_bytes;
  int fd;
  float *cos_t, *sin_t;
  /* semantics switches (orc_model_set_semantics): the reference's are eps 1e-4 + interleaved RoPE */
  float eps;
  int rope_half;
};

static const orc_tensor* find_tensor(const orc_model* m, const char* short_name, int layer) {
  for (int i = 0; i < m->nt; i++)
    if (strcmp(m->t[i].short_name, short_name) == 0 && (m->t[i].layer < 0 || m->t[i].layer == layer))
      return &m->t[i];
  return NULL;
}

const orc_bf16* orc_model_tensor(const orc_model* m, const char* short_name, int layer, size_t* n_elems) {
  const orc_tensor* t = find_tensor(m, short_name, layer);
  if (!t) return NULL;
  if (n_elems) *n_elems = (t->end - t->begin) / 2;
  return (const orc_bf16*)(m->blob + t->begin);
}

const orc_config* orc_model_config(const orc_model* m) { return &m->cfg; }
void orc_model_set_semantics(orc_model* m, int rope_half, float eps) {
  m->rope_half = rope_half;
  m->eps = eps > 0 ? eps : 1e-04f;
}

orc_model* orc_model_load(const char* meta_path, const char* weights_path, int head_dim_hint, int context) {
  FILE* f = fopen(meta_path, "r");
  if (!f) return NULL;
  orc_model* m = (orc_model*)calloc(1, sizeof(orc_model));
  m->fd = -1;
  int cap = 0;
  char line[512];
  orc_tensor cur;
  int have = 0;
  memset(&cur, 0, sizeof(cur));
  while (fgets(line, sizeof(line), f)) {
    if (strncmp(line, "Tensor: ", 8) == 0) {
      memset(&cur, 0, sizeof(cur));
      cur.layer = -1;
      sscanf(line + 8, "%159s", cur.name);
      have = 1;
    } else if (have && strstr(line, "layer:")) {
      sscanf(strstr(line, "layer:") + 6, "%d", &cur.layer);
    } else if (have && strstr(line, "short_name:")) {
      sscanf(strstr(line, "short_name:") + 11, "%95s", cur.short_name);
    } else if (have && strstr(line, "shape:")) {
      char* p = strchr(line, '[');
      cur.ndim = 0;
      if (p) {
        p++;
        while (*p && *p != ']') {
          while (*p == ' ') p++;
          if (*p == ']' || !*p) break;
          cur.shape[cur.ndim++] = strtoull(p, &p, 10);
          if (cur.ndim == 4) break;
        }
      }
    } else if (have && strstr(line, "offsets:")) {
      char* p = strchr(line, '[');
      if (p) {
        cur.begin = strtoull(p + 1, &p, 10);
        while (*p == ',' || *p == ' ') p++;
        cur.end = strtoull(p, &p, 10);
      }
      if (m->nt == cap) {
        cap = cap ? cap * 2 : 256;
        m->t = (orc_tensor*)realloc(m->t, sizeof(orc_tensor) * cap);
      }
      m->t[m->nt++] = cur;
      have = 0;
    }
  }
  fclose(f);

  /* config from shapes */
  const orc_tensor* emb = find_tensor(m, "embed_tokens.weight", 0);
  const orc_tensor* lm = find_tensor(m, "logits", 0);
  const orc_tensor* up = find_tensor(m, "mlp.up_proj.weight", 0);
  const orc_tensor* q = find_tensor(m, "self_attn.q_proj.weight", 0);
  const orc_tensor* k = find_tensor(m, "self_attn.k_proj.weight", 0);
  const orc_tensor* qn = find_tensor(m, "self_attn.q_norm.weight", 0);
  if (!emb || !lm || !up || !q || !k) {
    orc_model_free(m);
    return NULL;
  }
  int layers = 0;
  for (int i = 0; i < m->nt; i++)
    if (m->t[i].layer + 1 > layers) layers = m->t[i].layer + 1;
  m->cfg.hidden = (int)emb->shape[1];
  m->cfg.vocab = (int)lm->shape[0];
  m->cfg.inter = (int)up->shape[0];
  m->cfg.layers = layers;
  m->cfg.head_dim = qn ? (int)qn->shape[0] : head_dim_hint;
  m->cfg.n_q = (int)q->shape[0] / m->cfg.head_dim;
  m->cfg.n_kv = (int)k->shape[0] / m->cfg.head_dim;
  m->cfg.context = context;

  m->fd = open(weights_path, O_RDONLY);
  if (m->fd < 0) {
    orc_model_free(m);
    return NULL;
  }
  struct stat st;
  fstat(m->fd, &st);
  m->blob_bytes = (size_t)st.st_size;
  m->blob = (const uint8_t*)mmap(NULL, m->blob_bytes, PROT_READ, MAP_PRIVATE, m->fd, 0);
  if (m->blob == MAP_FAILED) {
    m->blob = NULL;
    orc_model_free(m);
    return NULL;
  }
  size_t n = (size_t)context * (m->cfg.head_dim / 2);
  m->cos_t = (float*)malloc(n * sizeof(float));
  m->sin_t = (float*)malloc(n * sizeof(float));
  orc_precompute_cos_sin(m->cos_t, m->sin_t, context, m->cfg.head_dim);
  return m;
}

void orc_model_free(orc_model* m) {
  if (!m) return;
  if (m->blob) munmap((void*)m->blob, m->blob_bytes);
  if (m->fd >= 0) close(m->fd);
  free(m->cos_t);
  free(m->sin_t);
  free(m->t);
  free(m);
}

/* ------------------------------------------------------------------ sequence + llm() */
struct orc_seq {
  const orc_model* m;
  orc_kv* kv;
  int sequence_len; /* ModelBuffers::sequence_len */
  int cap_tok;
  orc_bf16 *x, *rms_out, *Q, *K, *V, *att, *o, *up, *gate, *gate_out, *down, *logits;
  orc_dump_fn dump;
  void* dump_user;
};

static void seq_reserve(orc_seq* s, int n_tok) {
  if (n_tok <= s->cap_tok) return;
  const orc_config* c = &s->m->cfg;
  size_t H = c->hidden, Dq = (size_t)c->n_q * c->head_dim, Dkv = (size_t)c->n_kv * c->head_dim, I = c->inter;
#define RS(p, n) p = (orc_bf16*)realloc(p, sizeof(orc_bf16) * (n) * (size_t)n_tok)
  RS(s->x, H);
  RS(s->rms_out, H);
  RS(s->Q, Dq);
  RS(s->K, Dkv);
  RS(s->V, Dkv);
  RS(s->att, Dq);
  RS(s->o, H);
  RS(s->up, I);
  RS(s->gate, I);
  RS(s->gate_out, I);
  RS(s->down, H);
#undef RS
  s->cap_tok = n_tok;
}

orc_seq* orc_seq_new(const orc_model* m, int page_size) {
  orc_seq* s = (orc_seq*)calloc(1, sizeof(orc_seq));
  s->m = m;
  s->kv = orc_kv_new(page_size, m->cfg.layers, m->cfg.n_kv * m->cfg.head_dim);
  s->logits = (orc_bf16*)malloc(sizeof(orc_bf16) * (size_t)m->cfg.vocab);
  return s;
}

void orc_seq_free(orc_seq* s) {
  if (!s) return;
  orc_kv_free(s->kv);
  free(s->x);
  free(s->rms_out);
  free(s->Q);
  free(s->K);
  free(s->V);
  free(s->att);
  free(s->o);
  free(s->up);
  free(s->gate);
  free(s->gate_out);
  free(s->down);
  free(s->logits);
  free(s);
}

int orc_seq_len(const orc_seq* s) { return s->sequence_len; }
void orc_set_dump(orc_seq* s, orc_dump_fn fn, void* user) {
  s->dump = fn;
  s->dump_user = user;
}

#define DUMP(tag, layer, ptr, n) \
  if (s->dump) s->dump(s->dump_user, tag, layer, ptr, n)

/* Layer body shared by both branches of llm(): qwen_main.cu:77-222 (prefill) and
 * :271-365 (decode) issue the same 18 kernels; only m, the RoPE position base and the
 * attention arguments differ. */
static void layer_stack(orc_seq* s, int n_tok, int pos0, int causal) {
  const orc_model* m = s->m;
  const orc_config* c = &m->cfg;
  int H = c->hidden, hd = c->head_dim, Dq = c->n_q * hd, Dkv = c->n_kv * hd, I = c->inter;
  int half = hd / 2;
  for (int i = 0; i < c->layers; i++) {
    const float eps = m->eps > 0 ? m->eps : 1e-04f;
    orc_rmsnorm_eps(s->x, orc_model_tensor(m, "input_layernorm.weight", i, NULL), s->rms_out, H, n_tok, eps);
    DUMP("input_norm", i, s->rms_out, (size_t)n_tok * H);
    orc_matmul(s->rms_out, orc_model_tensor(m, "self_attn.q_proj.weight", i, NULL), s->Q, n_tok, H, Dq);
    orc_matmul(s->rms_out, orc_model_tensor(m, "self_attn.k_proj.weight", i, NULL), s->K, n_tok, H, Dkv);
    orc_matmul(s->rms_out, orc_model_tensor(m, "self_attn.v_proj.weight", i, NULL), s->V, n_tok, H, Dkv);
    const orc_bf16* qn = orc_model_tensor(m, "self_attn.q_norm.weight", i, NULL);
    const orc_bf16* kn = orc_model_tensor(m, "self_attn.k_norm.weight", i, NULL);
    const orc_bf16* qb = orc_model_tensor(m, "self_attn.q_proj.bias", i, NULL);
    const orc_bf16* kb = orc_model_tensor(m, "self_attn.k_proj.bias", i, NULL);
    const orc_bf16* vb = orc_model_tensor(m, "self_attn.v_proj.bias", i, NULL);
    if (qb) orc_add_bias(s->Q, qb, n_tok, Dq);
    if (kb) orc_add_bias(s->K, kb, n_tok, Dkv);
    if (vb) orc_add_bias(s->V, vb, n_tok, Dkv);
    if (qn) orc_qknorm_eps(s->Q, qn, hd, n_tok, Dq, c->n_q, eps);
    if (kn) orc_qknorm_eps(s->K, kn, hd, n_tok, Dkv, c->n_kv, eps);
    if (m->rope_half) {
      orc_rope_half(m->cos_t + (size_t)pos0 * half, m->sin_t + (size_t)pos0 * half, s->Q, n_tok, hd, Dq, c->n_q);
      orc_rope_half(m->cos_t + (size_t)pos0 * half, m->sin_t + (size_t)pos0 * half, s->K, n_tok, hd, Dkv, c->n_kv);
    } else {
      orc_rope(m->cos_t + (size_t)pos0 * half, m->sin_t + (size_t)pos0 * half, s->Q, n_tok, hd, Dq, c->n_q);
      orc_rope(m->cos_t + (size_t)pos0 * half, m->sin_t + (size_t)pos0 * half, s->K, n_tok, hd, Dkv, c->n_kv);
    }
    DUMP("q", i, s->Q, (size_t)n_tok * Dq);
    DUMP("k", i, s->K, (size_t)n_tok * Dkv);
    DUMP("v", i, s->V, (size_t)n_tok * Dkv);
    orc_kv_store(s->kv, i, pos0, n_tok, s->K, s->V);
    orc_attention(s->Q, s->att, n_tok, s->sequence_len, hd, Dq, Dkv, c->n_q, c->n_kv, causal, pos0, i, s->kv);
    DUMP("attn", i, s->att, (size_t)n_tok * Dq);
    orc_matmul(s->att, orc_model_tensor(m, "self_attn.o_proj.weight", i, NULL), s->o, n_tok, Dq, H);
    orc_residual_add(s->x, s->o, (size_t)n_tok * H);
    DUMP("x_attn", i, s->x, (size_t)n_tok * H);
    orc_rmsnorm_eps(s->x, orc_model_tensor(m, "post_attention_layernorm.weight", i, NULL), s->rms_out, H, n_tok, eps);
    orc_matmul(s->rms_out, orc_model_tensor(m, "mlp.up_proj.weight", i, NULL), s->up, n_tok, H, I);
    orc_matmul(s->rms_out, orc_model_tensor(m, "mlp.gate_proj.weight", i, NULL), s->gate, n_tok, H, I);
    orc_silu(s->gate, (size_t)n_tok * I);
    orc_elem_mul(s->up, s->gate, s->gate_out, (size_t)n_tok * I);
    DUMP("mlp_h", i, s->gate_out, (size_t)n_tok * I);
    orc_matmul(s->gate_out, orc_model_tensor(m, "mlp.down_proj.weight", i, NULL), s->down, n_tok, I, H);
    orc_residual_add(s->x, s->down, (size_t)n_tok * H);
    DUMP("x_out", i, s->x, (size_t)n_tok * H);
  }
}

static int head_and_sample(orc_seq* s, int row, int topk, float temperature, uint64_t seed,
                           orc_bf16* logits_out) {
  const orc_model* m = s->m;
  const orc_config* c = &m->cfg;
  int H = c->hidden;
  /* final norm is applied to every row in the reference (qwen_main.cu:227); only `row`
   * is consumed (:233 copy_last_vocab_vec / :368 copy_first_token). */
  orc_rmsnorm_eps(s->x + (size_t)row * H, orc_model_tensor(m, "norm.weight", 0, NULL), s->rms_out, H, 1, m->eps > 0 ? m->eps : 1e-04f);
  orc_matmul(s->rms_out, orc_model_tensor(m, "logits", 0, NULL), s->logits, 1, H, c->vocab);
  if (logits_out) memcpy(logits_out, s->logits, sizeof(orc_bf16) * (size_t)c->vocab);
  DUMP("logits", -1, s->logits, (size_t)c->vocab);
  return orc_sample_topk(s->logits, temperature, topk, (size_t)c->vocab, seed, 0);
}

int orc_prefill(orc_seq* s, const int* ids, int n_tok, int topk, float temperature, uint64_t seed,
                orc_bf16* logits_out) {
  const orc_config* c = &s->m->cfg;
  seq_reserve(s, n_tok);
  s->sequence_len = n_tok; /* utills.cu:18 */
  orc_embedding(s->x, orc_model_tensor(s->m, "embed_tokens.weight", 0, NULL), ids, c->hidden, n_tok);
  layer_stack(s, n_tok, 0, /*causal=*/1);
  return head_and_sample(s, n_tok - 1, topk, temperature, seed, logits_out);
}

int orc_decode(orc_seq* s, int token, int topk, float temperature, uint64_t seed, orc_bf16* logits_out) {
  const orc_config* c = &s->m->cfg;
  seq_reserve(s, 1);
  s->sequence_len += 1; /* qwen_main.cu:265 */
  orc_embedding(s->x, orc_model_tensor(s->m, "embed_tokens.weight", 0, NULL), &token, c->hidden, 1);
  layer_stack(s, 1, s->sequence_len - 1, /*causal=*/0);
  return head_and_sample(s, 0, topk, temperature, seed, logits_out);
}

/* ------------------------------------------------------------------ synthetic weights
 * splitmix64-style counter hash -> 8 x 16-bit uniforms summed (Irwin-Hall, sigma =
 * 65536*sqrt(8/12)); integer arithmetic + one exact int->float conversion + one fp32
 * multiply (+ one add for norm vectors): bit-identical on CPU, numpy and GPU. */
static uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

orc_bf16 orc_synth_value(uint64_t seed, uint64_t g, int kind) {
  uint64_t a = mix64(seed + (g + 1) * 0x9E3779B97F4A7C15ull);
  uint64_t b = mix64(a + 0x9E3779B97F4A7C15ull);
  int32_t sum = (int32_t)((a & 0xffff) + ((a >> 16) & 0xffff) + ((a >> 32) & 0xffff) + (a >> 48) +
                          (b & 0xffff) + ((b >> 16) & 0xffff) + ((b >> 32) & 0xffff) + (b >> 48));
  float c = (float)(sum - 262140);
  if (kind == 0) return orc_f2bf(c * (0.02f / 53509.92f));
  return orc_f2bf(1.0f + c * (0.05f / 53509.92f));
}

typedef struct {
  char name[160], short_name[96];
  int layer, ndim, kind;
  size_t shape[2];
} synth_t;

static int synth_cmp(const void* a, const void* b) {
  return strcmp(((const synth_t*)a)->name, ((const synth_t*)b)->name);
}

int orc_synth_write(const orc_config* c, uint64_t seed, const char* meta_path, const char* weights_path) {
  int per_layer = 11, nt = c->layers * per_layer + 3;
  synth_t* t = (synth_t*)calloc(nt, sizeof(synth_t));
  int n = 0;
  size_t H = c->hidden, hd = c->head_dim, Dq = (size_t)c->n_q * hd, Dkv = (size_t)c->n_kv * hd, I = c->inter;
#define ADD(fullname, sname, lyr, d0, d1, nd, kd)           \
  do {                                                      \
    snprintf(t[n].name, sizeof(t[n].name), "%s", fullname); \
    snprintf(t[n].short_name, sizeof(t[n].short_name), "%s", sname); \
    t[n].layer = lyr;                                       \
    t[n].shape[0] = d0;                                     \
    t[n].shape[1] = d1;                                     \
    t[n].ndim = nd;                                         \
    t[n].kind = kd;                                         \
    n++;                                                    \
  } while (0)
  ADD("lm_head.weight", "logits", -1, (size_t)c->vocab, H, 2, 0);
  ADD("model.embed_tokens.weight", "embed_tokens.weight", -1, (size_t)c->vocab, H, 2, 0);
  ADD("model.norm.weight", "norm.weight", -1, H, 0, 1, 1);
  static const struct {
    const char* s;
    int which;
  } names[11] = {{"input_layernorm.weight", 0},        {"mlp.down_proj.weight", 1},
                 {"mlp.gate_proj.weight", 2},          {"mlp.up_proj.weight", 3},
                 {"post_attention_layernorm.weight", 4}, {"self_attn.k_norm.weight", 5},
                 {"self_attn.k_proj.weight", 6},       {"self_attn.o_proj.weight", 7},
                 {"self_attn.q_norm.weight", 8},       {"self_attn.q_proj.weight", 9},
                 {"self_attn.v_proj.weight", 10}};
  for (int l = 0; l < c->layers; l++)
    for (int j = 0; j < 11; j++) {
      char full[160];
      snprintf(full, sizeof(full), "model.layers.%d.%s", l, names[j].s);
      switch (names[j].which) {
        case 0: case 4: ADD(full, names[j].s, l, H, 0, 1, 1); break;
        case 1: ADD(full, names[j].s, l, H, I, 2, 0); break;
        case 2: case 3: ADD(full, names[j].s, l, I, H, 2, 0); break;
        case 5: case 8: ADD(full, names[j].s, l, hd, 0, 1, 1); break;
        case 6: case 10: ADD(full, names[j].s, l, Dkv, H, 2, 0); break;
        case 7: ADD(full, names[j].s, l, H, Dq, 2, 0); break;
        case 9: ADD(full, names[j].s, l, Dq, H, 2, 0); break;
      }
    }
#undef ADD
  /* nlohmann::json objects iterate in byte-lexicographic key order (tensor_parser.cpp:71) */
  qsort(t, n, sizeof(synth_t), synth_cmp);
  FILE* fm = fopen(meta_path, "w");
  FILE* fw = fopen(weights_path, "wb");
  if (!fm || !fw) {
    if (fm) fclose(fm);
    if (fw) fclose(fw);
    free(t);
    return -1;
  }
  size_t off = 0;
  enum { CH = 1 << 16 };
  orc_bf16* buf = (orc_bf16*)malloc(sizeof(orc_bf16) * CH);
  for (int i = 0; i < n; i++) {
    size_t elems = t[i].shape[0] * (t[i].ndim == 2 ? t[i].shape[1] : 1);
    /* exact text of operator<< (tensor_parser.cpp:19-28) + the extra "\n" at :125 */
    fprintf(fm, "Tensor: %s\n  layer: %d\n  short_name: %s\n  shape: [ ", t[i].name, t[i].layer, t[i].short_name);
    for (int d = 0; d < t[i].ndim; d++) fprintf(fm, "%zu ", t[i].shape[d]);
    fprintf(fm, "]\n  offsets: [ %zu, %zu ]\n\n", off, off + elems * 2);
    uint64_t g0 = off / 2;
    for (size_t e = 0; e < elems; e += CH) {
      size_t m = elems - e < CH ? elems - e : CH;
      for (size_t j = 0; j < m; j++) buf[j] = orc_synth_value(seed, g0 + e + j, t[i].kind);
      fwrite(buf, sizeof(orc_bf16), m, fw);
    }
    off += elems * 2;
  }
  free(buf);
  fclose(fm);
  fclose(fw);
  free(t);
  return 0;
}
