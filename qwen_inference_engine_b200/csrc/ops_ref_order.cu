// ops_ref_order.cu -- the non-GEMM operators of the forward, written so that every
// floating-point result is produced by the same sequence of IEEE operations as the
// reference kernel it replaces (citations relative to /root/reference/layers/), while
// the work is spread over warps/blocks instead of one thread per token:
//   rmsNorm        src/normalization.cu:5-26   sequential FFMA chain (kept: it IS the spec)
//   qkNorm         src/qk_norm.cu:43-80        shared-memory tree  -> same tree in shuffles
//   RoPE           src/RoPE.cu:6-22            fma(x0,c,-(x1*s)) , fma(c,x1,x0*s)
//   selfattention  src/self_attension.cu:10-149
//   sampling       src/logit_decode.cu:149-274
// This file is compiled with -fmad=false; every fused multiply-add the reference's SASS
// contains is spelled __fmaf_rn here, everything else is an explicit _rn intrinsic.
#include <curand_kernel.h>
#include <math_constants.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"
#include "ref_math.cuh"

namespace qie {

// ------------------------------------------------------------------ embedding
__global__ void embedding_kernel(bf16* __restrict__ out, const bf16* __restrict__ table,
                                 const int* __restrict__ ids, size_t hidden) {
  pdl_wait();
  pdl_trigger();
  const size_t t = blockIdx.x;
  // a sampler that found no finite logit reports -1 (never a valid id); such a row reads row 0 instead of out of bounds
  const bf16* src = table + (size_t)max(ids[t], 0) * hidden;
  bf16* dst = out + t * hidden;
  if ((hidden & 7) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(dst);
    for (size_t i = threadIdx.x; i < hidden / 8; i += blockDim.x) d4[i] = s4[i];
  } else {
    for (size_t i = threadIdx.x; i < hidden; i += blockDim.x) dst[i] = src[i];
  }
}

cudaError_t launch_embedding(bf16* out, const bf16* table, const int* ids, size_t hidden, size_t n_tok,
                             cudaStream_t st) {
  if (n_tok == 0) return cudaSuccess;
  (void)launch_k(embedding_kernel, dim3((unsigned)n_tok), dim3(128), 0, st, out, table, ids, hidden);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ rmsNorm
// One block per token.  The reference's sum is a strictly sequential fp32 FFMA chain
// over the hidden dimension (one thread per token); fp32 addition is not associative, so
// the chain order is part of the result.  Thread 0 walks the chain out of shared memory
// (4-cycle dependent FFMA issue, loads hoisted), the block does the rest in parallel.
__global__ void __launch_bounds__(128) rmsnorm_ref_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                           bf16* __restrict__ y, int hidden, size_t x_stride, float eps) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ float xs[];
  __shared__ float s_rms;
  const bf16* xr = x + (size_t)blockIdx.x * x_stride;
  bf16* yr = y + (size_t)blockIdx.x * hidden;
  for (int i = threadIdx.x; i < hidden; i += blockDim.x) xs[i] = bf2f(xr[i]);
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
    int i = 0;
    for (; i + 8 <= hidden; i += 8) {
      float v0 = xs[i], v1 = xs[i + 1], v2 = xs[i + 2], v3 = xs[i + 3];
      float v4 = xs[i + 4], v5 = xs[i + 5], v6 = xs[i + 6], v7 = xs[i + 7];
      sum = __fmaf_rn(v0, v0, sum);
      sum = __fmaf_rn(v1, v1, sum);
      sum = __fmaf_rn(v2, v2, sum);
      sum = __fmaf_rn(v3, v3, sum);
      sum = __fmaf_rn(v4, v4, sum);
      sum = __fmaf_rn(v5, v5, sum);
      sum = __fmaf_rn(v6, v6, sum);
      sum = __fmaf_rn(v7, v7, sum);
    }
    for (; i < hidden; ++i) sum = __fmaf_rn(xs[i], xs[i], sum);
    s_rms = __fsqrt_rn(__fadd_rn(__fdiv_rn(sum, (float)hidden), eps));  // eps 1e-4 in the reference (normalization.cu:9)
  }
  __syncthreads();
  const float rms = s_rms;
  for (int i = threadIdx.x; i < hidden; i += blockDim.x)
    yr[i] = f2bf(__fmul_rn(__fdiv_rn(xs[i], rms), bf2f(w[i])));
}

cudaError_t launch_rmsnorm_ref(const bf16* x, const bf16* w, bf16* y, size_t hidden, size_t n_tok, size_t x_stride,
                               cudaStream_t st, float eps) {
  if (n_tok == 0) return cudaSuccess;
  (void)launch_k(rmsnorm_ref_kernel, dim3((unsigned)n_tok), dim3(128), hidden * sizeof(float), st, x, w, y, (int)hidden, x_stride, eps);
  return cudaGetLastError();
}

// fast-numerics variant: the sum of squares is reduced in parallel (warp shuffles), so it
// differs from the reference's sequential chain in the last fp32 bits.
__global__ void __launch_bounds__(256) rmsnorm_fast_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                            bf16* __restrict__ y, int hidden, size_t x_stride, float eps) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_part[8];
  const bf16* xr = x + (size_t)blockIdx.x * x_stride;
  bf16* yr = y + (size_t)blockIdx.x * hidden;
  float sum = 0.f;
  for (int i = threadIdx.x * 2; i < hidden; i += 512) {
    uint32_t v = *reinterpret_cast<const uint32_t*>(xr + i);
    float a = lo2f(v), b = hi2f(v);
    sum += a * a + b * b;
  }
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = sum;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += s_part[i];
  const float rms = __fsqrt_rn(__fadd_rn(__fdiv_rn(tot, (float)hidden), eps));
  for (int i = threadIdx.x * 2; i < hidden; i += 512) {
    uint32_t v = *reinterpret_cast<const uint32_t*>(xr + i), wv = *reinterpret_cast<const uint32_t*>(w + i);
    *reinterpret_cast<uint32_t*>(yr + i) = pack2(f2bf(__fmul_rn(__fdiv_rn(lo2f(v), rms), lo2f(wv))),
                                                 f2bf(__fmul_rn(__fdiv_rn(hi2f(v), rms), hi2f(wv))));
  }
}
// one WARP per row, 16-byte accesses, no block barrier -- the block-per-row kernel above ran at a quarter of the HBM
// rate on the 4096 rows of a prefill (15.9 us for 25 MB: latency of one tiny block per row)
__global__ void __launch_bounds__(256) rmsnorm_fast_rows_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                                 bf16* __restrict__ y, int hidden, size_t x_stride, float eps,
                                                                 int n_tok) {
  pdl_wait();
  pdl_trigger();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n_tok) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * x_stride);
  const uint4* wr = reinterpret_cast<const uint4*>(w);
  uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * hidden);
  const int nv = hidden >> 3;
  float sum = 0.f;
  for (int i = lane; i < nv; i += 32) {
    const uint4 v = xr[i];
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float a = lo2f(u[j]), b = hi2f(u[j]);
      sum += a * a + b * b;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float rms = __fsqrt_rn(__fadd_rn(__fdiv_rn(sum, (float)hidden), eps));
  for (int i = lane; i < nv; i += 32) {
    const uint4 v = xr[i], wv = wr[i];  // (the row is still in L1)
    const uint32_t u[4] = {v.x, v.y, v.z, v.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
    uint32_t o4[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      o4[j] = pack2(f2bf(__fmul_rn(__fdiv_rn(lo2f(u[j]), rms), lo2f(ww[j]))), f2bf(__fmul_rn(__fdiv_rn(hi2f(u[j]), rms), hi2f(ww[j]))));
    yr[i] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
  }
}

cudaError_t launch_rmsnorm_fast(const bf16* x, const bf16* w, bf16* y, size_t hidden, size_t n_tok, size_t x_stride,
                                cudaStream_t st, float eps) {
  if (n_tok == 0) return cudaSuccess;
  if (hidden & 1) return cudaErrorInvalidValue;
  // (every row count takes this kernel when the layout allows: a row's result must not depend on how a prompt is chunked)
  if (hidden % 8 == 0 && x_stride % 8 == 0 &&
      ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    (void)launch_k(rmsnorm_fast_rows_kernel, dim3((unsigned)((n_tok + 7) / 8)), dim3(256), 0, st, x, w, y, (int)hidden, x_stride, eps,
                   (int)n_tok);
    return cudaGetLastError();
  }
  (void)launch_k(rmsnorm_fast_kernel, dim3((unsigned)n_tok), dim3(256), 0, st, x, w, y, (int)hidden, x_stride, eps);
  return cudaGetLastError();
}

// ---- operator-level qkNorm / RoPE (one warp per (token, head)) ----------------------
template <int NP>
__global__ void qknorm_ref_kernel(bf16* __restrict__ x, const bf16* __restrict__ w, int n_tok, int row_dim,
                                  int n_heads) {
  pdl_wait();
  pdl_trigger();
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n_tok * n_heads) return;
  int tok = gw / n_heads, h = gw % n_heads;
  bf16* p = x + (size_t)tok * row_dim + (size_t)h * 64 * NP;
  float v[NP][2];
  head_load<NP>(v, p, lane);
  head_norm<NP>(v, w, lane);
  head_store<NP>(v, p, lane);
}

template <int NP>
__global__ void rope_ref_kernel(const float* __restrict__ cos_t, const float* __restrict__ sin_t,
                                bf16* __restrict__ x, int n_tok, const int* __restrict__ pos, int pos0, int row_dim,
                                int n_heads) {
  pdl_wait();
  pdl_trigger();
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= n_tok * n_heads) return;
  int tok = gw / n_heads, h = gw % n_heads;
  int ps = pos ? pos[tok] : pos0 + tok;
  bf16* p = x + (size_t)tok * row_dim + (size_t)h * 64 * NP;
  float v[NP][2];
  head_load<NP>(v, p, lane);
  head_rope<NP>(v, cos_t + (size_t)ps * 32 * NP, sin_t + (size_t)ps * 32 * NP, lane);
  head_store<NP>(v, p, lane);
}

#define QIE_DISPATCH_HD(hd, CALL)                  \
  switch (hd) {                                    \
    case 64: { constexpr int NP = 1; CALL; break; }  \
    case 128: { constexpr int NP = 2; CALL; break; } \
    case 256: { constexpr int NP = 4; CALL; break; } \
    default: return cudaErrorInvalidValue;         \
  }

cudaError_t launch_qknorm_ref(bf16* x, const bf16* w, int hd, int n_tok, int row_dim, int n_heads, cudaStream_t st) {
  int warps = n_tok * n_heads;
  if (warps == 0) return cudaSuccess;
  int blocks = (warps + 3) / 4;
  QIE_DISPATCH_HD(hd, ((void)launch_k(qknorm_ref_kernel<NP>, dim3(blocks), dim3(128), 0, st, x, w, n_tok, row_dim, n_heads)));
  return cudaGetLastError();
}

cudaError_t launch_rope_ref(const float* cos_t, const float* sin_t, bf16* x, int n_tok, const int* pos, int pos0,
                            int hd, int row_dim, int n_heads, cudaStream_t st) {
  int warps = n_tok * n_heads;
  if (warps == 0) return cudaSuccess;
  int blocks = (warps + 3) / 4;
  QIE_DISPATCH_HD(hd,
                  ((void)launch_k(rope_ref_kernel<NP>, dim3(blocks), dim3(128), 0, st, cos_t, sin_t, x, n_tok, pos, pos0, row_dim, n_heads)));
  return cudaGetLastError();
}

// ---- fused q/k-norm + RoPE + KV store: replaces 2x qkNorm + 2x RoPE + the
// cudaMemcpy2D scatter of kv_copy_layer_to_cache_* (include_cuda.cu:165-279) ------------
// One warp = up to QP_HEADS consecutive heads (q heads, then k heads, then v heads) of one token: the token's position,
// page and RoPE row are looked up once, the heads' rows are requested together and processed one after the other (r02:
// one warp per (token, head) ran a chain of three dependent global loads per 256 bytes of payload -- 23.9 us for 33 MB
// on a 4096-token prefill).  The arithmetic per head is unchanged (head_norm / head_rope).
constexpr int QP_HEADS = 4;
template <int NP>
__global__ void qkv_post_kernel(QkvPostArgs a) {
  pdl_wait();
  pdl_trigger();
  const int hd = 64 * NP;
  const int heads = a.n_q + 2 * a.kv.n_kv;  // q heads, k heads, v heads
  const int groups = (heads + QP_HEADS - 1) / QP_HEADS;
  int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= a.n_tok * groups) return;
  const int tok = gw / groups, h0 = (gw - tok * groups) * QP_HEADS;
  const int nh = min(QP_HEADS, heads - h0);
  const int ps = a.pos[tok];
  const float* cos_row = a.cos_t + (size_t)ps * 32 * NP;
  const float* sin_row = a.sin_t + (size_t)ps * 32 * NP;
  int page = 0, off = 0;
  if (h0 + nh > a.n_q) {  // the group holds k / v heads
    page = a.block_table[(size_t)a.slot[tok] * a.max_pages + ps / a.kv.page_size];
    off = ps % a.kv.page_size;
  }
  auto src_of = [&](int h) -> const bf16* {
    if (h < a.n_q) return a.q_in + (size_t)tok * a.q_in_stride + (size_t)h * hd;
    if (h < a.n_q + a.kv.n_kv) return a.k + (size_t)tok * a.kv_stride + (size_t)(h - a.n_q) * hd;
    return a.v + (size_t)tok * a.kv_stride + (size_t)(h - a.n_q - a.kv.n_kv) * hd;
  };
  float v[QP_HEADS][NP][2];
#pragma unroll
  for (int j = 0; j < QP_HEADS; ++j)
    if (j < nh) head_load<NP>(v[j], src_of(h0 + j), lane);
#pragma unroll
  for (int j = 0; j < QP_HEADS; ++j) {
    if (j >= nh) break;
    const int h = h0 + j;
    if (h < a.n_q) {
      if (a.q_bias) head_add_bias<NP>(v[j], a.q_bias + (size_t)h * hd, lane);
      if (a.q_norm_w) head_norm<NP>(v[j], a.q_norm_w, lane, a.eps);
      if (a.rope_half)
        head_rope_half<NP>(v[j], cos_row, sin_row, lane);
      else
        head_rope<NP>(v[j], cos_row, sin_row, lane);
      head_store<NP>(v[j], a.q + (size_t)tok * a.n_q * hd + (size_t)h * hd, lane);
    } else if (h < a.n_q + a.kv.n_kv) {
      const int kh = h - a.n_q;
      if (a.k_bias) head_add_bias<NP>(v[j], a.k_bias + (size_t)kh * hd, lane);
      if (a.k_norm_w) head_norm<NP>(v[j], a.k_norm_w, lane, a.eps);
      if (a.rope_half)
        head_rope_half<NP>(v[j], cos_row, sin_row, lane);
      else
        head_rope<NP>(v[j], cos_row, sin_row, lane);
      head_store<NP>(v[j], a.kv.chunk(page, a.layer, 0, kh) + (size_t)off * hd, lane);
    } else {
      const int vh = h - a.n_q - a.kv.n_kv;
      if (a.v_bias) head_add_bias<NP>(v[j], a.v_bias + (size_t)vh * hd, lane);
      head_store<NP>(v[j], a.kv.chunk(page, a.layer, 1, vh) + (size_t)off * hd, lane);  // (bf16 -> fp32 -> bf16 is the identity)
    }
  }
}

cudaError_t launch_qkv_post(const QkvPostArgs& a, cudaStream_t st) {
  const int heads = a.n_q + 2 * a.kv.n_kv;
  int warps = a.n_tok * ((heads + QP_HEADS - 1) / QP_HEADS);
  if (warps == 0) return cudaSuccess;
  int blocks = (warps + 3) / 4;
  QIE_DISPATCH_HD(a.kv.hd, ((void)launch_k(qkv_post_kernel<NP>, dim3(blocks), dim3(128), 0, st, a)));
  return cudaGetLastError();
}

// plain K/V scatter (operator-level twin of kv_copy_layer_to_cache_*)
__global__ void kv_store_kernel(KvGeom kv, int layer, const bf16* __restrict__ K, const bf16* __restrict__ V,
                                const int* __restrict__ pos, const int* __restrict__ slot,
                                const int* __restrict__ block_table, int max_pages, int n_tok) {
  pdl_wait();
  pdl_trigger();
  int tok = blockIdx.x;
  int ps = pos[tok];
  int page = block_table[(size_t)slot[tok] * max_pages + ps / kv.page_size];
  int off = ps % kv.page_size;
  int dkv = kv.n_kv * kv.hd;
  for (int i = threadIdx.x; i < dkv; i += blockDim.x) {
    int h = i / kv.hd, d = i % kv.hd;
    kv.chunk(page, layer, 0, h)[(size_t)off * kv.hd + d] = K[(size_t)tok * dkv + i];
    kv.chunk(page, layer, 1, h)[(size_t)off * kv.hd + d] = V[(size_t)tok * dkv + i];
  }
}

cudaError_t launch_kv_store(const KvGeom& kv, int layer, const bf16* K, const bf16* V, const int* pos,
                            const int* slot, const int* block_table, int max_pages, int n_tok, cudaStream_t st) {
  if (n_tok == 0) return cudaSuccess;
  (void)launch_k(kv_store_kernel, dim3(n_tok), dim3(128), 0, st, kv, layer, K, V, pos, slot, block_table, max_pages, n_tok);
  return cudaGetLastError();
}

// K/V rows -> a page list in the reference's layout (kv_copy_layer_to_cache_{prefill,decode}, include_cuda.cu:165-279:
// one cudaMemcpy2D pair per page per layer there; one launch here)
__global__ void kv_store_pagelist_kernel(bf16* const* __restrict__ k_pages, bf16* const* __restrict__ v_pages, int page_size,
                                         int n_layers, int layer, int kv_dim, const bf16* __restrict__ K,
                                         const bf16* __restrict__ V, int pos0, int n_tok) {
  pdl_wait();
  pdl_trigger();
  const int tok = blockIdx.x, pos = pos0 + tok;
  const size_t off = ((size_t)(pos % page_size) * n_layers + layer) * kv_dim;
  bf16* kd = k_pages[pos / page_size] + off;
  bf16* vd = v_pages[pos / page_size] + off;
  for (int i = threadIdx.x; i < kv_dim; i += blockDim.x) {
    kd[i] = K[(size_t)tok * kv_dim + i];
    vd[i] = V[(size_t)tok * kv_dim + i];
  }
}
cudaError_t launch_kv_store_pagelist(bf16* const* k_pages, bf16* const* v_pages, int page_size, int n_layers, int layer, int kv_dim,
                                     const bf16* K, const bf16* V, int pos0, int n_tok, cudaStream_t st) {
  if (n_tok == 0) return cudaSuccess;
  (void)launch_k(kv_store_pagelist_kernel, dim3(n_tok), dim3(128), 0, st, k_pages, v_pages, page_size, n_layers, layer, kv_dim, K, V, pos0, n_tok);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ elementwise
__global__ void silu_kernel(bf16* x, size_t n) {
  pdl_wait();
  pdl_trigger();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = bf2f(x[i]);
    float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-v)));
    x[i] = f2bf(__fmul_rn(v, sg));
  }
}
__global__ void elem_mul_kernel(const bf16* a, const bf16* b, bf16* c, size_t n) {
  pdl_wait();
  pdl_trigger();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) c[i] = f2bf(__fmul_rn(bf2f(a[i]), bf2f(b[i])));
}
__global__ void residual_add_kernel(bf16* a, const bf16* b, size_t n) {
  pdl_wait();
  pdl_trigger();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = f2bf(__fadd_rn(bf2f(a[i]), bf2f(b[i])));
}
cudaError_t launch_silu(bf16* x, size_t n, cudaStream_t st) {
  if (!n) return cudaSuccess;
  (void)launch_k(silu_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, x, n);
  return cudaGetLastError();
}
cudaError_t launch_elem_mul(const bf16* a, const bf16* b, bf16* c, size_t n, cudaStream_t st) {
  if (!n) return cudaSuccess;
  (void)launch_k(elem_mul_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, a, b, c, n);
  return cudaGetLastError();
}
cudaError_t launch_residual_add(bf16* a, const bf16* b, size_t n, cudaStream_t st) {
  if (!n) return cudaSuccess;
  (void)launch_k(residual_add_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, a, b, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ attention
// grid (q head, token); 256 threads.  Phase 1: one warp per kv position computes the
// score with the reference's 2^k tree over head_dim.  Phase 2: max (order free), expf,
// the reference's sequential sum, p = e/sum.  Phase 3: thread d accumulates
// out[d] = fma(p[k], v[k][d], out[d]) for k ascending (self_attension.cu:112-137).
// Positions > pos[t] contribute exact zeros in the reference (mask -1e9 -> expf = 0 ->
// fma(0, v, o) = o), so stopping at pos[t] is bit-identical.
template <int NP, bool PL>  // PL: page-list addressing (the reference's cache layout)
__global__ void __launch_bounds__(256) attention_ref_kernel(AttnArgs a) {
  pdl_wait();
  pdl_trigger();
  constexpr int hd = 64 * NP;
  extern __shared__ float sm[];
  float* score = sm;                                                   // [kv_len]
  int* pages = reinterpret_cast<int*>(sm + a.max_kv_len);              // [pages used]
  __shared__ float s_red[8];
  __shared__ float s_bcast;

  const int h = blockIdx.x, tok = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv_len = a.pos[tok] + 1;
  const int group = a.n_q / a.kv.n_kv;
  const int kvh = h / group;
  const int psz = a.kv.page_size;
  const int n_pages = (kv_len + psz - 1) / psz;
  if (!PL) {
    const int* bt = a.block_table + (size_t)a.slot[tok] * a.max_pages;
    for (int i = threadIdx.x; i < n_pages; i += blockDim.x) pages[i] = bt[i];
  }
  // row of cache position k: pool chunk [slot][hd], or the reference's page [slot][layer][kv_dim]
  const size_t pl_row = (size_t)a.pl_layers * a.kv.n_kv * hd, pl_off = (size_t)a.layer * a.kv.n_kv * hd + (size_t)kvh * hd;
  auto k_row = [&](int k) -> const bf16* {
    return PL ? a.k_pages[k / psz] + (size_t)(k % psz) * pl_row + pl_off : a.kv.chunk(pages[k / psz], a.layer, 0, kvh) + (size_t)(k % psz) * hd;
  };
  auto v_row = [&](int k) -> const bf16* {
    return PL ? a.v_pages[k / psz] + (size_t)(k % psz) * pl_row + pl_off : a.kv.chunk(pages[k / psz], a.layer, 1, kvh) + (size_t)(k % psz) * hd;
  };

  float q[NP][2];
  head_load<NP>(q, a.q + (size_t)tok * a.n_q * hd + (size_t)h * hd, lane);
  __syncthreads();

  const float inv_den = __fsqrt_rn((float)hd);
  for (int k = warp; k < kv_len; k += 8) {
    const bf16* kp = k_row(k);
    float pr[NP][2];
    head_load<NP>(pr, kp, lane);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      pr[p][0] = __fmul_rn(q[p][0], pr[p][0]);
      pr[p][1] = __fmul_rn(q[p][1], pr[p][1]);
    }
    float dot = head_tree_sum<NP>(pr);
    if (lane == 0) score[k] = __fdiv_rn(dot, inv_den);
  }
  __syncthreads();

  // max
  float m = -1e9f;
  for (int k = threadIdx.x; k < kv_len; k += blockDim.x) m = fmaxf(m, score[k]);
  m = warp_max(m);
  if (lane == 0) s_red[warp] = m;
  __syncthreads();
  m = s_red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, s_red[i]);
  for (int k = threadIdx.x; k < kv_len; k += blockDim.x) score[k] = expf(__fsub_rn(score[k], m));
  __syncthreads();
  if (threadIdx.x == 0) {
    float sum = 0.f;
    int k = 0;
    for (; k + 8 <= kv_len; k += 8) {
      float e0 = score[k], e1 = score[k + 1], e2 = score[k + 2], e3 = score[k + 3];
      float e4 = score[k + 4], e5 = score[k + 5], e6 = score[k + 6], e7 = score[k + 7];
      sum = __fadd_rn(sum, e0);
      sum = __fadd_rn(sum, e1);
      sum = __fadd_rn(sum, e2);
      sum = __fadd_rn(sum, e3);
      sum = __fadd_rn(sum, e4);
      sum = __fadd_rn(sum, e5);
      sum = __fadd_rn(sum, e6);
      sum = __fadd_rn(sum, e7);
    }
    for (; k < kv_len; ++k) sum = __fadd_rn(sum, score[k]);
    s_bcast = sum;
  }
  __syncthreads();
  const float sum = s_bcast;
  for (int k = threadIdx.x; k < kv_len; k += blockDim.x) score[k] = __fdiv_rn(score[k], sum);
  __syncthreads();

  if (threadIdx.x < hd) {
    const int d = threadIdx.x;
    float o = 0.f;
    int k = 0;
    for (; k + 4 <= kv_len; k += 4) {
      float v0 = bf2f(v_row(k)[d]);
      float v1 = bf2f(v_row(k + 1)[d]);
      float v2 = bf2f(v_row(k + 2)[d]);
      float v3 = bf2f(v_row(k + 3)[d]);
      o = __fmaf_rn(score[k], v0, o);
      o = __fmaf_rn(score[k + 1], v1, o);
      o = __fmaf_rn(score[k + 2], v2, o);
      o = __fmaf_rn(score[k + 3], v3, o);
    }
    for (; k < kv_len; ++k) o = __fmaf_rn(score[k], bf2f(v_row(k)[d]), o);
    a.out[(size_t)tok * a.n_q * hd + (size_t)h * hd + d] = f2bf(o);
  }
}

cudaError_t launch_attention_ref(const AttnArgs& a, cudaStream_t st) {
  if (a.n_tok == 0) return cudaSuccess;
  int max_pages_used = (a.max_kv_len + a.kv.page_size - 1) / a.kv.page_size;
  size_t smem = (size_t)a.max_kv_len * sizeof(float) + (size_t)max_pages_used * sizeof(int);
  dim3 grid(a.n_q, a.n_tok);
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
#define QIE_ATTN(NPV)                                                                                              \
  {                                                                                                                \
    /* per-device opt-in to > 48 KiB of dynamic shared memory: set on every launch (cheap), not once per process */ \
    if (a.k_pages) {                                                                                               \
      cudaError_t e = cudaFuncSetAttribute(attention_ref_kernel<NPV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (e != cudaSuccess) return e;                                                                              \
      (void)launch_k(attention_ref_kernel<NPV, true>, dim3(grid), dim3(256), smem, st, a);                          \
    } else {                                                                                                       \
      cudaError_t e = cudaFuncSetAttribute(attention_ref_kernel<NPV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); \
      if (e != cudaSuccess) return e;                                                                              \
      (void)launch_k(attention_ref_kernel<NPV, false>, dim3(grid), dim3(256), smem, st, a);                         \
    }                                                                                                              \
  }
  switch (a.kv.hd) {
    case 64: QIE_ATTN(1); break;
    case 128: QIE_ATTN(2); break;
    case 256: QIE_ATTN(4); break;
    default: return cudaErrorInvalidValue;
  }
#undef QIE_ATTN
  return cudaGetLastError();
}

// ------------------------------------------------------------------ sampling
// The reference runs ONE block of 256 threads and, for each of k rounds, a strided scan
// + shared-memory arg-max whose tie-break is a function of (idx & 255):
// value, then the larger bit-reversed low byte of idx, then the lower idx.  That is a
// total order, so the scan can use any number of threads and still pick the same token.
__global__ void __launch_bounds__(1024) sample_topk_kernel(const bf16* __restrict__ logits, int* __restrict__ out,
                                                            size_t vocab, float temperature, int k, uint64_t seed,
                                                            uint64_t seed_stride, const int* __restrict__ step_ptr,
                                                            unsigned long long subsequence) {
  pdl_wait();
  pdl_trigger();
  __shared__ float topk_vals[256];
  __shared__ int topk_idxs[256];
  __shared__ Cand s_c[32];
  const bf16* row = logits + (size_t)blockIdx.x * vocab;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (k <= 0) {
    if (tid == 0) out[blockIdx.x] = -1;
    return;
  }
  k = min(k, (int)min(vocab, (size_t)256));
  if (!(temperature > 0.0f)) temperature = 1.0f;
  if (tid < 256) {
    topk_vals[tid] = -CUDART_INF_F;
    topk_idxs[tid] = -1;
  }
  __syncthreads();
  for (int sel = 0; sel < k; ++sel) {
    float bv = -CUDART_INF_F;
    int bi = -1;
    for (size_t idx = tid; idx < vocab; idx += blockDim.x) {
      float v = bf2f(row[idx]);
      if (!(v > -CUDART_INF_F)) continue;  // -inf / NaN are never selected (v > local.val fails)
      bool chosen = false;
      for (int t = 0; t < sel; ++t)
        if ((int)idx == topk_idxs[t]) {
          chosen = true;
          break;
        }
      if (!chosen && cand_better(v, (int)idx, bv, bi)) {
        bv = v;
        bi = (int)idx;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (cand_better(ov, oi, bv, bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) {
      s_c[warp].val = bv;
      s_c[warp].idx = bi;
    }
    __syncthreads();
    if (warp == 0) {
      bv = lane < (blockDim.x >> 5) ? s_c[lane].val : -CUDART_INF_F;
      bi = lane < (blockDim.x >> 5) ? s_c[lane].idx : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (cand_better(ov, oi, bv, bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (lane == 0) {
        topk_vals[sel] = bi >= 0 ? bv : -CUDART_INF_F;
        topk_idxs[sel] = bi;
      }
    }
    __syncthreads();
    if (topk_idxs[sel] == -1) break;
  }
  if (tid == 0) {
    int actual_k = 0;
    for (int i = 0; i < k; ++i) {
      if (topk_idxs[i] != -1)
        actual_k++;
      else
        break;
    }
    if (actual_k == 0) {
      out[blockIdx.x] = -1;
      return;
    }
    if (actual_k == 1) {  // softmax over one candidate is 1; u in (0,1] always selects it
      out[blockIdx.x] = topk_idxs[0];
      return;
    }
    float max_val = __fdiv_rn(topk_vals[0], temperature);
    for (int i = 1; i < actual_k; ++i) {
      float v = __fdiv_rn(topk_vals[i], temperature);
      if (v > max_val) max_val = v;
      topk_vals[i] = v;
    }
    topk_vals[0] = __fdiv_rn(topk_vals[0], temperature);
    float sum = 0.0f;
    for (int i = 0; i < actual_k; ++i) {
      topk_vals[i] = expf(__fsub_rn(topk_vals[i], max_val));
      sum = __fadd_rn(sum, topk_vals[i]);
    }
    curandState rng;
    unsigned long long sd = seed + (unsigned long long)blockIdx.x * seed_stride + (step_ptr ? (unsigned)step_ptr[blockIdx.x] : 0u);
    curand_init(sd, subsequence, 0, &rng);  // logit_decode.cu:256-257: (seed, subsequence, offset 0)
    float u = __fmul_rn(curand_uniform(&rng), sum);
    float cum = 0.0f;
    int picked = topk_idxs[actual_k - 1];
    for (int i = 0; i < actual_k; ++i) {
      cum = __fadd_rn(cum, topk_vals[i]);
      if (u <= cum) {
        picked = topk_idxs[i];
        break;
      }
    }
    out[blockIdx.x] = picked;
  }
}

// Greedy mode (k == 1, what the persistent kernel folds into its lm_head epilogue): the sampled token is the
// arg-max in the reference's total order whatever the random draw, so the row can be scanned by several CTAs with
// 16-byte loads -- the one-block-per-row scan above is latency bound (85 us for 152 k logits, 64 SMs busy at batch
// 64).  Stage 1: CTA (slice, row) reduces its slice to one candidate; stage 2: one warp per row merges the slices.
constexpr int GREEDY_SLICES = 8;
__global__ void __launch_bounds__(256) greedy_slice_kernel(const bf16* __restrict__ logits, Cand* __restrict__ cands, size_t vocab) {
  pdl_wait();
  pdl_trigger();
  __shared__ Cand s_c[8];
  const int slice = blockIdx.x, rowi = blockIdx.y;
  const bf16* row = logits + (size_t)rowi * vocab;
  const size_t nvec = vocab >> 3;  // 8 logits per 16-byte load (rows are 16-byte aligned: vocab % 8 == 0, checked by the launcher)
  const size_t per = (nvec + GREEDY_SLICES - 1) / GREEDY_SLICES;
  const size_t v0 = (size_t)slice * per, v1 = min(nvec, v0 + per);
  float bv = -CUDART_INF_F;
  int bi = -1;
  for (size_t v = v0 + threadIdx.x; v < v1; v += 256) {
    const uint4 q = ld_nc_v4(reinterpret_cast<const uint4*>(row) + v);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float f0 = lo2f(w[e]), f1 = hi2f(w[e]);
      const int i0 = (int)(v * 8) + 2 * e;
      if (f0 > -CUDART_INF_F && cand_better(f0, i0, bv, bi)) {
        bv = f0;
        bi = i0;
      }
      if (f1 > -CUDART_INF_F && cand_better(f1, i0 + 1, bv, bi)) {
        bv = f1;
        bi = i0 + 1;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (cand_better(ov, oi, bv, bi)) {
      bv = ov;
      bi = oi;
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) s_c[warp] = Cand{bv, bi};
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w2 = 1; w2 < 8; ++w2)
      if (cand_better(s_c[w2].val, s_c[w2].idx, bv, bi)) {
        bv = s_c[w2].val;
        bi = s_c[w2].idx;
      }
    cands[(size_t)rowi * GREEDY_SLICES + slice] = Cand{bv, bi};
  }
}
__global__ void greedy_merge_kernel(const Cand* __restrict__ cands, int* __restrict__ out, int n_rows) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  float bv = -CUDART_INF_F;
  int bi = -1;
  for (int s2 = 0; s2 < GREEDY_SLICES; ++s2) {
    const Cand c = cands[(size_t)r * GREEDY_SLICES + s2];
    if (cand_better(c.val, c.idx, bv, bi)) {
      bv = c.val;
      bi = c.idx;
    }
  }
  out[r] = bi;  // -1 if the row holds no finite logit, like the scan above
}

// Top-k (k > 1) without k scans.  The reference selects by k rounds of "arg-max over everything not chosen yet"
// (logit_decode.cu:182-223), i.e. the k best elements of the row in its total order -- value, then the larger
// bit-reversed low byte of the index, then the lower index (cand_better) -- best first.  That order is the numeric
// order of a 48-bit key [16: order-preserving bf16][8: brev(idx & 255)][24: 0xFFFFFF - idx], all keys distinct, so the
// k-th best key is found by a radix select: six 8-bit digits from the top, one 256-bin shared-memory histogram pass
// over the row per digit (the row is 300 KB and stays in L2), then one pass collects the k keys >= the threshold, a
// bitonic sort orders them, and thread 0 runs the reference's tail (temperature, softmax over k, XORWOW draw, CDF
// scan) unchanged.  7 passes instead of k = 50 (x exclusion-list checks); one block of 1024 threads per row.
__device__ __forceinline__ unsigned long long topk_key(uint16_t bits, int idx) {
  if (bits == 0x8000u) bits = 0;  // -0.0 == +0.0 in the reference's float compares
  const uint32_t v16 = (bits & 0x8000u) ? (uint32_t)(uint16_t)~bits : (uint32_t)(bits | 0x8000u);
  const uint32_t rev = __brev((unsigned)idx << 24);  // brev of the low byte, in the low byte
  return ((unsigned long long)v16 << 32) | ((unsigned long long)rev << 24) | (unsigned long long)(0xFFFFFFu - (unsigned)idx);
}
__global__ void __launch_bounds__(1024) sample_topk_select_kernel(const bf16* __restrict__ logits, int* __restrict__ out, size_t vocab,
                                                                    float temperature, int k, uint64_t seed, uint64_t seed_stride,
                                                                    const int* __restrict__ step_ptr, unsigned long long subsequence) {
  pdl_wait();
  pdl_trigger();
  __shared__ unsigned hist[256];
  __shared__ unsigned long long sel[256];  // the selected keys, then sorted best first
  __shared__ float topk_vals[256];
  __shared__ int topk_idxs[256];
  __shared__ unsigned long long s_prefix;
  __shared__ int s_need, s_count;
  const uint16_t* row = reinterpret_cast<const uint16_t*>(logits) + (size_t)blockIdx.x * vocab;
  const int tid = threadIdx.x;
  if (!(temperature > 0.0f)) temperature = 1.0f;
  k = min(k, (int)min(vocab, (size_t)256));
  auto valid = [](uint16_t b) {  // -inf and NaN are never selected (the reference's v > best fails for both)
    const uint16_t a = b & 0x7fffu;
    return !(a > 0x7f80u) && b != 0xff80u;
  };
  // how many candidates exist at all
  unsigned mine = 0;
  for (size_t i = tid; i < vocab; i += blockDim.x) mine += valid(row[i]) ? 1u : 0u;
  if (tid == 0) s_count = 0;
  __syncthreads();
  mine = __reduce_add_sync(0xffffffffu, mine);
  if ((tid & 31) == 0 && mine) atomicAdd(&s_count, (int)mine);
  __syncthreads();
  const int actual_k = min(k, s_count);
  if (actual_k <= 0) {
    if (tid == 0) out[blockIdx.x] = -1;
    return;
  }
  if (tid == 0) {
    s_prefix = 0ull;
    s_need = actual_k;
  }
  // ---- radix select of the actual_k-th largest key
  for (int d = 5; d >= 0; --d) {
    if (tid < 256) hist[tid] = 0u;
    __syncthreads();
    const unsigned long long prefix = s_prefix;
    const int shift = 8 * d;
    for (size_t i = tid; i < vocab; i += blockDim.x) {
      const uint16_t b = row[i];
      if (!valid(b)) continue;
      const unsigned long long key = topk_key(b, (int)i);
      if (d == 5 || (key >> (shift + 8)) == (prefix >> (shift + 8))) atomicAdd(&hist[(unsigned)(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int need = s_need, dig = 255;
      for (; dig > 0; --dig) {  // digits above the threshold digit are taken whole
        if ((int)hist[dig] >= need) break;
        need -= (int)hist[dig];
      }
      s_prefix = prefix | ((unsigned long long)dig << shift);
      s_need = need;
    }
    __syncthreads();
  }
  // ---- collect the keys >= threshold (exactly actual_k: keys are distinct)
  const unsigned long long thr = s_prefix;
  if (tid == 0) s_count = 0;
  if (tid < 256) sel[tid] = 0ull;
  __syncthreads();
  for (size_t i = tid; i < vocab; i += blockDim.x) {
    const uint16_t b = row[i];
    if (!valid(b)) continue;
    const unsigned long long key = topk_key(b, (int)i);
    if (key >= thr) {
      const int slot = atomicAdd(&s_count, 1);
      if (slot < 256) sel[slot] = key;
    }
  }
  __syncthreads();
  // ---- bitonic sort of 256 keys, descending (empty slots are 0 = below every real key)
  for (int size = 2; size <= 256; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < 128) {
        const int lo = 2 * tid - (tid & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long a = sel[lo], b2 = sel[hi];
        if ((a < b2) == desc) {
          sel[lo] = b2;
          sel[hi] = a;
        }
      }
      __syncthreads();
    }
  if (tid < actual_k) {
    const int idx = (int)(0xFFFFFFu - (unsigned)(sel[tid] & 0xFFFFFFull));
    topk_idxs[tid] = idx;
    topk_vals[tid] = bf2f(logits[(size_t)blockIdx.x * vocab + idx]);
  }
  __syncthreads();
  if (tid == 0) {  // the reference's tail, logit_decode.cu:225-273 (as in sample_topk_kernel)
    if (actual_k == 1) {
      out[blockIdx.x] = topk_idxs[0];
      return;
    }
    float max_val = __fdiv_rn(topk_vals[0], temperature);
    for (int i = 1; i < actual_k; ++i) {
      float v = __fdiv_rn(topk_vals[i], temperature);
      if (v > max_val) max_val = v;
      topk_vals[i] = v;
    }
    topk_vals[0] = __fdiv_rn(topk_vals[0], temperature);
    float sum = 0.0f;
    for (int i = 0; i < actual_k; ++i) {
      topk_vals[i] = expf(__fsub_rn(topk_vals[i], max_val));
      sum = __fadd_rn(sum, topk_vals[i]);
    }
    curandState rng;
    unsigned long long sd = seed + (unsigned long long)blockIdx.x * seed_stride + (step_ptr ? (unsigned)step_ptr[blockIdx.x] : 0u);
    curand_init(sd, subsequence, 0, &rng);
    float u = __fmul_rn(curand_uniform(&rng), sum);
    float cum = 0.0f;
    int picked = topk_idxs[actual_k - 1];
    for (int i = 0; i < actual_k; ++i) {
      cum = __fadd_rn(cum, topk_vals[i]);
      if (u <= cum) {
        picked = topk_idxs[i];
        break;
      }
    }
    out[blockIdx.x] = picked;
  }
}

cudaError_t launch_sample_topk(const bf16* logits, int* out_tokens, int n_rows, size_t vocab, float temperature,
                               int k, uint64_t seed, uint64_t seed_stride, const int* step_ptr, cudaStream_t st,
                               uint64_t subsequence) {
  if (n_rows == 0) return cudaSuccess;
  if (k == 1 && vocab % 8 == 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 && vocab >= 4096) {
    // candidate scratch per stream (engines of one process run on their own streams); buffers are never freed or
    // moved once handed out, because captured CUDA graphs keep their address
    struct Scratch {
      Cand* p = nullptr;
      int cap = 0;
    };
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, std::vector<Scratch>> pool;  // keyed by (device, stream): a stream handle
                                                                               // value can come back on another device
    Cand* cands = nullptr;
    {
      int dev = 0;
      (void)cudaGetDevice(&dev);
      std::lock_guard<std::mutex> lk(mu);
      std::vector<Scratch>& v = pool[std::make_pair(dev, st)];
      for (const Scratch& sc : v)
        if (sc.cap >= n_rows) cands = sc.p;
      if (!cands) {
        Scratch sc;
        sc.cap = std::max(n_rows, 256);
        cudaError_t e = cudaMalloc(&sc.p, (size_t)sc.cap * GREEDY_SLICES * sizeof(Cand));
        if (e != cudaSuccess) return e;
        v.push_back(sc);
        cands = sc.p;
      }
    }
    (void)launch_k(greedy_slice_kernel, dim3(GREEDY_SLICES, n_rows), dim3(256), 0, st, logits, cands, vocab);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    (void)launch_k(greedy_merge_kernel, dim3((n_rows + 63) / 64), dim3(64), 0, st, (const Cand*)cands, out_tokens, n_rows);
    return cudaGetLastError();
  }
  // k > 1: radix-select kernel (7 passes over the row); QIE_TOPK_SCAN=1 keeps the k-round scan of the reference (A/B knob,
  // and the only path for vocabularies beyond 2^24 entries, which the 24-bit index field of the key cannot hold)
  static const bool scan_env = [] { const char* v = getenv("QIE_TOPK_SCAN"); return v && v[0] == '1'; }();
  if (k > 1 && !scan_env && vocab < (1u << 24))
    (void)launch_k(sample_topk_select_kernel, dim3(n_rows), dim3(1024), 0, st, logits, out_tokens, vocab, temperature, k, seed, seed_stride,
                   step_ptr, (unsigned long long)subsequence);
  else
    (void)launch_k(sample_topk_kernel, dim3(n_rows), dim3(1024), 0, st, logits, out_tokens, vocab, temperature, k, seed, seed_stride, step_ptr,
                   (unsigned long long)subsequence);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ repetition penalty
// apply_repetition_penalty_kernel is declared by the reference (include/layers_include.cuh:33) and never defined; the
// conventional semantics (CTRL / HF RepetitionPenaltyLogitsProcessor) are implemented: every DISTINCT token id of
// a row's context has its logit divided (positive) or multiplied (otherwise) by the penalty.  Row r's context is
// ctx[ctx_row[r] * ctx_stride + 0 .. ctx_len[r]) (ctx_row == nullptr: row r, ctx_len == nullptr: fixed_len).
// One thread per context position; only the first occurrence of a token applies, so no logit is written twice.
__global__ void repetition_penalty_kernel(bf16* __restrict__ logits, const int* __restrict__ ctx, const int* __restrict__ ctx_row,
                                          const int* __restrict__ ctx_len, int len_bias, size_t fixed_len, size_t ctx_stride, int vocab,
                                          float penalty) {
  pdl_wait();
  pdl_trigger();
  const int r = blockIdx.y;
  const size_t n = ctx_len ? (size_t)(ctx_len[r] + len_bias) : fixed_len;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int* c = ctx + (size_t)(ctx_row ? ctx_row[r] : r) * ctx_stride;
  const int t = c[i];
  if (t < 0 || t >= vocab) return;
  for (size_t j = 0; j < i; ++j)
    if (c[j] == t) return;
  bf16* p = logits + (size_t)r * vocab + t;
  float v = bf2f(*p);
  v = v > 0.0f ? __fdiv_rn(v, penalty) : __fmul_rn(v, penalty);
  *p = f2bf(v);
}
cudaError_t launch_repetition_penalty(bf16* logits, const int* ctx, const int* ctx_row, const int* ctx_len, int len_bias, size_t fixed_len,
                                      size_t max_len, size_t ctx_stride, int n_rows, int vocab, float penalty, cudaStream_t st) {
  if (n_rows == 0 || max_len == 0) return cudaSuccess;
  (void)launch_k(repetition_penalty_kernel, dim3((unsigned)((max_len + 255) / 256), n_rows), dim3(256), 0, st, logits, ctx, ctx_row, ctx_len,
                 len_bias, fixed_len, ctx_stride, vocab, penalty);
  return cudaGetLastError();
}
// token history of the engine's sequences (only kept while a repetition penalty is set): the input token of every row
// of a forward goes to hist[slot[t]][pos[t]]
__global__ void history_append_kernel(int* __restrict__ hist, size_t stride, const int* __restrict__ ids, const int* __restrict__ pos,
                                      const int* __restrict__ slot, int n) {
  pdl_wait();
  pdl_trigger();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n && (size_t)pos[t] < stride) hist[(size_t)slot[t] * stride + pos[t]] = ids[t];
}
cudaError_t launch_history_append(int* hist, size_t stride, const int* ids, const int* pos, const int* slot, int n, cudaStream_t st) {
  if (!n) return cudaSuccess;
  (void)launch_k(history_append_kernel, dim3((n + 127) / 128), dim3(128), 0, st, hist, stride, ids, pos, slot, n);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ synthetic weights
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__global__ void synth_fill_kernel(bf16* blob, size_t elem_begin, size_t n, unsigned long long seed, int kind) {
  pdl_wait();
  pdl_trigger();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  unsigned long long g = elem_begin + i;
  unsigned long long a = mix64(seed + (g + 1) * 0x9E3779B97F4A7C15ull);
  unsigned long long b = mix64(a + 0x9E3779B97F4A7C15ull);
  int sum = (int)((a & 0xffff) + ((a >> 16) & 0xffff) + ((a >> 32) & 0xffff) + (a >> 48) + (b & 0xffff) +
                  ((b >> 16) & 0xffff) + ((b >> 32) & 0xffff) + (b >> 48));
  float c = (float)(sum - 262140);
  float v = kind == 0 ? __fmul_rn(c, 0.02f / 53509.92f) : __fadd_rn(1.0f, __fmul_rn(c, 0.05f / 53509.92f));
  blob[g] = f2bf(v);
}
cudaError_t launch_synth_fill(bf16* blob, size_t elem_begin, size_t n_elems, uint64_t seed, int kind,
                              cudaStream_t st) {
  if (!n_elems) return cudaSuccess;
  (void)launch_k(synth_fill_kernel, dim3((unsigned)((n_elems + 255) / 256)), dim3(256), 0, st, blob, elem_begin, n_elems, seed, kind);
  return cudaGetLastError();
}

// fill positions [pos0, pos0+n_pos) of one sequence's KV pages (all layers, K and V) with
// seeded N(0,1)-like values: bench-only stand-in for a long prefill
__global__ void kv_fill_kernel(KvGeom kv, const int* __restrict__ block_row, int pos0, int n_pos,
                               unsigned long long seed) {
  pdl_wait();
  pdl_trigger();
  const size_t per_pos = (size_t)kv.n_layers * 2 * kv.n_kv * kv.hd;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_pos * n_pos) return;
  int p = (int)(i / per_pos);
  size_t r = i % per_pos;
  int d = (int)(r % kv.hd);
  r /= kv.hd;
  int head = (int)(r % kv.n_kv);
  r /= kv.n_kv;
  int which = (int)(r % 2);
  int layer = (int)(r / 2);
  int pos = pos0 + p;
  int page = block_row[pos / kv.page_size];
  unsigned long long a = mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ull);
  int sum = (int)((a & 0xffff) + ((a >> 16) & 0xffff) + ((a >> 32) & 0xffff) + (a >> 48));
  float v = __fmul_rn((float)(sum - 131070), 1.0f / 37837.0f);
  kv.chunk(page, layer, which, head)[(size_t)(pos % kv.page_size) * kv.hd + d] = f2bf(v);
}
cudaError_t launch_kv_fill(const KvGeom& kv, const int* block_row, int pos0, int n_pos, uint64_t seed,
                           cudaStream_t st) {
  size_t n = (size_t)kv.n_layers * 2 * kv.n_kv * kv.hd * n_pos;
  if (!n) return cudaSuccess;
  (void)launch_k(kv_fill_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, kv, block_row, pos0, n_pos, seed);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ tensor-parallel arg-max merge
__global__ void tp_cand_make_kernel(const bf16* __restrict__ logits, const int* __restrict__ sampled, TpCand* cand, int n,
                                    size_t vocab_local, int vocab_offset) {
  pdl_wait();
  pdl_trigger();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int li = sampled[i];
  TpCand c;
  c.val = li >= 0 ? bf2f(logits[(size_t)i * vocab_local + li]) : -CUDART_INF_F;
  c.idx = li >= 0 ? li + vocab_offset : -1;
  cand[i] = c;
}
__global__ void tp_cand_merge_kernel(const TpCand* __restrict__ all, int tp, int n, int* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float bv = -CUDART_INF_F;
  int bi = -1;
  for (int r = 0; r < tp; ++r) {
    const TpCand c = all[(size_t)r * n + i];
    if (cand_better(c.val, c.idx, bv, bi)) {
      bv = c.val;
      bi = c.idx;
    }
  }
  out[i] = bi;
}
__global__ void residual_add_f32_kernel(bf16* __restrict__ x, const float* __restrict__ y, size_t n) {
  pdl_wait();
  pdl_trigger();
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = f2bf(__fadd_rn(bf2f(x[i]), bf2f(f2bf(y[i]))));
}
cudaError_t launch_residual_add_f32(bf16* x, const float* y, size_t n, cudaStream_t st) {
  if (!n) return cudaSuccess;
  (void)launch_k(residual_add_f32_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, x, y, n);
  return cudaGetLastError();
}
cudaError_t launch_tp_cand_make(const bf16* logits_local, const int* sampled_local, TpCand* cand, int n_rows, size_t vocab_local,
                                int vocab_offset, cudaStream_t st) {
  if (!n_rows) return cudaSuccess;
  (void)launch_k(tp_cand_make_kernel, dim3((n_rows + 127) / 128), dim3(128), 0, st, logits_local, sampled_local, cand, n_rows,
                 vocab_local, vocab_offset);
  return cudaGetLastError();
}
cudaError_t launch_tp_cand_merge(const TpCand* all, int tp, int n_rows, int* out_tokens, cudaStream_t st) {
  if (!n_rows) return cudaSuccess;
  (void)launch_k(tp_cand_merge_kernel, dim3((n_rows + 127) / 128), dim3(128), 0, st, all, tp, n_rows, out_tokens);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ step bookkeeping
__global__ void advance_kernel(int* pos, int* ids, const int* sampled, int n, int* step_ptr) {
  pdl_wait();
  pdl_trigger();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    pos[i] += 1;
    ids[i] = sampled[i];
    if (step_ptr) step_ptr[i] += 1;
  }
}
cudaError_t launch_advance(int* pos, int* ids, const int* sampled, int n, int* step_ptr, cudaStream_t st) {
  if (!n) return cudaSuccess;
  (void)launch_k(advance_kernel, dim3((n + 127) / 128), dim3(128), 0, st, pos, ids, sampled, n, step_ptr);
  return cudaGetLastError();
}

}  // namespace qie
