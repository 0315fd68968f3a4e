// ref_math.cuh -- device functions that reproduce the reference's arithmetic op for op
// (citations relative to /root/reference/layers/).  Shared by the per-operator kernels
// (ops_ref_order.cu, gemm_ref_order.cu) and the persistent decode kernel (decode_mega.cu).
// Translation units including this header are compiled with -fmad=false: every fused
// multiply-add the reference's SASS contains is spelled __fmaf_rn, everything else is an
// explicit _rn intrinsic.
#pragma once
#include <math_constants.h>

#include "common.cuh"

namespace qie {

// ------------------------------------------------------------------ per-head tree sum
// Lane l holds, for p in [0,NP): elements t = 64p + 2l and 64p + 2l + 1 of a head of
// hd = 64*NP values.  Reproduces  for (stride = hd/2; stride > 0; stride >>= 1)
//   buf[t] += buf[t+stride]  (t < stride)   exactly; the result is valid in lane 0.
template <int NP>
__device__ __forceinline__ float head_tree_sum(float (&v)[NP][2]) {
#pragma unroll
  for (int s = NP / 2; s >= 1; s >>= 1)  // strides >= 64: partners live in the same lane
#pragma unroll
    for (int p = 0; p < s; ++p) {
      v[p][0] = __fadd_rn(v[p][0], v[p + s][0]);
      v[p][1] = __fadd_rn(v[p][1], v[p + s][1]);
    }
  float a = v[0][0], b = v[0][1];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {  // strides 32,16,8,4,2
    a = __fadd_rn(a, __shfl_down_sync(0xffffffffu, a, o));
    b = __fadd_rn(b, __shfl_down_sync(0xffffffffu, b, o));
  }
  return __fadd_rn(a, b);  // stride 1
}

// q/k-norm of one head held in registers (in place, values become the bf16-rounded
// results as floats): qk_norm.cu:46-78.
template <int NP>
__device__ __forceinline__ void head_norm(float (&x)[NP][2], const bf16* __restrict__ w, int lane, float eps = 1e-04f) {
  float sq[NP][2];
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    sq[p][0] = __fmul_rn(x[p][0], x[p][0]);
    sq[p][1] = __fmul_rn(x[p][1], x[p][1]);
  }
  float tot = head_tree_sum<NP>(sq);
  float rms = __fsqrt_rn(__fadd_rn(__fdiv_rn(tot, (float)(64 * NP)), eps));
  rms = __shfl_sync(0xffffffffu, rms, 0);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    uint32_t wp = *reinterpret_cast<const uint32_t*>(w + 64 * p + 2 * lane);
    x[p][0] = bf2f(f2bf(__fmul_rn(__fdiv_rn(x[p][0], rms), lo2f(wp))));
    x[p][1] = bf2f(f2bf(__fmul_rn(__fdiv_rn(x[p][1], rms), hi2f(wp))));
  }
}

// RoPE of one head held in registers: RoPE.cu:15-18 (pair j = 32p + lane).
template <int NP>
__device__ __forceinline__ void head_rope(float (&x)[NP][2], const float* __restrict__ cos_row,
                                          const float* __restrict__ sin_row, int lane) {
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    float c = cos_row[32 * p + lane], s = sin_row[32 * p + lane];
    float x0 = x[p][0], x1 = x[p][1];
    float v1 = __fmaf_rn(x0, c, -__fmul_rn(x1, s));
    float v2 = __fmaf_rn(c, x1, __fmul_rn(x0, s));
    x[p][0] = bf2f(f2bf(v1));
    x[p][1] = bf2f(f2bf(v2));
  }
}

// "HF-correct" RoPE (transformers' rotate_half; Qwen2.5 / Qwen3 checkpoints): element j of a head
// pairs with j + hd/2 and both use table column j.  Lane l holds elements 64p + 2l, 64p + 2l + 1: for hd = 128 the
// partner (+64) is the same lane's p = 1 value; for hd = 64 the partner (+32) lives in lane l ^ 16, same slot.
template <int NP>
__device__ __forceinline__ void head_rope_half(float (&x)[NP][2], const float* __restrict__ cos_row,
                                               const float* __restrict__ sin_row, int lane) {
  if (NP == 1) {
    const bool hi = lane >= 16;  // this lane holds elements of the second half
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int j = (2 * lane + s) & 31;  // table column of this element's pair
      const float c = cos_row[j], sn = sin_row[j];
      const float mine = x[0][s], other = __shfl_xor_sync(0xffffffffu, x[0][s], 16);
      // first half: x0*c - x1*s ; second half: x1*c + x0*s  (x0 = first-half element, x1 = its partner)
      const float v = hi ? __fmaf_rn(c, mine, __fmul_rn(other, sn)) : __fmaf_rn(mine, c, -__fmul_rn(other, sn));
      x[0][s] = bf2f(f2bf(v));
    }
  } else {
#pragma unroll
    for (int p = 0; p < NP / 2; ++p)
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int j = 64 * p + 2 * lane + s;
        const float c = cos_row[j], sn = sin_row[j];
        const float x0 = x[p][s], x1 = x[p + NP / 2][s];
        x[p][s] = bf2f(f2bf(__fmaf_rn(x0, c, -__fmul_rn(x1, sn))));
        x[p + NP / 2][s] = bf2f(f2bf(__fmaf_rn(c, x1, __fmul_rn(x0, sn))));
      }
  }
}
// projection bias (Qwen2.5): y = bf16(y + b) on the rounded projection output
template <int NP>
__device__ __forceinline__ void head_add_bias(float (&x)[NP][2], const bf16* __restrict__ b, int lane) {
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const uint32_t bp = *reinterpret_cast<const uint32_t*>(b + 64 * p + 2 * lane);
    x[p][0] = bf2f(f2bf(__fadd_rn(x[p][0], lo2f(bp))));
    x[p][1] = bf2f(f2bf(__fadd_rn(x[p][1], hi2f(bp))));
  }
}

template <int NP>
__device__ __forceinline__ void head_load(float (&x)[NP][2], const bf16* __restrict__ src, int lane) {
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    uint32_t v = *reinterpret_cast<const uint32_t*>(src + 64 * p + 2 * lane);
    x[p][0] = lo2f(v);
    x[p][1] = hi2f(v);
  }
}
template <int NP>
__device__ __forceinline__ void head_store(const float (&x)[NP][2], bf16* __restrict__ dst, int lane) {
#pragma unroll
  for (int p = 0; p < NP; ++p)
    *reinterpret_cast<uint32_t*>(dst + 64 * p + 2 * lane) = pack2(f2bf(x[p][0]), f2bf(x[p][1]));
}

// SiLU.cu:6-8,19-20:  y = x * (1 / (1 + expf(-x)))
__device__ __forceinline__ float silu_ref(float x) {
  float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
  return __fmul_rn(x, sg);
}


// arg-max candidate order of the reference sampler (logit_decode.cu:15-33,182-223; SURVEY 8a
// S1): value, then the larger bit-reversed low byte of the index, then the lower index.
struct Cand {
  float val;
  int idx;
};
__device__ __forceinline__ bool cand_better(float va, int ia, float vb, int ib) {
  if (ib < 0) return ia >= 0;
  if (ia < 0) return false;
  if (va > vb) return true;
  if (va < vb) return false;
  unsigned ra = __brev((unsigned)ia << 24), rb = __brev((unsigned)ib << 24);
  if (ra != rb) return ra > rb;
  return ia < ib;
}


}  // namespace qie
