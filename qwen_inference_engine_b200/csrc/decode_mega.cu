// decode_mega.cu -- the whole decode step as ONE persistent cooperative kernel (sm_100a).
//
// The reference runs a decode token as 18 kernels + 2-4 memcpys per layer, most of them
// followed by a device sync (llm(), /root/reference/layers/src/qwen_main.cu:250-404).
// Even with 8 fused launches per layer the step stays launch/latency bound
// (profiles/r01a_*).  Here one CTA per SM stays resident for the whole step:
//
//   * a PRODUCER warp walks the static schedule of weight tiles of this CTA for ALL layers
//     and streams them HBM -> shared memory with TMA bulk copies (cp.async.bulk +
//     mbarrier complete_tx) into a ring.  Weights do not depend on activations, so the
//     stream runs ahead across phase boundaries and grid barriers: HBM never idles while
//     the dependent chain of a token (norm -> qkv -> attention -> o -> norm -> gate/up ->
//     down) is resolved.
//   * 8 CONSUMER warps do the math in the reference's order: one warp = 8 weight rows over
//     the whole K through mma.sync.m16n8k16 with k-chunks ascending (matrix_mul.cu:206-259),
//     the sequential RMSNorm FFMA chain (normalization.cu:11-15), the 2^k tree of
//     qkNorm / attention scores, the serial softmax sum and PV chain
//     (self_attension.cu:94-137).  Results are bit-identical to the per-operator kernels
//     and therefore to the reference's kernels.
//   * phases are separated by a grid barrier (one atomic + spin per CTA); activations
//     travel between phases through L2 ([B, *] bf16, a few KB).
//
// Phases per layer: QKV (RMSNorm fused in front) | attention (q/k-norm + RoPE + KV store
// fused in front) | O (+residual) | GATE/UP (RMSNorm in front, SiLU*up behind) | DOWN
// (+residual); then final norm + lm_head with the greedy arg-max (reference tie-break,
// logit_decode.cu:15-33,182-223) folded into the epilogue, and the step bookkeeping.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "ref_math.cuh"

namespace qie {
namespace {

constexpr int NW = 8;                         // consumer warps
constexpr int NTC = NW * 32;                  // consumer threads
constexpr int MEGA_THREADS = NTC + 32;        // + producer warp
constexpr int MAX_SLOTS = 32;
constexpr int MAX_ROWS = 8;                   // rows per launch (A fragments by predicated LDS)
constexpr int OFF_FULL = 0, OFF_EMPTY = 256, OFF_RMS = 512, OFF_CAND = 1024;
constexpr int HDR_BYTES = OFF_CAND + NW * MAX_ROWS * 8;  // 1536

enum { PH_QKV = 0, PH_O = 1, PH_GATEUP = 2, PH_DOWN = 3, PH_LMHEAD = 4 };

// ---------------------------------------------------------------- mbarrier / TMA bulk
__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MB_DONE_%=;\n"
      "bra MB_WAIT_%=;\n"
      "MB_DONE_%=:\n"
      "}\n" ::"r"(addr),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(NTC) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// grid barrier for the consumer warps of all CTAs (cooperative launch: all CTAs resident)
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& epoch) {
  bar_consumers();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
    } while (v < epoch);
    __threadfence();
  }
  bar_consumers();
}

// ---------------------------------------------------------------- GEMM phase description
struct Phase {
  const bf16* W[3];
  const bf16* W2;
  int rows[3];
  int ubeg[4];
  int nseg, K, dual, units, kind;
};

__device__ __forceinline__ void make_phase(const MegaArgs& a, int idx, Phase& p) {
  const int hd = a.hd, Dq = a.n_q * hd, Dkv = a.n_kv * hd;
  p.W2 = nullptr;
  p.dual = 0;
  p.nseg = 1;
  if (idx >= 4 * a.L) {
    p.kind = PH_LMHEAD;
    p.W[0] = a.lm_head;
    p.rows[0] = a.V;
    p.K = a.H;
  } else {
    const MegaLayer& w = a.layers[idx >> 2];
    p.kind = idx & 3;
    switch (p.kind) {
      case PH_QKV:
        p.nseg = 3;
        p.W[0] = w.q; p.W[1] = w.k; p.W[2] = w.v;
        p.rows[0] = Dq; p.rows[1] = Dkv; p.rows[2] = Dkv;
        p.K = a.H;
        break;
      case PH_O:
        p.W[0] = w.o; p.rows[0] = a.H; p.K = Dq;
        break;
      case PH_GATEUP:
        p.W[0] = w.gate; p.W2 = w.up; p.rows[0] = a.I; p.K = a.H; p.dual = 1;
        break;
      default:
        p.W[0] = w.down; p.rows[0] = a.H; p.K = a.I;
        break;
    }
  }
  int u = 0;
  for (int s = 0; s < p.nseg; ++s) {
    p.ubeg[s] = u;
    u += (p.rows[s] + 7) >> 3;
  }
  p.ubeg[p.nseg] = u;
  p.units = u;
}

__device__ __forceinline__ int units_of_cta(int units, int cta, int grid) {
  return cta < units ? (units - cta + grid - 1) / grid : 0;
}

// ---------------------------------------------------------------- producer
// Walks every GEMM phase of the step in order; per phase the units of this CTA in rounds of
// NW (one unit per consumer warp), chunk-major inside a round so the warps advance together.
__device__ void producer_loop(const MegaArgs& a, uint32_t smem_base, int n_phases) {
  const int lane = threadIdx.x & 31;
  const int S = a.n_slots, KC = a.KC, RS = (KC + 8) * 2;
  const uint32_t ring = smem_base + HDR_BYTES + a.act_bytes;
  uint32_t job = 0;
  for (int ph = 0; ph < n_phases; ++ph) {
    Phase p;
    make_phase(a, ph, p);
    const int n_c = units_of_cta(p.units, blockIdx.x, gridDim.x);
    const int nch = (p.K + KC - 1) / KC;
    const int nrows = p.dual ? 16 : 8;
    for (int r0 = 0; r0 < n_c; r0 += NW) {
      const int nact = min(NW, n_c - r0);
      // this lane's source row for each unit slot of the round is recomputed per job (cheap)
      for (int ch = 0; ch < nch; ++ch) {
        const int k0 = ch * KC;
        const int klen = min(KC, p.K - k0);
        for (int s = 0; s < nact; ++s, ++job) {
          const int u = blockIdx.x + (r0 + s) * gridDim.x;
          int seg = 0;
          while (seg + 1 < p.nseg && u >= p.ubeg[seg + 1]) ++seg;
          const int row0 = (u - p.ubeg[seg]) << 3;
          const uint32_t slot = job % S, par = (job / S) & 1;
          const uint32_t full = smem_base + OFF_FULL + slot * 8, empty = smem_base + OFF_EMPTY + slot * 8;
          if (lane == 0) {
            mbar_wait(empty, par ^ 1);
            mbar_expect_tx(full, (uint32_t)(nrows * klen * 2));
          }
          __syncwarp();
          if (lane < nrows) {
            const bf16* Wm = (lane < 8) ? p.W[seg] : p.W2;
            int gr = row0 + (lane & 7);
            if (gr >= p.rows[seg]) gr = p.rows[seg] - 1;  // duplicate rows are never stored
            bulk_g2s(ring + slot * a.slot_bytes + lane * RS, Wm + (size_t)gr * p.K + k0, (uint32_t)(klen * 2), full);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- consumer: GEMM
struct Best {
  float v;
  int i;
};

// One warp accumulates its unit (8 weight rows, or 8 gate + 8 up rows) over one chunk.
template <bool DUAL>
__device__ __forceinline__ void mma_chunk(float (&acc)[4], float (&acc2)[4], uint32_t slot_addr, uint32_t a_addr,
                                          int RS, int nk16, bool row_valid, int lane) {
  // ldmatrix source rows: lanes 0-7 rows 0-7 @k, 8-15 rows 0-7 @k+8 (x2); DUAL adds rows 8-15 (up)
  const uint32_t baddr = slot_addr + (lane & 7) * RS + ((lane >> 3) & 1) * 16 + (DUAL ? ((lane >> 4) & 1) * 8 * RS : 0);
#pragma unroll 4
  for (int j = 0; j < nk16; ++j) {
    uint32_t b0, b1, b2 = 0, b3 = 0;
    if (DUAL)
      ldmatrix_x4(b0, b1, b2, b3, baddr + j * 32);
    else
      ldmatrix_x2(b0, b1, baddr + j * 32);
    uint32_t af[4] = {0u, 0u, 0u, 0u};
    if (row_valid) {
      af[0] = lds32(a_addr + j * 32);
      af[2] = lds32(a_addr + j * 32 + 16);
    }
    mma_bf16_16816(acc, af, b0, b1);
    if (DUAL) mma_bf16_16816(acc2, af, b2, b3);
  }
}

__device__ void gemm_phase(const MegaArgs& a, const Phase& p, unsigned char* smem, uint32_t smem_base, uint32_t& job_base,
                           Best& best) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.n_slots, KC = a.KC, RS = (KC + 8) * 2;
  const uint32_t ring = smem_base + HDR_BYTES + a.act_bytes;
  const int AS = (p.K + 8) * 2;  // activation row stride in shared memory (bytes)
  const int g = lane >> 2, c = lane & 3;
  const bool row_valid = g < a.B;
  const uint32_t a_row = smem_base + HDR_BYTES + g * AS + c * 4;
  const int n_c = units_of_cta(p.units, blockIdx.x, gridDim.x);
  const int nch = (p.K + KC - 1) / KC;
  const int Dq = a.n_q * a.hd, Dkv = a.n_kv * a.hd;

  for (int r0 = 0; r0 < n_c; r0 += NW) {
    const int nact = min(NW, n_c - r0);
    if (warp >= nact) break;
    const int u = blockIdx.x + (r0 + warp) * gridDim.x;
    int seg = 0;
    while (seg + 1 < p.nseg && u >= p.ubeg[seg + 1]) ++seg;
    const int row0 = (u - p.ubeg[seg]) << 3;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ch = 0; ch < nch; ++ch) {
      const uint32_t job = job_base + (uint32_t)(r0 * nch + ch * nact + warp);
      const uint32_t slot = job % S, par = (job / S) & 1;
      const int k0 = ch * KC;
      const int klen = min(KC, p.K - k0);
      mbar_wait(smem_base + OFF_FULL + slot * 8, par);
      if (p.dual)
        mma_chunk<true>(acc, acc2, ring + slot * a.slot_bytes, a_row + k0 * 2, RS, klen >> 4, row_valid, lane);
      else
        mma_chunk<false>(acc, acc2, ring + slot * a.slot_bytes, a_row + k0 * 2, RS, klen >> 4, row_valid, lane);
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_base + OFF_EMPTY + slot * 8);
    }
    // epilogue: c0,c1 -> (token g, n = row0 + 2c + {0,1}); tokens 8..15 do not exist here
    if (row_valid) {
      const int n = row0 + c * 2;
      if (n < p.rows[seg]) {  // rows are multiples of 2 everywhere (checked by the launcher)
        const float v0 = acc[0], v1 = acc[1];
        switch (p.kind) {
          case PH_QKV: {
            const int col = (seg == 0 ? 0 : (seg == 1 ? Dq : Dq + Dkv)) + n;
            *reinterpret_cast<uint32_t*>(a.qkv + (size_t)g * (Dq + 2 * Dkv) + col) = pack2(f2bf(v0), f2bf(v1));
            break;
          }
          case PH_O:
          case PH_DOWN: {
            uint32_t* dst = reinterpret_cast<uint32_t*>(a.x + (size_t)g * a.H + n);
            const uint32_t old = __ldcg(dst);
            const float y0 = bf2f(f2bf(v0)), y1 = bf2f(f2bf(v1));
            *dst = pack2(f2bf(__fadd_rn(lo2f(old), y0)), f2bf(__fadd_rn(hi2f(old), y1)));
            break;
          }
          case PH_GATEUP: {
            const float g0 = bf2f(f2bf(v0)), g1 = bf2f(f2bf(v1));
            const float u0 = bf2f(f2bf(acc2[0])), u1 = bf2f(f2bf(acc2[1]));
            const float s0 = bf2f(f2bf(silu_ref(g0))), s1 = bf2f(f2bf(silu_ref(g1)));
            *reinterpret_cast<uint32_t*>(a.h + (size_t)g * a.I + n) =
                pack2(f2bf(__fmul_rn(u0, s0)), f2bf(__fmul_rn(u1, s1)));
            break;
          }
          default: {  // PH_LMHEAD
            const bf16 l0 = f2bf(v0), l1 = f2bf(v1);
            *reinterpret_cast<uint32_t*>(a.logits + (size_t)g * a.V + n) = pack2(l0, l1);
            const float f0 = bf2f(l0), f1 = bf2f(l1);
            if (f0 > -CUDART_INF_F && cand_better(f0, n, best.v, best.i)) {
              best.v = f0;
              best.i = n;
            }
            if (f1 > -CUDART_INF_F && cand_better(f1, n + 1, best.v, best.i)) {
              best.v = f1;
              best.i = n + 1;
            }
            break;
          }
        }
      }
    }
  }
  job_base += (uint32_t)(n_c * nch);
}

// ---------------------------------------------------------------- consumer: activations
// rows [B, K] bf16 from global (L2) into the activation region, row stride (K+8)*2 bytes
template <typename RowPtr>
__device__ __forceinline__ void load_rows(uint32_t act, int B, int K, RowPtr row_ptr) {
  const int AS = (K + 8) * 2, per = K >> 3;
  for (int i = threadIdx.x; i < B * per; i += NTC) {
    const int b = i / per, cc = i - b * per;
    cp_async16(act + b * AS + cc * 16, row_ptr(b) + cc * 8);
  }
  cp_async_commit();
  cp_async_wait<0>();
  bar_consumers();
}

// RMSNorm of the rows in the activation region, in place (normalization.cu:5-26): the sum
// of squares is the reference's sequential FFMA chain, one thread per row.
__device__ void rmsnorm_rows(unsigned char* smem, int B, int H, const bf16* __restrict__ w) {
  const int AS = (H + 8) * 2;
  unsigned char* act = smem + HDR_BYTES;
  float* rms_s = reinterpret_cast<float*>(smem + OFF_RMS);
  if (threadIdx.x < B) {
    const uint4* row = reinterpret_cast<const uint4*>(act + threadIdx.x * AS);
    float sum = 0.f;
    for (int i = 0; i < (H >> 3); ++i) {
      const uint4 v = row[i];
      float f;
      f = lo2f(v.x); sum = __fmaf_rn(f, f, sum);
      f = hi2f(v.x); sum = __fmaf_rn(f, f, sum);
      f = lo2f(v.y); sum = __fmaf_rn(f, f, sum);
      f = hi2f(v.y); sum = __fmaf_rn(f, f, sum);
      f = lo2f(v.z); sum = __fmaf_rn(f, f, sum);
      f = hi2f(v.z); sum = __fmaf_rn(f, f, sum);
      f = lo2f(v.w); sum = __fmaf_rn(f, f, sum);
      f = hi2f(v.w); sum = __fmaf_rn(f, f, sum);
    }
    rms_s[threadIdx.x] = __fsqrt_rn(__fadd_rn(__fdiv_rn(sum, (float)H), 1e-04f));
  }
  bar_consumers();
  const int per = H >> 3;
  for (int i = threadIdx.x; i < B * per; i += NTC) {
    const int b = i / per, cc = i - b * per;
    uint4* p = reinterpret_cast<uint4*>(act + b * AS + cc * 16);
    const uint4 v = *p;
    const uint4 wv = *reinterpret_cast<const uint4*>(w + cc * 8);
    const float rms = rms_s[b];
    uint4 o;
    o.x = pack2(f2bf(__fmul_rn(__fdiv_rn(lo2f(v.x), rms), lo2f(wv.x))), f2bf(__fmul_rn(__fdiv_rn(hi2f(v.x), rms), hi2f(wv.x))));
    o.y = pack2(f2bf(__fmul_rn(__fdiv_rn(lo2f(v.y), rms), lo2f(wv.y))), f2bf(__fmul_rn(__fdiv_rn(hi2f(v.y), rms), hi2f(wv.y))));
    o.z = pack2(f2bf(__fmul_rn(__fdiv_rn(lo2f(v.z), rms), lo2f(wv.z))), f2bf(__fmul_rn(__fdiv_rn(hi2f(v.z), rms), hi2f(wv.z))));
    o.w = pack2(f2bf(__fmul_rn(__fdiv_rn(lo2f(v.w), rms), lo2f(wv.w))), f2bf(__fmul_rn(__fdiv_rn(hi2f(v.w), rms), hi2f(wv.w))));
    *p = o;
  }
  bar_consumers();
}

// ---------------------------------------------------------------- consumer: attention
// score of one cache position against one query head: products, then the reference's
// shared-memory tree (stride hd/2 ... 1, self_attension.cu:63-71) evaluated in registers.
template <int HD>
__device__ __forceinline__ float dot_tree(const float* __restrict__ q, const uint32_t (&kr)[HD / 2]) {
  float d[HD / 2];
#pragma unroll
  for (int j = 0; j < HD / 2; j += 4) {
    const float4 qa = *reinterpret_cast<const float4*>(q + j);
    const float4 qb = *reinterpret_cast<const float4*>(q + j + HD / 2);
    const uint32_t ka0 = kr[j / 2], ka1 = kr[j / 2 + 1], kb0 = kr[(j + HD / 2) / 2], kb1 = kr[(j + HD / 2) / 2 + 1];
    d[j] = __fadd_rn(__fmul_rn(qa.x, lo2f(ka0)), __fmul_rn(qb.x, lo2f(kb0)));
    d[j + 1] = __fadd_rn(__fmul_rn(qa.y, hi2f(ka0)), __fmul_rn(qb.y, hi2f(kb0)));
    d[j + 2] = __fadd_rn(__fmul_rn(qa.z, lo2f(ka1)), __fmul_rn(qb.z, lo2f(kb1)));
    d[j + 3] = __fadd_rn(__fmul_rn(qa.w, hi2f(ka1)), __fmul_rn(qb.w, hi2f(kb1)));
  }
#pragma unroll
  for (int s = HD / 4; s >= 1; s >>= 1)
#pragma unroll
    for (int j = 0; j < s; ++j) d[j] = __fadd_rn(d[j], d[j + s]);
  return d[0];
}

template <int NP>
__device__ __forceinline__ void head_load_cg(float (&x)[NP][2], const bf16* src, int lane) {
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    const uint32_t v = __ldcg(reinterpret_cast<const uint32_t*>(src + 64 * p + 2 * lane));
    x[p][0] = lo2f(v);
    x[p][1] = hi2f(v);
  }
}

// Tasks: (row, q head) while they fit one wave of CTAs, else (row, kv head) with the whole
// query group sharing the K/V stream.  q/k-norm + RoPE + the KV store of the new position
// are done here (replaces 2x qkNorm, 2x RoPE, kv_copy_layer_to_cache_decode).
template <int NP>
__device__ void attention_phase(const MegaArgs& a, int layer, unsigned char* smem) {
  constexpr int HD = 64 * NP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Gq = a.n_q / a.n_kv;
  const int Dq = a.n_q * HD, Dkv = a.n_kv * HD, QKV = Dq + 2 * Dkv;
  const bool per_head = a.B * a.n_q <= (int)gridDim.x;
  const int hs = per_head ? 1 : Gq;
  const int ntask = per_head ? a.B * a.n_q : a.B * a.n_kv;
  const int tmax = (a.max_kv_len + 3) & ~3;
  const int psz = a.kv.page_size;
  const MegaLayer& w = a.layers[layer];

  float* q_s = reinterpret_cast<float*>(smem + HDR_BYTES);          // [hs][HD] fp32
  bf16* knew = reinterpret_cast<bf16*>(q_s + hs * HD);              // [HD] bf16 (16-byte aligned)
  bf16* vnew = knew + HD;
  float* score = reinterpret_cast<float*>(vnew + HD);               // [hs][tmax]
  int* pages = reinterpret_cast<int*>(score + hs * tmax);

  for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
    int b, kvh, h0;
    if (per_head) {
      b = task / a.n_q;
      h0 = task - b * a.n_q;
      kvh = h0 / Gq;
    } else {
      b = task / a.n_kv;
      kvh = task - b * a.n_kv;
      h0 = kvh * Gq;
    }
    const bool writer = !per_head || (h0 == kvh * Gq);
    const int ps = a.pos[b];
    const int n_pages = ps / psz + 1;
    const int* bt = a.block_table + (size_t)a.slot[b] * a.max_pages;
    for (int i = threadIdx.x; i < n_pages; i += NTC) pages[i] = bt[i];
    const bf16* row = a.qkv + (size_t)b * QKV;
    const float* cos_row = a.cos_t + (size_t)ps * 32 * NP;
    const float* sin_row = a.sin_t + (size_t)ps * 32 * NP;
    const int new_page = bt[ps / psz], new_off = ps % psz;
    for (int r = warp; r < hs + 2; r += NW) {
      float v[NP][2];
      if (r < hs) {
        head_load_cg<NP>(v, row + (size_t)(h0 + r) * HD, lane);
        if (w.q_norm) head_norm<NP>(v, w.q_norm, lane);
        head_rope<NP>(v, cos_row, sin_row, lane);
#pragma unroll
        for (int p = 0; p < NP; ++p)
          *reinterpret_cast<float2*>(q_s + r * HD + 64 * p + 2 * lane) = make_float2(v[p][0], v[p][1]);
      } else if (r == hs) {
        head_load_cg<NP>(v, row + Dq + (size_t)kvh * HD, lane);
        if (w.k_norm) head_norm<NP>(v, w.k_norm, lane);
        head_rope<NP>(v, cos_row, sin_row, lane);
        head_store<NP>(v, knew, lane);
        if (writer) head_store<NP>(v, a.kv.chunk(new_page, layer, 0, kvh) + (size_t)new_off * HD, lane);
      } else {
        head_load_cg<NP>(v, row + Dq + Dkv + (size_t)kvh * HD, lane);
        head_store<NP>(v, vnew, lane);
        if (writer) head_store<NP>(v, a.kv.chunk(new_page, layer, 1, kvh) + (size_t)new_off * HD, lane);
      }
    }
    bar_consumers();

    // scores: one thread per cache position, all heads of the task
    const float den = __fsqrt_rn((float)HD);
    for (int k = threadIdx.x; k <= ps; k += NTC) {
      const uint4* kp = (k == ps) ? reinterpret_cast<const uint4*>(knew)
                                  : reinterpret_cast<const uint4*>(a.kv.chunk(pages[k / psz], layer, 0, kvh) +
                                                                   (size_t)(k % psz) * HD);
      uint32_t kr[HD / 2];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        const uint4 t = kp[i];
        kr[4 * i] = t.x;
        kr[4 * i + 1] = t.y;
        kr[4 * i + 2] = t.z;
        kr[4 * i + 3] = t.w;
      }
      for (int i = 0; i < hs; ++i) score[i * tmax + k] = __fdiv_rn(dot_tree<HD>(q_s + i * HD, kr), den);
    }
    bar_consumers();

    // softmax + PV: one warp per head (self_attension.cu:94-137)
    for (int i = warp; i < hs; i += NW) {
      float* s = score + i * tmax;
      float m = -1e9f;
      for (int k = lane; k <= ps; k += 32) m = fmaxf(m, s[k]);
      m = warp_max(m);
      for (int k = lane; k <= ps; k += 32) s[k] = expf(__fsub_rn(s[k], m));
      __syncwarp();
      float sum = 0.f;
      if (lane == 0) {
        int k = 0;
        for (; k + 4 <= ps + 1; k += 4) {
          const float4 e = *reinterpret_cast<const float4*>(s + k);
          sum = __fadd_rn(sum, e.x);
          sum = __fadd_rn(sum, e.y);
          sum = __fadd_rn(sum, e.z);
          sum = __fadd_rn(sum, e.w);
        }
        for (; k <= ps; ++k) sum = __fadd_rn(sum, s[k]);
      }
      sum = __shfl_sync(0xffffffffu, sum, 0);
      for (int k = lane; k <= ps; k += 32) s[k] = __fdiv_rn(s[k], sum);
      __syncwarp();
      float o[NP][2];
#pragma unroll
      for (int p = 0; p < NP; ++p) o[p][0] = o[p][1] = 0.f;
      int k = 0;
      for (; k + 4 <= ps; k += 4) {  // cached positions 0..ps-1 from the pool
        uint32_t vv[4][NP];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bf16* vp = a.kv.chunk(pages[(k + j) / psz], layer, 1, kvh) + (size_t)((k + j) % psz) * HD;
#pragma unroll
          for (int p = 0; p < NP; ++p) vv[j][p] = *reinterpret_cast<const uint32_t*>(vp + 64 * p + 2 * lane);
        }
        const float4 pr = *reinterpret_cast<const float4*>(s + k);
        const float prj[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            o[p][0] = __fmaf_rn(prj[j], lo2f(vv[j][p]), o[p][0]);
            o[p][1] = __fmaf_rn(prj[j], hi2f(vv[j][p]), o[p][1]);
          }
      }
      for (; k <= ps; ++k) {
        const bf16* vp = (k == ps) ? vnew : a.kv.chunk(pages[k / psz], layer, 1, kvh) + (size_t)(k % psz) * HD;
        const float pk = s[k];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const uint32_t vv = *reinterpret_cast<const uint32_t*>(vp + 64 * p + 2 * lane);
          o[p][0] = __fmaf_rn(pk, lo2f(vv), o[p][0]);
          o[p][1] = __fmaf_rn(pk, hi2f(vv), o[p][1]);
        }
      }
      head_store<NP>(o, a.att + (size_t)b * Dq + (size_t)(h0 + i) * HD, lane);
    }
    bar_consumers();  // shared memory is reused by the next task / phase
  }
}

// ---------------------------------------------------------------- the kernel
template <int NP>
__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_mega_kernel(const __grid_constant__ MegaArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5;
  const int L = a.n_layers_run > 0 ? min(a.n_layers_run, a.L) : a.L;
  const bool with_head = a.n_layers_run <= 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.n_slots; ++s) {
      mbar_init(smem_base + OFF_FULL + s * 8, 1);
      mbar_init(smem_base + OFF_EMPTY + s * 8, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == NW) {  // producer warp
    producer_loop(a, smem_base, with_head ? 4 * a.L + 1 : 4 * L);
    return;
  }

  // ---- consumers
  unsigned epoch = 0;
  uint32_t job_base = 0;
  int prof_i = 0;
  auto stamp = [&]() {
    if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) a.prof[prof_i] = globaltimer();
    ++prof_i;
  };
  const uint32_t act = smem_base + HDR_BYTES;
  const int H = a.H, Dq = a.n_q * a.hd;
  Best best{-CUDART_INF_F, -1};

  // embedding rows (embedded_matrix.cu:5-17): x[b] = E[ids[b]]; layer 0 reads E directly
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(a.embed + (size_t)a.ids[b] * H);
    uint4* dst = reinterpret_cast<uint4*>(a.x + (size_t)b * H);
    for (int i = threadIdx.x; i < (H >> 3); i += NTC) dst[i] = src[i];
  }
  stamp();
  for (int l = 0; l < L; ++l) {
    const MegaLayer& w = a.layers[l];
    Phase p;
    // ---- QKV
    if (l == 0)
      load_rows(act, a.B, H, [&](int b) { return a.embed + (size_t)a.ids[b] * H; });
    else
      load_rows(act, a.B, H, [&](int b) { return a.x + (size_t)b * H; });
    rmsnorm_rows(smem, a.B, H, w.in_ln);
    make_phase(a, 4 * l + PH_QKV, p);
    gemm_phase(a, p, smem, smem_base, job_base, best);
    grid_sync(a.bar, epoch);
    stamp();
    // ---- attention
    if (NP == 1)
      attention_phase<1>(a, l, smem);
    else
      attention_phase<NP>(a, l, smem);
    grid_sync(a.bar, epoch);
    stamp();
    // ---- O + residual
    load_rows(act, a.B, Dq, [&](int b) { return a.att + (size_t)b * Dq; });
    make_phase(a, 4 * l + PH_O, p);
    gemm_phase(a, p, smem, smem_base, job_base, best);
    grid_sync(a.bar, epoch);
    stamp();
    // ---- gate/up + SiLU*up
    load_rows(act, a.B, H, [&](int b) { return a.x + (size_t)b * H; });
    rmsnorm_rows(smem, a.B, H, w.post_ln);
    make_phase(a, 4 * l + PH_GATEUP, p);
    gemm_phase(a, p, smem, smem_base, job_base, best);
    grid_sync(a.bar, epoch);
    stamp();
    // ---- down + residual
    load_rows(act, a.B, a.I, [&](int b) { return a.h + (size_t)b * a.I; });
    make_phase(a, 4 * l + PH_DOWN, p);
    gemm_phase(a, p, smem, smem_base, job_base, best);
    grid_sync(a.bar, epoch);
    stamp();
  }
  if (!with_head) return;

  // ---- final norm + lm_head (+ greedy arg-max candidates)
  {
    Phase p;
    load_rows(act, a.B, H, [&](int b) { return a.x + (size_t)b * H; });
    rmsnorm_rows(smem, a.B, H, a.final_norm);
    make_phase(a, 4 * a.L, p);
    gemm_phase(a, p, smem, smem_base, job_base, best);
    // candidates: lanes (g, c) of a warp hold token g; reduce over c, then over warps
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best.v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best.i, o);
      if (cand_better(ov, oi, best.v, best.i)) {
        best.v = ov;
        best.i = oi;
      }
    }
    MegaCand* cs = reinterpret_cast<MegaCand*>(smem + OFF_CAND);
    if ((lane & 3) == 0 && (lane >> 2) < MAX_ROWS) cs[warp * MAX_ROWS + (lane >> 2)] = MegaCand{best.v, best.i};
    bar_consumers();
    if (threadIdx.x < a.B) {
      float bv = cs[threadIdx.x].val;
      int bi = cs[threadIdx.x].idx;
      for (int wv = 1; wv < NW; ++wv) {
        const MegaCand o = cs[wv * MAX_ROWS + threadIdx.x];
        if (cand_better(o.val, o.idx, bv, bi)) {
          bv = o.val;
          bi = o.idx;
        }
      }
      a.cand[(size_t)blockIdx.x * a.B + threadIdx.x] = MegaCand{bv, bi};
    }
    grid_sync(a.bar, epoch);
    stamp();
  }
  // ---- arg-max over the CTAs' candidates + step bookkeeping (advance_kernel)
  if (a.greedy && warp == 0) {
    const int lane = threadIdx.x & 31;
    for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
      float bv = -CUDART_INF_F;
      int bi = -1;
      for (int cta = lane; cta < (int)gridDim.x; cta += 32) {
        const MegaCand* cp = a.cand + (size_t)cta * a.B + b;
        const float ov = __ldcg(&cp->val);
        const int oi = __ldcg(&cp->idx);
        if (cand_better(ov, oi, bv, bi)) {
          bv = ov;
          bi = oi;
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
      if (lane == 0) {
        a.sampled[b] = bi;
        if (a.advance) {
          a.pos[b] += 1;
          a.ids[b] = bi;
          if (a.rowstep) a.rowstep[b] += 1;
        }
      }
    }
  }
  stamp();
}

struct Geom {
  int KC, slot_bytes, act_bytes, n_slots;
  size_t smem;
};

bool mega_geometry(int H, int I, int n_q, int n_kv, int hd, int B, int max_kv_len, int grid, Geom* g) {
  const int Dq = n_q * hd;
  if ((H % 32) || (I % 32) || (Dq % 32) || (hd != 64 && hd != 128 && hd != 256)) return false;
  if (B < 1 || B > MAX_ROWS || n_q % n_kv) return false;
  const int nch = (H + 511) / 512;
  int KC = ((H + nch - 1) / nch + 31) & ~31;
  const int Kmax = std::max(H, std::max(I, Dq));
  int act = B * (Kmax + 8) * 2;
  const int hs = (B * n_q <= grid) ? 1 : n_q / n_kv;
  const int tmax = (max_kv_len + 3) & ~3;
  const int attn = hs * hd * 4 + 2 * hd * 2 + hs * tmax * 4 + (max_kv_len / 1 + 16) * 4;  // pages: <= one per position
  act = std::max(act, attn);
  act = (act + 127) & ~127;
  const int slot = 16 * (KC + 8) * 2;
  const int budget = 227 * 1024 - HDR_BYTES - act;
  int S = budget / slot;
  if (S > MAX_SLOTS) S = MAX_SLOTS;
  if (S < 3) return false;
  g->KC = KC;
  g->slot_bytes = slot;
  g->act_bytes = act;
  g->n_slots = S;
  g->smem = (size_t)HDR_BYTES + act + (size_t)S * slot;
  return true;
}

}  // namespace

int decode_mega_max_rows(int H, int I, int n_q, int n_kv, int hd, int max_kv_len) {
  Geom g;
  int best = 0;
  for (int B = 1; B <= MAX_ROWS; ++B)
    if (mega_geometry(H, I, n_q, n_kv, hd, B, max_kv_len, 148, &g)) best = B;
  return best;
}

int decode_mega_prof_slots(int L) { return 5 * L + 4; }

cudaError_t launch_decode_mega(MegaArgs a, int num_sms, cudaStream_t st) {
  Geom g;
  if (!mega_geometry(a.H, a.I, a.n_q, a.n_kv, a.hd, a.B, a.max_kv_len, num_sms, &g)) return cudaErrorInvalidValue;
  a.KC = g.KC;
  a.slot_bytes = g.slot_bytes;
  a.act_bytes = g.act_bytes;
  a.n_slots = g.n_slots;
  void (*kern)(MegaArgs) = nullptr;
  switch (a.hd) {
    case 64: kern = decode_mega_kernel<1>; break;
    case 128: kern = decode_mega_kernel<2>; break;
    default: kern = decode_mega_kernel<4>; break;
  }
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  void* params[] = {&a};
  return cudaLaunchCooperativeKernel((const void*)kern, dim3(num_sms), dim3(MEGA_THREADS), params, g.smem, st);
}

}  // namespace qie
