// decode_gemv.cu -- the fast-numerics decode step of <= 4 sequences as ONE persistent cooperative kernel built from
// memory-bound GEMVs (sm_100a).  BASELINE.json's named target is batch-1 decode against the HBM roofline; the
// reference-order kernel (decode_mega.cu) cannot get there because the reference's arithmetic order is a chain of
// dependent HMMAs / FFMAs per output.  This kernel keeps the reference's ROUNDING POINTS (every operator
// output is rounded to bf16 where the reference rounds: matrix_mul.cu, normalization.cu, qk_norm.cu, RoPE.cu, SiLU.cu,
// element_add.cu, residual_add.cu, self_attension.cu) but sums in whatever order is fastest, so results are within the
// north star's 1e-2 tolerance per layer, not bit-exact (tests/test_gpu_mega.py, tests/test_gpu_layer_isolation.py).
//
//   * one CTA per SM, resident for the whole step; weight rows are dealt to the CTAs as CONTIGUOUS row ranges, so a
//     CTA's share of a projection is one contiguous piece of HBM.  A producer warp streams it with 1-D bulk copies
//     (cp.async.bulk + mbarrier complete_tx, ~19 KB per request) into a shared-memory ring for ALL phases of ALL layers;
//     weights do not depend on activations, so the stream runs ahead through the grid barriers: ~200 KB per SM
//     (= one layer of the 0.5B model across the chip) is in flight or resident while the token's dependent chain
//     (norm -> qkv -> attention -> o -> norm -> gate/up -> down) is resolved.
//   * 8 consumer warps = 16 half-warps; a half-warp owns one weight row: 16 lanes x 128-bit shared-memory loads of
//     the row and of the bf16 activation vector, mixed-precision FMAs (fma.rn.f32.bf16 = FHFMA.BF16: bf16 x bf16 + fp32
//     with the halves selected by the instruction, no unpack), 4 independent chains per lane, 4 shuffles to reduce.
//   * epilogues are fused: +residual (o_proj, down_proj), SiLU(gate) * up, logits + the greedy arg-max candidate
//     (reference tie-break, logit_decode.cu:15-33) ; RMSNorm runs in front of the QKV / gate+up / lm_head GEMVs on
//     the freshly loaded residual row; q/k-norm + RoPE + KV store run in front of the attention task.
//   * attention = split-KV flash decoding: task = (row, q head, KV split); scores, soft-max and PV of the split's
//     positions, partial (max, sum, o[hd]) per task; the o_proj phase combines the partials when it loads its input.
//   * phases are separated by a grid barrier (5 per layer); activations travel through L2.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "ref_math.cuh"

namespace qie {
namespace {

enum { PH_QKV = 0, PH_O = 1, PH_GATEUP = 2, PH_DOWN = 3, PH_LMHEAD = 4 };

constexpr int GV_CW = 8;                // consumer warps
constexpr int GV_CT = GV_CW * 32;       // consumer threads
constexpr int GV_THREADS = GV_CT + 32;  // + the producer warp
constexpr int GV_HW = GV_CW * 2;        // half-warps
constexpr int GV_MAX_SLOTS = 16;
constexpr int GV_PT = 2048;             // cached positions one attention task takes
constexpr int GV_MAX_LAYERS = 64;
constexpr int GV_NPJ = 7;               // bf16 pairs of a residual row per thread: H <= 2 * 256 * 7
constexpr int GV_SMEM_MAX = 227 * 1024;

struct GemvArgs {
  int n_slots, slot_bytes, off_act, off_ring;
  int n_split;   // KV splits per (row, q head)
  float* part;   // [tasks][hd + 2] partial attention results: max, sum, o[hd]
};

// shared-memory header
constexpr int GO_FULL = 0, GO_EMPTY = 128, GO_RED = 256, GO_CAND = 512, GO_LAYERS = 1024;
constexpr int GO_XS = GO_LAYERS + GV_MAX_LAYERS * (int)sizeof(MegaLayer);  // residual rows [B][H] bf16
// attention scratch inside the activation area
constexpr int GA_Q = 0, GA_K = 256, GA_V = 512, GA_SCORE = 768, GA_PAGES = GA_SCORE + GV_PT * 4,
              GA_PV = GA_PAGES + (GV_PT + 16) * 4, GA_END = GA_PV + GV_CW * 128 * 4;

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
constexpr long long GV_SPIN_LIMIT = 4000000000ll;  // ~2 s of SM clocks: trap instead of hanging the GPU
__device__ __forceinline__ bool mbar_try(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t addr, uint32_t parity) {
  if (mbar_try(addr, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(addr, parity))
    if (clock64() - t0 > GV_SPIN_LIMIT) __trap();
}
// 1-D bulk copy global -> shared memory of this CTA, completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(GV_CT) : "memory"); }
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// grid barrier of the consumer warps (see decode_mega.cu grid_sync: release add + acquire poll)
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& epoch) {
  bar_consumers();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (clock64() - t0 > GV_SPIN_LIMIT) __trap();
    } while (v < epoch);
  }
  bar_consumers();
}

// acc += w.lo * x.lo ; acc2 += w.hi * x.hi  (bf16 x bf16 products are exact in fp32; one rounding per FMA)
__device__ __forceinline__ void fma2_bf16(float& acc_lo, float& acc_hi, uint32_t w, uint32_t x) {
  asm("{\n"
      ".reg .b16 wl, wh, xl, xh;\n"
      "mov.b32 {wl, wh}, %2;\n"
      "mov.b32 {xl, xh}, %3;\n"
      "fma.rn.f32.bf16 %0, wl, xl, %0;\n"
      "fma.rn.f32.bf16 %1, wh, xh, %1;\n"
      "}\n"
      : "+f"(acc_lo), "+f"(acc_hi)
      : "r"(w), "r"(x));
}
// 8 products of one 16-byte piece into 4 chains
__device__ __forceinline__ void dot8(float (&c)[4], const uint4& w, const uint4& x) {
  fma2_bf16(c[0], c[1], w.x, x.x);
  fma2_bf16(c[2], c[3], w.y, x.y);
  fma2_bf16(c[0], c[1], w.z, x.z);
  fma2_bf16(c[2], c[3], w.w, x.w);
}

struct GPhase {
  const bf16* seg[3];
  int seg_rows[3];
  int nseg;
  const bf16* w2;  // up_proj rows travel with the gate_proj rows of the same index
  int K, rows, ch; // ch: rows (row pairs) per ring slot
};
__device__ __forceinline__ void gv_phase(const MegaArgs& a, const GemvArgs& g, const MegaLayer* layers, int kind, int l, GPhase& p) {
  const int Dq = a.n_q * a.hd, Dkv = a.n_kv * a.hd;
  p.w2 = nullptr;
  p.nseg = 1;
  p.seg[1] = p.seg[2] = nullptr;
  p.seg_rows[1] = p.seg_rows[2] = 0;
  switch (kind) {
    case PH_QKV:
      p.seg[0] = layers[l].q;
      p.seg[1] = layers[l].k;
      p.seg[2] = layers[l].v;
      p.seg_rows[0] = Dq;
      p.seg_rows[1] = p.seg_rows[2] = Dkv;
      p.nseg = 3;
      p.K = a.H;
      p.rows = Dq + 2 * Dkv;
      break;
    case PH_O:
      p.seg[0] = layers[l].o;
      p.K = Dq;
      p.rows = a.H;
      break;
    case PH_GATEUP:
      p.seg[0] = layers[l].gate;
      p.w2 = layers[l].up;
      p.K = a.H;
      p.rows = a.I;
      break;
    case PH_DOWN:
      p.seg[0] = layers[l].down;
      p.K = a.I;
      p.rows = a.H;
      break;
    default:
      p.seg[0] = a.lm_head;
      p.K = a.H;
      p.rows = a.V;
      break;
  }
  if (p.nseg == 1) p.seg_rows[0] = p.rows;
  p.ch = max(1, g.slot_bytes / (p.K * 2 * (p.w2 ? 2 : 1)));
}
// rows [r0, r1) of a phase that CTA c owns (contiguous, sizes differ by at most one)
__device__ __forceinline__ void gv_range(int rows, int& r0, int& r1) {
  r0 = (int)((long long)blockIdx.x * rows / gridDim.x);
  r1 = (int)((long long)(blockIdx.x + 1) * rows / gridDim.x);
}

// ---------------------------------------------------------------- producer: the weight stream of this CTA
__device__ __forceinline__ void gv_producer(const MegaArgs& a, const GemvArgs& g, uint32_t smem_base, const MegaLayer* layers) {
  if ((threadIdx.x & 31) != 0) return;
  uint32_t slot = 0, round = 0;
  const int nph = 4 * a.L + 1;
  for (int idx = 0; idx < nph; ++idx) {
    const int kind = idx < 4 * a.L ? (idx & 3) : PH_LMHEAD;
    GPhase p;
    gv_phase(a, g, layers, kind, idx >> 2, p);
    int r0, r1;
    gv_range(p.rows, r0, r1);
    const uint32_t row_bytes = (uint32_t)p.K * 2u;
    for (int r = r0; r < r1; r += p.ch) {
      const int n = min(p.ch, r1 - r);
      if (round > 0) mbar_wait(smem_base + GO_EMPTY + slot * 8, (round - 1) & 1);
      const uint32_t dst = smem_base + g.off_ring + slot * g.slot_bytes;
      const uint32_t full = smem_base + GO_FULL + slot * 8;
      mbar_expect_tx(full, (uint32_t)n * row_bytes * (p.w2 ? 2u : 1u));
      if (p.w2) {
        bulk_g2s(dst, p.seg[0] + (size_t)r * p.K, (uint32_t)n * row_bytes, full);
        bulk_g2s(dst + (uint32_t)n * row_bytes, p.w2 + (size_t)r * p.K, (uint32_t)n * row_bytes, full);
      } else {
        int s0 = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int s1 = s0 + p.seg_rows[s];
          const int lo = max(r, s0), hi = min(r + n, s1);
          if (lo < hi)
            bulk_g2s(dst + (uint32_t)(lo - r) * row_bytes, p.seg[s] + (size_t)(lo - s0) * p.K, (uint32_t)(hi - lo) * row_bytes, full);
          s0 = s1;
        }
      }
      if (++slot == (uint32_t)g.n_slots) {
        slot = 0;
        ++round;
      }
    }
  }
}

struct GvBest {
  float v;
  int i;
};
struct GvRing {
  uint32_t slot, round;
  int rot;  // rows dealt so far modulo the half-warps: row i of a chunk goes to half-warp (rot + i) % GV_HW
};

// ---------------------------------------------------------------- consumers: one GEMV phase
// act: [NB][K] bf16 in shared memory.  KIND selects the epilogue.
template <int NB, int KIND>
__device__ __forceinline__ void gv_gemv(const MegaArgs& a, const GemvArgs& g, const GPhase& p, unsigned char* smem, uint32_t smem_base,
                                        GvRing& ring, const unsigned char* act, GvBest& best) {
  constexpr bool PAIR = KIND == PH_GATEUP;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hw = warp * 2 + (lane >> 4), l16 = lane & 15;
  const unsigned hmask = 0xffffu << (lane & 16);
  const int K = p.K, nv = K >> 3;
  const uint32_t row_bytes = (uint32_t)K * 2u;
  const bf16* xs = reinterpret_cast<const bf16*>(smem + GO_XS);
  int r0, r1;
  gv_range(p.rows, r0, r1);
  for (int r = r0; r < r1; r += p.ch) {
    const int n = min(p.ch, r1 - r);
    mbar_wait(smem_base + GO_FULL + ring.slot * 8, ring.round & 1);
    const unsigned char* base = smem + g.off_ring + ring.slot * g.slot_bytes;
    int i = hw - ring.rot;
    if (i < 0) i += GV_HW;
    for (; i < n; i += GV_HW) {
      const uint4* wr = reinterpret_cast<const uint4*>(base + (size_t)i * row_bytes);
      const uint4* wr2 = reinterpret_cast<const uint4*>(base + (size_t)(n + i) * row_bytes);
      float c[NB][4], c2[PAIR ? NB : 1][4];
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int k = 0; k < 4; ++k) c[b][k] = 0.f;
#pragma unroll
      for (int b = 0; b < (PAIR ? NB : 1); ++b)
#pragma unroll
        for (int k = 0; k < 4; ++k) c2[b][k] = 0.f;
#pragma unroll 4
      for (int v = l16; v < nv; v += 16) {
        const uint4 w = wr[v];
        uint4 w2 = make_uint4(0u, 0u, 0u, 0u);
        if (PAIR) w2 = wr2[v];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const uint4 x = reinterpret_cast<const uint4*>(act + (size_t)b * row_bytes)[v];
          dot8(c[b], w, x);
          if (PAIR) dot8(c2[b], w2, x);
        }
      }
      float s[NB], s2[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        s[b] = (c[b][0] + c[b][1]) + (c[b][2] + c[b][3]);
        s2[b] = PAIR ? (c2[b][0] + c2[b][1]) + (c2[b][2] + c2[b][3]) : 0.f;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          s[b] += __shfl_xor_sync(hmask, s[b], o);
          if (PAIR) s2[b] += __shfl_xor_sync(hmask, s2[b], o);
        }
      }
      // lane b of the half-warp finishes row b of the batch
      float v1 = 0.f, v2 = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if (l16 == b) {
          v1 = s[b];
          v2 = s2[b];
        }
      if (l16 < a.B) {
        const int b = l16, gr = r + i;
        const float y = bf2f(f2bf(v1));  // the projection output as the reference stores it (matrix_mul.cu: bf16)
        if (KIND == PH_QKV) {
          a.qkv[(size_t)b * p.rows + gr] = f2bf(v1);
        } else if (KIND == PH_O || KIND == PH_DOWN) {
          a.x[(size_t)b * a.H + gr] = f2bf(bf2f(xs[b * a.H + gr]) + y);  // residual_add.cu:7
        } else if (KIND == PH_GATEUP) {
          const float sg = bf2f(f2bf(y * (1.0f / (1.0f + expf(-y)))));     // SiLU.cu:6-8, stored as bf16
          a.h[(size_t)b * a.I + gr] = f2bf(sg * bf2f(f2bf(v2)));            // element_add.cu (element-wise product)
        } else {
          a.logits[(size_t)b * a.V + gr] = f2bf(v1);
          if (cand_better(y, gr, best.v, best.i)) {
            best.v = y;
            best.i = gr;
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_base + GO_EMPTY + ring.slot * 8);
    ring.rot = (ring.rot + n) % GV_HW;
    if (++ring.slot == (uint32_t)g.n_slots) {
      ring.slot = 0;
      ++ring.round;
    }
  }
}

// norm weights of the next RMSNorm into registers (requested in front of the grid barrier that precedes their use)
__device__ __forceinline__ void gv_norm_w(const bf16* w, int H, uint32_t (&wr)[GV_NPJ]) {
  const uint32_t* wp = reinterpret_cast<const uint32_t*>(w);
#pragma unroll
  for (int j = 0; j < GV_NPJ; ++j) {
    const int idx = threadIdx.x + j * GV_CT;
    wr[j] = idx < (H >> 1) ? __ldg(wp + idx) : 0u;
  }
}
// residual rows -> shared memory (raw, for the residual epilogues) + their RMSNorm (normalization.cu:9-21 rounding:
// bf16((x / rms) * w)) into the activation area; the sum of squares is a tree, not the reference's chain
template <int NB>
__device__ __forceinline__ void gv_load_norm(const MegaArgs& a, unsigned char* smem, const GemvArgs& g, bool from_embed,
                                             const uint32_t (&wr)[GV_NPJ]) {
  const int H = a.H, hp = H >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t xv[NB][GV_NPJ];
  float ss[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    ss[b] = 0.f;
    const bool live = b < a.B;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(
        live ? (from_embed ? a.embed + (size_t)max(a.ids[b], 0) * H : a.x + (size_t)b * H) : a.x);
#pragma unroll
    for (int j = 0; j < GV_NPJ; ++j) {
      const int idx = threadIdx.x + j * GV_CT;
      xv[b][j] = (live && idx < hp) ? __ldcg(src + idx) : 0u;
    }
  }
  uint32_t* xs = reinterpret_cast<uint32_t*>(smem + GO_XS);
  float* red = reinterpret_cast<float*>(smem + GO_RED);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
#pragma unroll
    for (int j = 0; j < GV_NPJ; ++j) {
      const int idx = threadIdx.x + j * GV_CT;
      if (idx < hp) {
        xs[b * hp + idx] = xv[b][j];
        const float lo = lo2f(xv[b][j]), hi = hi2f(xv[b][j]);
        ss[b] += lo * lo + hi * hi;
      }
    }
    ss[b] = warp_sum(ss[b]);
    if (lane == 0) red[b * GV_CW + warp] = ss[b];
  }
  bar_consumers();
  uint32_t* xn = reinterpret_cast<uint32_t*>(smem + g.off_act);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float tot = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < GV_CW; ++w8) tot += red[b * GV_CW + w8];
    const float rms = sqrtf(tot / (float)H + 1e-04f);
#pragma unroll
    for (int j = 0; j < GV_NPJ; ++j) {
      const int idx = threadIdx.x + j * GV_CT;
      if (idx < hp)
        xn[b * hp + idx] = pack2(f2bf(__fdividef(lo2f(xv[b][j]), rms) * lo2f(wr[j])), f2bf(__fdividef(hi2f(xv[b][j]), rms) * hi2f(wr[j])));
    }
  }
  bar_consumers();
}

// ---------------------------------------------------------------- attention: split-KV flash decoding
template <int NP>
__device__ __forceinline__ void gv_attention(const MegaArgs& a, const GemvArgs& g, int layer, unsigned char* smem, const MegaLayer& w) {
  constexpr int HD = 64 * NP, LPP = HD / 8, GPW = 32 / LPP, NG = GV_CW * GPW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp * GPW + lane / LPP, j = lane % LPP;
  const unsigned gmask = (LPP == 16 ? 0xffffu : 0xffu) << (lane - j);
  const int S = g.n_split, Gq = a.n_q / a.n_kv;
  const int Dq = a.n_q * HD, Dkv = a.n_kv * HD, QKV = Dq + 2 * Dkv;
  const int ntask = a.B * a.n_q * S;
  const int psz = a.kv.page_size;
  const int psh = (psz & (psz - 1)) == 0 ? __ffs(psz) - 1 : -1;
  unsigned char* sc = smem + g.off_act;
  bf16* q_s = reinterpret_cast<bf16*>(sc + GA_Q);
  bf16* knew = reinterpret_cast<bf16*>(sc + GA_K);
  bf16* vnew = reinterpret_cast<bf16*>(sc + GA_V);
  float* score = reinterpret_cast<float*>(sc + GA_SCORE);
  int* pages = reinterpret_cast<int*>(sc + GA_PAGES);
  float* pv = reinterpret_cast<float*>(sc + GA_PV);
  float* red = reinterpret_cast<float*>(smem + GO_RED);
  const float rs = 1.0f / sqrtf((float)HD);
  for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
    const int b = task / (a.n_q * S), rem = task - b * a.n_q * S, h = rem / S, sp = rem - h * S, kvh = h / Gq;
    const int ps = a.pos[b], npos = ps + 1;
    const int len = (npos + S - 1) / S;
    const int p0 = min(npos, sp * len), p1 = min(npos, p0 + len);
    float* part = g.part + (size_t)((b * a.n_q + h) * S + sp) * (HD + 2);
    if (p0 >= p1) {  // empty split
      if (threadIdx.x < HD) part[2 + threadIdx.x] = 0.f;
      if (threadIdx.x == 0) {
        part[0] = -CUDART_INF_F;
        part[1] = 0.f;
      }
      continue;
    }
    const bool has_new = p1 == npos;  // the split holds the position this step appends
    const bool writer = has_new && h == kvh * Gq;
    const int pc1 = has_new ? p1 - 1 : p1;  // cached positions [p0, pc1)
    const int* bt = a.block_table + (size_t)a.slot[b] * a.max_pages;
    const int pg0 = psh >= 0 ? (p0 >> psh) : p0 / psz;
    const bf16* row = a.qkv + (size_t)b * QKV;
    const float* cos_row = a.cos_t + (size_t)ps * 32 * NP;
    const float* sin_row = a.sin_t + (size_t)ps * 32 * NP;
    if (warp == 0) {
      float v[NP][2];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const uint32_t u = __ldcg(reinterpret_cast<const uint32_t*>(row + (size_t)h * HD + 64 * p + 2 * lane));
        v[p][0] = lo2f(u);
        v[p][1] = hi2f(u);
      }
      if (w.q_norm) head_norm<NP>(v, w.q_norm, lane);
      head_rope<NP>(v, cos_row, sin_row, lane);
      head_store<NP>(v, q_s, lane);
    } else if (warp == 1) {
      if (has_new) {
        float v[NP][2];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const uint32_t u = __ldcg(reinterpret_cast<const uint32_t*>(row + Dq + (size_t)kvh * HD + 64 * p + 2 * lane));
          v[p][0] = lo2f(u);
          v[p][1] = hi2f(u);
        }
        if (w.k_norm) head_norm<NP>(v, w.k_norm, lane);
        head_rope<NP>(v, cos_row, sin_row, lane);
        head_store<NP>(v, knew, lane);
        if (writer) head_store<NP>(v, a.kv.chunk(bt[ps / psz], layer, 0, kvh) + (size_t)(ps % psz) * HD, lane);
      }
    } else if (warp == 2) {
      if (has_new) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const uint32_t u = __ldcg(reinterpret_cast<const uint32_t*>(row + Dq + Dkv + (size_t)kvh * HD + 64 * p + 2 * lane));
          reinterpret_cast<uint32_t*>(vnew)[32 * p + lane] = u;
          if (writer)
            reinterpret_cast<uint32_t*>(a.kv.chunk(bt[ps / psz], layer, 1, kvh) + (size_t)(ps % psz) * HD)[32 * p + lane] = u;
        }
      }
    } else if (pc1 > p0) {
      const int pgn = (psh >= 0 ? ((pc1 - 1) >> psh) : (pc1 - 1) / psz) - pg0 + 1;
      for (int i = threadIdx.x - 96; i < pgn; i += GV_CT - 96) pages[i] = bt[pg0 + i];
    }
    bar_consumers();
    // ---- scores: LPP lanes per cached position, 16 bytes of the row each
    const uint4 qv = reinterpret_cast<const uint4*>(q_s)[j];
    float mx = -CUDART_INF_F;
    for (int base = p0; base < pc1; base += 4 * NG) {
      uint4 kr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = base + u * NG + grp;
        kr[u] = make_uint4(0u, 0u, 0u, 0u);
        if (p < pc1) {
          const int pi = psh >= 0 ? (p >> psh) : p / psz;
          kr[u] = ld_nc_v4(a.kv.chunk(pages[pi - pg0], layer, 0, kvh) + (size_t)(p - pi * psz) * HD + j * 8);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = base + u * NG + grp;
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        dot8(c, qv, kr[u]);
        float d = (c[0] + c[1]) + (c[2] + c[3]);
#pragma unroll
        for (int o = LPP / 2; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        d *= rs;
        if (p < pc1) {
          if (j == 0) score[p - p0] = d;
          mx = fmaxf(mx, d);
        }
      }
    }
    if (has_new && grp == 0) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      dot8(c, qv, reinterpret_cast<const uint4*>(knew)[j]);
      float d = (c[0] + c[1]) + (c[2] + c[3]);
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) d += __shfl_xor_sync(gmask, d, o);
      d *= rs;
      if (j == 0) score[ps - p0] = d;
      mx = fmaxf(mx, d);
    }
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    bar_consumers();
    float M = red[0];
#pragma unroll
    for (int w8 = 1; w8 < GV_CW; ++w8) M = fmaxf(M, red[w8]);
    // ---- exp + sum
    const int n = p1 - p0;
    float ls = 0.f;
    for (int i = threadIdx.x; i < n; i += GV_CT) {
      const float e = __expf(score[i] - M);
      score[i] = e;
      ls += e;
    }
    ls = warp_sum(ls);
    if (lane == 0) red[GV_CW + warp] = ls;
    bar_consumers();
    // ---- PV: the same position groups, 8 output columns per lane
    float o8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o8[k] = 0.f;
    auto fold = [&](const uint4& vr, float e) {
      o8[0] = fmaf(e, lo2f(vr.x), o8[0]);
      o8[1] = fmaf(e, hi2f(vr.x), o8[1]);
      o8[2] = fmaf(e, lo2f(vr.y), o8[2]);
      o8[3] = fmaf(e, hi2f(vr.y), o8[3]);
      o8[4] = fmaf(e, lo2f(vr.z), o8[4]);
      o8[5] = fmaf(e, hi2f(vr.z), o8[5]);
      o8[6] = fmaf(e, lo2f(vr.w), o8[6]);
      o8[7] = fmaf(e, hi2f(vr.w), o8[7]);
    };
    for (int base = p0; base < pc1; base += 4 * NG) {
      uint4 vr[4];
      float e4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = base + u * NG + grp;
        vr[u] = make_uint4(0u, 0u, 0u, 0u);
        e4[u] = 0.f;
        if (p < pc1) {
          const int pi = psh >= 0 ? (p >> psh) : p / psz;
          vr[u] = ld_nc_v4(a.kv.chunk(pages[pi - pg0], layer, 1, kvh) + (size_t)(p - pi * psz) * HD + j * 8);
          e4[u] = score[p - p0];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) fold(vr[u], e4[u]);
    }
    if (has_new && grp == 0) fold(reinterpret_cast<const uint4*>(vnew)[j], score[ps - p0]);
#pragma unroll
    for (int off = LPP; off < 32; off <<= 1)
#pragma unroll
      for (int k = 0; k < 8; ++k) o8[k] += __shfl_xor_sync(0xffffffffu, o8[k], off);
    if (lane < LPP) {
#pragma unroll
      for (int k = 0; k < 8; ++k) pv[warp * HD + j * 8 + k] = o8[k];
    }
    bar_consumers();
    if (threadIdx.x < HD) {
      float t = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < GV_CW; ++w8) t += pv[w8 * HD + threadIdx.x];
      part[2 + threadIdx.x] = t;
    }
    if (threadIdx.x == 0) {
      float l = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < GV_CW; ++w8) l += red[GV_CW + w8];
      part[0] = M;
      part[1] = l;
    }
    bar_consumers();  // the scratch is reused by the next task
  }
}

// attention output rows [B][Dq] bf16 (self_attension.cu:137: rounded once) from the tasks' partial results
__device__ __forceinline__ void gv_load_att(const MegaArgs& a, const GemvArgs& g, unsigned char* smem) {
  const int HD = a.hd, Dq = a.n_q * HD, S = g.n_split;
  bf16* att = reinterpret_cast<bf16*>(smem + g.off_act);
  for (int e = threadIdx.x; e < a.B * Dq; e += GV_CT) {
    const int b = e / Dq, r = e - b * Dq, h = r / HD, d = r - h * HD;
    const float* pp = g.part + (size_t)((b * a.n_q + h) * S) * (HD + 2);
    float o;
    if (S == 1) {
      o = __ldcg(pp + 2 + d) / __ldcg(pp + 1);
    } else {
      float M = -CUDART_INF_F;
      for (int s = 0; s < S; ++s) M = fmaxf(M, __ldcg(pp + s * (HD + 2)));
      float L = 0.f;
      o = 0.f;
      for (int s = 0; s < S; ++s) {
        const float m = __ldcg(pp + s * (HD + 2));
        const float wgt = m == -CUDART_INF_F ? 0.f : __expf(m - M);
        L += __ldcg(pp + s * (HD + 2) + 1) * wgt;
        o += __ldcg(pp + s * (HD + 2) + 2 + d) * wgt;
      }
      o /= L;
    }
    att[e] = f2bf(o);
  }
  bar_consumers();
}
__device__ __forceinline__ void gv_load_rows(unsigned char* dst, const bf16* src, int n_elems) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (int i = threadIdx.x; i < (n_elems >> 3); i += GV_CT) d4[i] = __ldcg(s4 + i);
  bar_consumers();
}

template <int NP, int NB>
__global__ void __launch_bounds__(GV_THREADS, 1) decode_gemv_kernel(const __grid_constant__ MegaArgs a, const __grid_constant__ GemvArgs g) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < g.n_slots; ++s) {
      mbar_init(smem_base + GO_FULL + s * 8, 1);
      mbar_init(smem_base + GO_EMPTY + s * 8, GV_CW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.layers);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + GO_LAYERS);
    for (int i = threadIdx.x; i < a.L * (int)(sizeof(MegaLayer) / 4); i += GV_THREADS) dst[i] = src[i];
  }
  __syncthreads();
  const MegaLayer* layers = reinterpret_cast<const MegaLayer*>(smem + GO_LAYERS);
  if (warp == GV_CW) {
    gv_producer(a, g, smem_base, layers);
    return;
  }
  unsigned epoch = 0;
  GvRing ring{0u, 0u, 0};
  int prof_i = 0;
  auto stamp = [&]() {
    if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) {
      a.prof[prof_i] = globaltimer();
      a.prof[a.prof_stride + prof_i] = (unsigned long long)clock64();
    }
    ++prof_i;
  };
  const unsigned char* act = smem + g.off_act;
  GvBest best{-CUDART_INF_F, -1};
  uint32_t wr[GV_NPJ];
  gv_norm_w(layers[0].in_ln, a.H, wr);
  stamp();
  GPhase p;
  for (int l = 0; l < a.L; ++l) {
    // ---- RMSNorm + QKV
    stamp();  // (qkv.load: the row load is part of the norm here)
    gv_load_norm<NB>(a, smem, g, l == 0, wr);
    stamp();
    gv_phase(a, g, layers, PH_QKV, l, p);
    gv_gemv<NB, PH_QKV>(a, g, p, smem, smem_base, ring, act, best);
    gv_norm_w(layers[l].post_ln, a.H, wr);
    stamp();
    grid_sync(a.bar, epoch);
    stamp();
    // ---- q/k-norm + RoPE + KV store + attention
    gv_attention<NP>(a, g, l, smem, layers[l]);
    stamp();
    grid_sync(a.bar, epoch);
    stamp();
    // ---- O + residual
    gv_load_att(a, g, smem);
    stamp();
    gv_phase(a, g, layers, PH_O, l, p);
    gv_gemv<NB, PH_O>(a, g, p, smem, smem_base, ring, act, best);
    stamp();
    grid_sync(a.bar, epoch);
    stamp();
    // ---- RMSNorm + gate/up + SiLU * up
    stamp();
    gv_load_norm<NB>(a, smem, g, false, wr);
    stamp();
    gv_phase(a, g, layers, PH_GATEUP, l, p);
    gv_gemv<NB, PH_GATEUP>(a, g, p, smem, smem_base, ring, act, best);
    gv_norm_w(l + 1 < a.L ? layers[l + 1].in_ln : a.final_norm, a.H, wr);
    stamp();
    grid_sync(a.bar, epoch);
    stamp();
    // ---- down + residual
    gv_load_rows(smem + g.off_act, a.h, a.B * a.I);
    stamp();
    gv_phase(a, g, layers, PH_DOWN, l, p);
    gv_gemv<NB, PH_DOWN>(a, g, p, smem, smem_base, ring, act, best);
    stamp();
    grid_sync(a.bar, epoch);
    stamp();
  }
  // ---- final norm + lm_head + arg-max candidates
  stamp();
  gv_load_norm<NB>(a, smem, g, false, wr);
  stamp();
  gv_phase(a, g, layers, PH_LMHEAD, 0, p);
  gv_gemv<NB, PH_LMHEAD>(a, g, p, smem, smem_base, ring, act, best);
  stamp();
  {
    MegaCand* cs = reinterpret_cast<MegaCand*>(smem + GO_CAND);  // [GV_HW][4]
    const int hw = warp * 2 + (lane >> 4), l16 = lane & 15;
    if (l16 < 4) cs[hw * 4 + l16] = MegaCand{best.v, best.i};
    bar_consumers();
    if (threadIdx.x < a.B) {
      float bv = cs[threadIdx.x].val;
      int bi = cs[threadIdx.x].idx;
      for (int k = 1; k < GV_HW; ++k) {
        const MegaCand o = cs[k * 4 + threadIdx.x];
        if (cand_better(o.val, o.idx, bv, bi)) {
          bv = o.val;
          bi = o.idx;
        }
      }
      a.cand[(size_t)blockIdx.x * a.B + threadIdx.x] = MegaCand{bv, bi};
    }
  }
  grid_sync(a.bar, epoch);
  stamp();
  if (a.greedy && warp == 0) {
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

struct GvGeom {
  GemvArgs g;
  size_t smem;
};
bool gv_geometry(int H, int I, int L, int n_q, int n_kv, int hd, int B, int max_kv_len, int grid, GvGeom* out) {
  const int Dq = n_q * hd;
  if (B < 1 || B > DECODE_GEMV_MAX_ROWS || L < 1 || L > GV_MAX_LAYERS) return false;
  if (hd != 64 && hd != 128) return false;
  if (n_kv < 1 || n_q % n_kv != 0) return false;
  if (H % 128 || I % 128 || Dq % 128 || H > 2 * GV_CT * GV_NPJ) return false;
  const int nb = B <= 1 ? 1 : (B <= 2 ? 2 : 4);
  const int maxk = std::max(std::max(H, I), Dq);
  const int act_bytes = std::max(GA_END, nb * maxk * 2);
  GemvArgs g{};
  g.off_act = (GO_XS + nb * H * 2 + 127) & ~127;
  g.off_ring = (g.off_act + act_bytes + 127) & ~127;
  int rd = std::max(1, 20480 / (I * 2));
  int slot = rd * I * 2;
  slot = std::max(slot, std::max(4 * H, 2 * Dq));
  slot = (slot + 127) & ~127;
  g.slot_bytes = slot;
  g.n_slots = std::min(GV_MAX_SLOTS, (GV_SMEM_MAX - g.off_ring) / slot);
  if (g.n_slots < 3) return false;
  // KV splits: ~256 positions per task while the tasks fit one wave
  int S = std::max(1, std::min((max_kv_len + 255) / 256, grid / std::max(1, B * n_q)));
  if ((max_kv_len + 1 + S - 1) / S + 1 > GV_PT) return false;
  g.n_split = S;
  out->g = g;
  out->smem = (size_t)g.off_ring + (size_t)g.n_slots * slot;
  return true;
}

}  // namespace

bool decode_gemv_supports(int H, int I, int L, int n_q, int n_kv, int hd, int B, int max_kv_len, int num_sms) {
  GvGeom gg;
  return gv_geometry(H, I, L, n_q, n_kv, hd, B, max_kv_len, num_sms, &gg);
}
size_t decode_gemv_part_floats(int n_q, int hd, int num_sms) {
  // partial attention results: (max, sum, o[hd]) per task, tasks <= max(rows * heads, SMs) splits included
  return (size_t)(DECODE_GEMV_MAX_ROWS * n_q + num_sms) * (size_t)(hd + 2);
}

cudaError_t launch_decode_gemv(MegaArgs a, float* part, int num_sms, cudaStream_t st) {
  GvGeom gg;
  if (!gv_geometry(a.H, a.I, a.L, a.n_q, a.n_kv, a.hd, a.B, a.max_kv_len, num_sms, &gg)) return cudaErrorInvalidValue;
  gg.g.part = part;
  const int nb = a.B <= 1 ? 1 : (a.B <= 2 ? 2 : 4);
  void (*kern)(MegaArgs, GemvArgs);
  if (a.hd == 64)
    kern = nb == 1 ? decode_gemv_kernel<1, 1> : (nb == 2 ? decode_gemv_kernel<1, 2> : decode_gemv_kernel<1, 4>);
  else
    kern = nb == 1 ? decode_gemv_kernel<2, 1> : (nb == 2 ? decode_gemv_kernel<2, 2> : decode_gemv_kernel<2, 4>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gg.smem);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  void* params[] = {&a, &gg.g};
  return cudaLaunchCooperativeKernel((const void*)kern, dim3(num_sms), dim3(GV_THREADS), params, gg.smem, st);
}

}  // namespace qie
