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
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "ref_math.cuh"

#ifndef QIE_PV_PIPE
#define QIE_PV_PIPE 1  // PV loop: request the next group's operands before the current FFMA chain (A/B knob)
#endif

namespace qie {
namespace {

constexpr int NW = 7;                         // consumer warps (7 + the producer warp = 256 threads: 255 registers per thread)
constexpr int NTC = NW * 32;                  // consumer threads
constexpr int MEGA_THREADS = NTC + 32;        // + producer warp
constexpr int MAX_SLOTS = 32;
constexpr int VT = 128;                       // cached positions per V tile in shared memory
constexpr int MAX_ROWS = 64;                  // rows (sequences) per launch
constexpr int MAX_LAYERS = 64;
// shared-memory header
constexpr int OFF_FULL = 0, OFF_EMPTY = 256, OFF_ISSUED = 512, OFF_DBG = 528, OFF_REL = 768, OFF_RMS = 1024, OFF_CAND = 1280;  // OFF_DBG: 16 x u64 cycle counters; OFF_REL: u32 per slot
constexpr int OFF_LAYERS = OFF_CAND + NW * MAX_ROWS * 8;                 // MegaLayer[MAX_LAYERS]
constexpr int OFF_WNORM = OFF_LAYERS + MAX_LAYERS * (int)sizeof(MegaLayer);  // 2 x [H] bf16 norm weights
constexpr int OFF_TSF = 944;    // 4 mbarriers: stage of the tile-split A stream filled (gemm_ts)
constexpr int OFF_TSE = 976;    // 4 mbarriers: ... read by every warp that has a unit (count = units of this CTA, set at kernel start)
constexpr int OFF_TSSEQ = 748;  // u32: chunks of that stream requested so far (stage = seq & 3, fill parity = (seq >> 2) & 1)
static_assert(sizeof(MegaLayer) == 88, "MegaLayer layout");

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
constexpr long long SPIN_LIMIT = 4000000000ll;  // ~2 s of SM clocks: trap instead of hanging the GPU
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
    if (clock64() - t0 > SPIN_LIMIT) __trap();
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, uint32_t mbar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(NTC) : "memory"); }
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32_volatile(uint32_t addr, uint32_t v) {
  asm volatile("st.volatile.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Grid barrier for the consumer warps of all CTAs (cooperative launch: all CTAs resident).
// bar.sync orders the CTA's writes before thread 0's release; the release/acquire pair on
// the counter publishes them device-wide.  Readers fetch activations with ld.cg /
// cp.async.cg (L2), never through L1.
// The poll is an acquire load (r02: ld.relaxed + a trailing fence.acq_rel.gpu cost ~0.3 us more per barrier --
// batch 64 2954 -> 2918 us per step, batch 1 1042 -> 1007 us per token); it reads the head of the release sequence the
// CTAs' red.release additions form, so it synchronises with all of them.
// Measured alternatives (B200, batch 1 and 64): arrivals sharded over 8 counters on separate
// lines -> same step time (the wait is the slowest CTA of the phase, not the atomics); one flag
// word per CTA polled by 148 threads of every CTA -> 2x slower barriers (hot lines).
__device__ __forceinline__ void grid_sync(unsigned* ctr, unsigned& epoch) {
  bar_consumers();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    unsigned v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (clock64() - t0 > SPIN_LIMIT) __trap();
    } while (v < epoch);
  }
  bar_consumers();
}

// Tensor-parallel exchange point (behind o_proj and down_proj): local grid barrier, then CTA 0 tells every
// peer "this rank has finished exchange number xepoch" with a system-scope release store into the peer's flag
// word, and every CTA waits until all peers have said the same.  Each CTA's thread 0 fences at system scope
// before it arrives at the local barrier, so the CTA's peer stores are ordered before CTA 0's flag.
__device__ __forceinline__ void tp_exchange_sync(const MegaArgs& a, unsigned& epoch, unsigned xepoch) {
  bar_consumers();
  if (threadIdx.x == 0) {
    asm volatile("fence.acq_rel.sys;" ::: "memory");
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
    unsigned v;
    long long t0 = clock64();
    do {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.bar) : "memory");
      if (clock64() - t0 > SPIN_LIMIT) __trap();
    } while (v < epoch);
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    if (blockIdx.x == 0)
      for (int r = 0; r < a.tp_size; ++r)
        if (r != a.tp_rank)
          asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(a.tp_flag[r] + a.tp_rank), "r"(xepoch) : "memory");
    for (int r = 0; r < a.tp_size; ++r) {
      if (r == a.tp_rank) continue;
      t0 = clock64();
      do {
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(a.tp_flag[a.tp_rank] + r) : "memory");
        if (clock64() - t0 > 4 * SPIN_LIMIT) __trap();  // the peer may still be launching its kernel
      } while ((int)(v - xepoch) < 0);
    }
    asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
  bar_consumers();
}

// Residual rows after an exchange: x_new = bf16(x + bf16(sum over ranks of the fp32 partial sums)), ranks in
// ascending order on every rank (identical bits everywhere), written into the activation region (layout of
// load_rows) by every CTA and into the other residual buffer by CTA 0.
__device__ __forceinline__ void load_rows_tp(const MegaArgs& a, uint32_t act, int B, int H, const bf16* x_cur, bf16* x_next, int xpar) {
  const int AS = (H + 8) * 2, hv = H >> 3;  // 8 elements per thread and step: one 16-byte x load, 2 x 16 bytes per rank
  const float* part = a.tp_part[a.tp_rank] + (size_t)xpar * a.tp_size * MEGA_TP_ROWS * H;
  const size_t rstride = (size_t)MEGA_TP_ROWS * H;
#pragma unroll 2
  for (int i = threadIdx.x; i < B * hv; i += NTC) {
    const int b = i / hv, n = (i - b * hv) * 8;
    const size_t e = (size_t)b * H + n;
    const uint4 old = __ldcg(reinterpret_cast<const uint4*>(x_cur + e));
    float4 s0 = __ldcg(reinterpret_cast<const float4*>(part + e));
    float4 s1 = __ldcg(reinterpret_cast<const float4*>(part + e + 4));
    for (int r = 1; r < a.tp_size; ++r) {
      const float4 p0 = __ldcg(reinterpret_cast<const float4*>(part + r * rstride + e));
      const float4 p1 = __ldcg(reinterpret_cast<const float4*>(part + r * rstride + e + 4));
      s0.x = __fadd_rn(s0.x, p0.x); s0.y = __fadd_rn(s0.y, p0.y); s0.z = __fadd_rn(s0.z, p0.z); s0.w = __fadd_rn(s0.w, p0.w);
      s1.x = __fadd_rn(s1.x, p1.x); s1.y = __fadd_rn(s1.y, p1.y); s1.z = __fadd_rn(s1.z, p1.z); s1.w = __fadd_rn(s1.w, p1.w);
    }
    auto add2 = [](uint32_t o, float y0, float y1) {
      return pack2(f2bf(__fadd_rn(lo2f(o), bf2f(f2bf(y0)))), f2bf(__fadd_rn(hi2f(o), bf2f(f2bf(y1)))));
    };
    uint4 nw;
    nw.x = add2(old.x, s0.x, s0.y);
    nw.y = add2(old.y, s0.z, s0.w);
    nw.z = add2(old.z, s1.x, s1.y);
    nw.w = add2(old.w, s1.z, s1.w);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(act + b * AS + n * 2), "r"(nw.x), "r"(nw.y), "r"(nw.z), "r"(nw.w) : "memory");
    if (blockIdx.x == 0) *reinterpret_cast<uint4*>(x_next + e) = nw;
  }
  bar_consumers();
}

// One row (dist_norm: CTA b owns row b): the sum lands in row 0 of the activation region and in the other
// residual buffer; nobody else computes this row.
__device__ __forceinline__ void load_row_tp(const MegaArgs& a, uint32_t act, int b, int H, const bf16* x_cur, bf16* x_next, int xpar) {
  const int hv = H >> 3;
  const float* part = a.tp_part[a.tp_rank] + (size_t)xpar * a.tp_size * MEGA_TP_ROWS * H;
  const size_t rstride = (size_t)MEGA_TP_ROWS * H;
  for (int i = threadIdx.x; i < hv; i += NTC) {
    const int n = i * 8;
    const size_t e = (size_t)b * H + n;
    const uint4 old = __ldcg(reinterpret_cast<const uint4*>(x_cur + e));
    float4 s0 = __ldcg(reinterpret_cast<const float4*>(part + e));
    float4 s1 = __ldcg(reinterpret_cast<const float4*>(part + e + 4));
    for (int r = 1; r < a.tp_size; ++r) {
      const float4 p0 = __ldcg(reinterpret_cast<const float4*>(part + r * rstride + e));
      const float4 p1 = __ldcg(reinterpret_cast<const float4*>(part + r * rstride + e + 4));
      s0.x = __fadd_rn(s0.x, p0.x); s0.y = __fadd_rn(s0.y, p0.y); s0.z = __fadd_rn(s0.z, p0.z); s0.w = __fadd_rn(s0.w, p0.w);
      s1.x = __fadd_rn(s1.x, p1.x); s1.y = __fadd_rn(s1.y, p1.y); s1.z = __fadd_rn(s1.z, p1.z); s1.w = __fadd_rn(s1.w, p1.w);
    }
    auto add2 = [](uint32_t o, float y0, float y1) {
      return pack2(f2bf(__fadd_rn(lo2f(o), bf2f(f2bf(y0)))), f2bf(__fadd_rn(hi2f(o), bf2f(f2bf(y1)))));
    };
    uint4 nw;
    nw.x = add2(old.x, s0.x, s0.y);
    nw.y = add2(old.y, s0.z, s0.w);
    nw.z = add2(old.z, s1.x, s1.y);
    nw.w = add2(old.w, s1.z, s1.w);
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(act + n * 2), "r"(nw.x), "r"(nw.y), "r"(nw.z), "r"(nw.w) : "memory");
    *reinterpret_cast<uint4*>(x_next + e) = nw;
  }
  bar_consumers();
}

// The same sum for more than a few rows: every CTA reading every rank's fp32 partial sums of the whole batch
// costs (2 + 4 tp) bytes per element and CTA of L2 traffic, so the rows are combined ONCE -- item i by CTA
// i mod grid -- into the other residual buffer, followed by a grid barrier and a plain bf16 row load.
__device__ __forceinline__ void combine_rows_tp(const MegaArgs& a, int B, int H, const bf16* x_cur, bf16* x_next, int xpar) {
  const int hv = H >> 3;
  const float* part = a.tp_part[a.tp_rank] + (size_t)xpar * a.tp_size * MEGA_TP_ROWS * H;
  const size_t rstride = (size_t)MEGA_TP_ROWS * H;
  for (int i = blockIdx.x + threadIdx.x * gridDim.x; i < B * hv; i += NTC * gridDim.x) {
    const int b = i / hv, n = (i - b * hv) * 8;
    const size_t e = (size_t)b * H + n;
    const uint4 old = __ldcg(reinterpret_cast<const uint4*>(x_cur + e));
    float4 s0 = __ldcg(reinterpret_cast<const float4*>(part + e));
    float4 s1 = __ldcg(reinterpret_cast<const float4*>(part + e + 4));
    for (int r = 1; r < a.tp_size; ++r) {
      const float4 p0 = __ldcg(reinterpret_cast<const float4*>(part + r * rstride + e));
      const float4 p1 = __ldcg(reinterpret_cast<const float4*>(part + r * rstride + e + 4));
      s0.x = __fadd_rn(s0.x, p0.x); s0.y = __fadd_rn(s0.y, p0.y); s0.z = __fadd_rn(s0.z, p0.z); s0.w = __fadd_rn(s0.w, p0.w);
      s1.x = __fadd_rn(s1.x, p1.x); s1.y = __fadd_rn(s1.y, p1.y); s1.z = __fadd_rn(s1.z, p1.z); s1.w = __fadd_rn(s1.w, p1.w);
    }
    auto add2 = [](uint32_t o, float y0, float y1) {
      return pack2(f2bf(__fadd_rn(lo2f(o), bf2f(f2bf(y0)))), f2bf(__fadd_rn(hi2f(o), bf2f(f2bf(y1)))));
    };
    uint4 nw;
    nw.x = add2(old.x, s0.x, s0.y);
    nw.y = add2(old.y, s0.z, s0.w);
    nw.z = add2(old.z, s1.x, s1.y);
    nw.w = add2(old.w, s1.z, s1.w);
    *reinterpret_cast<uint4*>(x_next + e) = nw;
  }
}

// ---------------------------------------------------------------- GEMM phase description
struct Phase {
  const TensorMap2D* W[3];  // weight tensor maps (3-D view, box = 8 rows x KC)
  const TensorMap2D* W2;
  int rows[3];
  int ubeg[4];
  int nseg, K, dual, units, kind;
  int n_c, nch, nu;  // units of THIS CTA, chunks per unit, units per round (host-precomputed tables)
  int upr, row2;     // weight rows per unit (8, or 16 for lm_head pairs), row offset of the second tile of a dual unit
  int xpar;          // tensor parallel: which half of the exchange buffer this phase's partial sums go to
  // unit i of this CTA is ufirst + i * ustep.  Default: blockIdx.x + i * gridDim.x.  Tile-split phases (ts): the CTA
  // owns ONE 16-row token tile (tile) and the units grp, grp + G, ... of ALL output rows, see gemm_ts
  int ufirst, ustep, cls, ts, tile;
};

__device__ __forceinline__ void make_phase(const MegaArgs& a, int idx, Phase& p) {
  const int hd = a.hd, Dq = a.n_q * hd, Dkv = a.n_kv * hd;
  p.W2 = nullptr;
  p.dual = 0;
  p.nseg = 1;
  p.xpar = 0;
  if (idx >= 4 * a.L) {
    // lm_head: a unit is a PAIR of 8-row tiles (rows r..r+7 and r+8..r+15) accumulated side by side
    // like gate/up: two independent HMMA chains per warp and each activation fragment feeds both
    p.kind = PH_LMHEAD;
    p.W[0] = a.wmaps + 7 * a.L;
    p.W2 = p.W[0];
    p.dual = 1;
    p.rows[0] = a.V;
    p.K = a.H;
  } else {
    const TensorMap2D* m = a.wmaps + 7 * (idx >> 2);
    p.kind = idx & 3;
    switch (p.kind) {
      case PH_QKV:
        p.nseg = 3;
        p.W[0] = m; p.W[1] = m + 1; p.W[2] = m + 2;
        p.rows[0] = Dq; p.rows[1] = Dkv; p.rows[2] = Dkv;
        p.K = a.H;
        break;
      case PH_O:
        p.W[0] = m + 3; p.rows[0] = a.H; p.K = Dq;
        break;
      case PH_GATEUP:
        p.W[0] = m + 4; p.W2 = m + 5; p.rows[0] = a.I; p.K = a.H; p.dual = 1;
        break;
      default:
        p.W[0] = m + 6; p.rows[0] = a.H; p.K = a.I;
        break;
    }
  }
  p.upr = p.kind == PH_LMHEAD ? 16 : 8;
  p.row2 = p.kind == PH_LMHEAD ? 8 : 0;
  int u = 0;
  for (int s = 0; s < p.nseg; ++s) {
    p.ubeg[s] = u;
    u += (p.rows[s] + p.upr - 1) / p.upr;
  }
  p.ubeg[p.nseg] = u;
  p.units = u;
  p.ts = a.ph_ts[p.kind];
  if (p.ts) {
    const int G = a.ph_G[p.kind], grp = (int)blockIdx.x / a.mtt;
    p.tile = (int)blockIdx.x - grp * a.mtt;
    p.ufirst = grp;
    p.ustep = G;
    p.cls = grp < a.ph_r[p.kind] ? 1 : 0;
    p.n_c = grp < G ? a.ph_q[p.kind] + p.cls : 0;  // CTAs beyond mtt * G (grid not a multiple of the tile count) sit the phase out
  } else {
    p.tile = 0;
    p.ufirst = (int)blockIdx.x;
    p.ustep = (int)gridDim.x;
    p.cls = (int)blockIdx.x < a.ph_r[p.kind] ? 1 : 0;
    p.n_c = a.ph_q[p.kind] + p.cls;
  }
  p.nch = a.ph_nch[p.kind];
  p.nu = a.ph_nu[p.kind];
}

struct RingPos {  // position of a job in the weight ring, advanced without divisions
  uint32_t job, slot, par;
  __device__ __forceinline__ void advance(uint32_t jobs, uint32_t dslot, uint32_t dpar, uint32_t S) {  // dslot = jobs % S
    job += jobs;
    slot += dslot;
    par ^= dpar;
    if (slot >= S) {
      slot -= S;
      par ^= 1;
    }
  }
  __device__ __forceinline__ void step(uint32_t d, uint32_t S) {  // small d (< a few S)
    job += d;
    slot += d;
    while (slot >= S) {
      slot -= S;
      par ^= 1;
    }
  }
};

// ---------------------------------------------------------------- producer
// One elected thread walks every GEMM phase of the step in order; per phase the units of this
// CTA in rounds of NW (one unit per consumer warp), chunk-major inside a round so the warps
// advance together.  A job = ONE TMA request (two for gate+up): a box of 8 weight rows x KC
// k-elements lands in the slot as [k-block][row][64] with the 128-byte swizzle.
// `issued` (shared memory) tells consumers that job j's mbarrier phase has been armed: a
// consumer that is a whole ring lap ahead must not test the parity of an older phase.
__device__ __forceinline__ void producer_loop(const MegaArgs& a, uint32_t smem_base, int n_phases) {
  if ((threadIdx.x & 31) != 0) return;
  const uint32_t S = a.n_slots;
  const int KC = a.KC;
  const uint32_t ring = smem_base + a.off_ring;
  const uint32_t box_bytes = (uint32_t)(8 * KC * 2);
  uint32_t job = 0, slot = 0, par = 0;
  for (int ph = 0; ph < n_phases; ++ph) {
    Phase p;
    make_phase(a, ph, p);
    const int n_c = p.n_c, nch = p.nch;
    for (int r0 = 0; r0 < n_c; r0 += p.nu) {
      const int nact = min(p.nu, n_c - r0);
      for (int ch = 0; ch < nch; ++ch) {
        for (int s = 0; s < nact; ++s) {
          const int u = p.ufirst + (r0 + s) * p.ustep;
          int seg = 0;
          while (seg + 1 < p.nseg && u >= p.ubeg[seg + 1]) ++seg;
          const int row0 = (u - p.ubeg[seg]) * p.upr;
          const uint32_t full = smem_base + OFF_FULL + slot * 8, empty = smem_base + OFF_EMPTY + slot * 8;
          mbar_wait(empty, par ^ 1);
          mbar_expect_tx(full, p.dual ? 2 * box_bytes : box_bytes);  // out-of-range rows / k-blocks are zero-filled
          __threadfence_block();
          sts32_volatile(smem_base + OFF_ISSUED, job + 1);
          const uint32_t dst = ring + slot * a.slot_bytes;
          tma_load_3d(dst, p.W[seg], full, 0, row0, ch * (KC / 64));
          if (p.dual) tma_load_3d(dst + box_bytes, p.W2, full, 0, row0 + p.row2, ch * (KC / 64));
          ++job;
          if (++slot == S) {
            slot = 0;
            par ^= 1;
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

// One warp accumulates its unit (8 weight rows, or 8 gate + 8 up rows) over one chunk for
// all token tiles.  MT == 0: up to 8 tokens, A fragments by predicated 32-bit LDS (rows
// 8..15 of the MMA tile are register zeros).  MT >= 1: 16*MT token rows by ldmatrix.
// The accumulation is a dependent HMMA chain per output tile (k ascending: the reference's
// order); measured on B200 the dependent HMMA.16816 latency is 20.7 cycles and the issue
// interval 8 cycles per SM sub-partition (tools/ubench/hmma_lat.cu), so the fragment loads
// (LDSM/LDS, ~33 cycles) are software-pipelined one group of k-steps ahead in registers and
// never sit on the chain.
template <int MT, bool DUAL>  // MT here = token tiles handled by ONE warp (0: the <= 8 token path)
struct Frags {
  static constexpr int MTT = MT == 0 ? 1 : MT;
  static constexpr int GS = MT <= 1 ? 4 : (MT == 2 ? 2 : 1);  // k16 steps per group
  uint32_t b[GS][DUAL ? 4 : 2];
  uint32_t a[GS][MTT][4];  // MT == 0: [1] and [3] (token rows 8..15) stay zero
};

// ASWZ (one 16-row token tile, GS == 4): the A operand is a TMA box [k-block of 64][16 rows][64] with the 128-byte
// swizzle (the layout of the weight tiles) instead of padded rows.
template <int MT, bool DUAL, bool ASWZ = false>
__device__ __forceinline__ void mma_chunk(float (&acc)[MT == 0 ? 1 : MT][4], float (&acc2)[MT == 0 ? 1 : MT][4],
                                          uint32_t slot_addr, uint32_t a_addr, int AS, int up_off, int nk16, int B,
                                          int lane) {
  using F = Frags<MT, DUAL>;
  constexpr int MTT = F::MTT, GS = F::GS;
  static_assert(!ASWZ || (MT == 1 && GS == 4), "swizzled A: one token tile per warp");
  // weight tile: [k-block of 64][row 0-7][64] bf16, 16-byte chunk c of row r stored at c ^ r;
  // ldmatrix lanes 0-7 -> rows @k, 8-15 -> rows @k+8 (x2); DUAL: lanes 16-31 -> the up tile
  const int br = lane & 7, bh = (lane >> 3) & 1;
  const uint32_t brow = slot_addr + br * 128 + (DUAL ? ((lane >> 4) & 1) * up_off : 0);
  uint32_t boff[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) boff[s] = (uint32_t)((((s << 1) + bh) ^ br) << 4);
  const bool row_valid = (lane >> 2) < B;
  const int arow = (lane & 7) + ((lane >> 3) & 1) * 8;  // ldmatrix.x4: lanes 0-15 rows @k, 16-31 rows @k+8
  const uint32_t a0 = MT == 0 ? a_addr + (lane >> 2) * AS + (lane & 3) * 4
                              : (ASWZ ? a_addr + arow * 128 : a_addr + arow * AS + (lane >> 4) * 16);
  uint32_t aoff[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) aoff[s] = (uint32_t)((((s << 1) + (lane >> 4)) ^ (arow & 7)) << 4);
  auto load_step = [&](F& f, int s, int j) {  // fragments of k-step j -> slot s of f
    const uint32_t baddr = brow + (uint32_t)(j >> 2) * 1024u + boff[(GS == 4) ? s : (j & 3)];
    if (DUAL)
      ldmatrix_x4(f.b[s][0], f.b[s][1], f.b[s][2], f.b[s][3], baddr);
    else
      ldmatrix_x2(f.b[s][0], f.b[s][1], baddr);
#pragma unroll
    for (int m = 0; m < MTT; ++m) {
      if (MT == 0) {
        if (row_valid) {  // invalid token rows keep the zeros they were initialised with
          f.a[s][m][0] = lds32(a0 + j * 32);
          f.a[s][m][2] = lds32(a0 + j * 32 + 16);
        }
      } else {
        if (ASWZ)
          ldmatrix_x4(f.a[s][m][0], f.a[s][m][1], f.a[s][m][2], f.a[s][m][3], a0 + (uint32_t)(j >> 2) * 2048u + aoff[s]);
        else
          ldmatrix_x4(f.a[s][m][0], f.a[s][m][1], f.a[s][m][2], f.a[s][m][3], a0 + m * 16 * AS + j * 32);
      }
    }
  };
  auto compute_step = [&](const F& f, int s) {
#pragma unroll
    for (int m = 0; m < MTT; ++m) {
      mma_bf16_16816(acc[m], f.a[s][m], f.b[s][0], f.b[s][1]);
      if (DUAL) mma_bf16_16816(acc2[m], f.a[s][m], f.b[s][2], f.b[s][3]);
    }
  };
  F f0, f1;
  if (MT == 0) {
    // Token rows 8..15 of the MMA tile are zero.  The zeros are made opaque to the compiler so
    // that each {a0, 0, a2, 0} operand quad stays allocated across the loop: with literal zeros
    // ptxas re-creates the quad every k-step on top of a just-written accumulator, which puts
    // HMMA -> MOV -> LDS -> HMMA (~60 cycles) on the dependent chain instead of 20.7.
#pragma unroll
    for (int s = 0; s < GS; ++s)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        asm volatile("mov.u32 %0, 0;" : "=r"(f0.a[s][0][i]));
        asm volatile("mov.u32 %0, 0;" : "=r"(f1.a[s][0][i]));
      }
  }
  // nk16 is a multiple of 4 (chunks are whole 64-wide k blocks), GS divides 4.  The HMMAs of a chain are dependent
  // (20.7 cycles apart) and a warp issues in order, so the fragment loads of the NEXT group are placed BETWEEN the
  // HMMAs of the current one -- step s of the other set right behind step s -- where they issue in the latency
  // shadow of the HMMA just issued.  (r01 issued a whole group of loads, then the group's HMMAs: both are asm
  // volatile and keep their order, so ~36 load / address instructions sat on the chain: 59 cycles per k-step at one
  // chain per warp.)
#pragma unroll
  for (int s = 0; s < GS; ++s) load_step(f0, s, s);
  int j0 = 0;
  for (; j0 + 2 * GS < nk16; j0 += 2 * GS) {  // at least three groups left: no per-step branches in the steady state
#pragma unroll
    for (int s = 0; s < GS; ++s) {
      compute_step(f0, s);
      load_step(f1, s, j0 + GS + s);
    }
#pragma unroll
    for (int s = 0; s < GS; ++s) {
      compute_step(f1, s);
      load_step(f0, s, j0 + 2 * GS + s);
    }
  }
  if (j0 + GS < nk16) {  // two groups left (f0 is loaded)
#pragma unroll
    for (int s = 0; s < GS; ++s) {
      compute_step(f0, s);
      load_step(f1, s, j0 + GS + s);
    }
#pragma unroll
    for (int s = 0; s < GS; ++s) compute_step(f1, s);
  } else {  // one group left
#pragma unroll
    for (int s = 0; s < GS; ++s) compute_step(f0, s);
  }
}

// a_src == nullptr: the A operand [B, K] is resident in the activation region (row stride
// (K+8)*2).  Otherwise it is streamed from global memory (row stride K) in KC-wide chunks
// through two buffers of the activation region, all consumer threads copying.
// lm_head output pair: bf16 logits (R10) + the running arg-max candidate of this lane
__device__ __forceinline__ void lm_store(const MegaArgs& a, int tok, int n, float v0, float v1, Best& bb) {
  const bf16 l0 = f2bf(v0), l1 = f2bf(v1);
  *reinterpret_cast<uint32_t*>(a.logits + (size_t)tok * a.V + n) = pack2(l0, l1);
  const float f0 = bf2f(l0), f1 = bf2f(l1);
  // take a candidate when it is finite-or-+inf and better in the reference's total order (cand_better).  A larger
  // value wins outright (also against the initial {-inf, -1}); only an exact tie needs the index comparison -- the
  // full comparator on every logit cost as many issue slots as the phase's HMMAs.
  if (f0 > bb.v || (f0 == bb.v && f0 > -CUDART_INF_F && cand_better(f0, n, bb.v, bb.i))) {
    bb.v = f0;
    bb.i = n;
  }
  if (f1 > bb.v || (f1 == bb.v && f1 > -CUDART_INF_F && cand_better(f1, n + 1, bb.v, bb.i))) {
    bb.v = f1;
    bb.i = n + 1;
  }
}

// One output pair (token tok, columns n, n+1) of a GEMM phase: the reference's rounding points
// (R3, R8, R9 of SURVEY 8a).
template <int MT>
__device__ __forceinline__ void epilogue_store(const MegaArgs& a, const Phase& p, int seg, int tok, int n, float v0, float v1,
                                               float w0, float w1, uint32_t res_old) {
  const int Dq = a.n_q * a.hd, Dkv = a.n_kv * a.hd;
  switch (p.kind) {
    case PH_QKV: {
      const int col = (seg == 0 ? 0 : (seg == 1 ? Dq : Dq + Dkv)) + n;
      *reinterpret_cast<uint32_t*>(a.qkv + (size_t)tok * (Dq + 2 * Dkv) + col) = pack2(f2bf(v0), f2bf(v1));
      break;
    }
    case PH_O:
    case PH_DOWN: {
      if (a.tp_size > 1) {  // fp32 partial sums of this rank -> every rank's exchange buffer (peer stores over NVLink)
        const size_t off = (((size_t)p.xpar * a.tp_size + a.tp_rank) * MEGA_TP_ROWS + tok) * a.H + n;
        for (int r = 0; r < a.tp_size; ++r) *reinterpret_cast<float2*>(a.tp_part[r] + off) = make_float2(v0, v1);
        break;
      }
      uint32_t* dst = reinterpret_cast<uint32_t*>(a.x + (size_t)tok * a.H + n);
      const uint32_t old = MT == 0 ? res_old : __ldcg(dst);
      const float y0 = bf2f(f2bf(v0)), y1 = bf2f(f2bf(v1));
      *dst = pack2(f2bf(__fadd_rn(lo2f(old), y0)), f2bf(__fadd_rn(hi2f(old), y1)));
      break;
    }
    case PH_GATEUP: {
      const float g0 = bf2f(f2bf(v0)), g1 = bf2f(f2bf(v1));
      const float u0 = bf2f(f2bf(w0)), u1 = bf2f(f2bf(w1));
      const float s0 = bf2f(f2bf(silu_ref(g0))), s1 = bf2f(f2bf(silu_ref(g1)));
      *reinterpret_cast<uint32_t*>(a.h + (size_t)tok * a.I + n) = pack2(f2bf(__fmul_rn(u0, s0)), f2bf(__fmul_rn(u1, s1)));
      break;
    }
    default:
      break;
  }
}

// TW = token tiles per warp: TW == MT -> one warp per unit (all tiles); TW < MT -> the MT/TW warps
// of a unit share its weight tile, each with its own token tiles (phases with one unit per
// CTA: more dependent HMMA chains run side by side instead of back to back in one warp).
template <int MT, int TW>
__device__ __forceinline__ void gemm_inner(const MegaArgs& a, const Phase& p, uint32_t smem_base, RingPos& base,
                                           const bf16* a_src, Best (&best)[MT == 0 ? 1 : MT][2],
                                           unsigned char* dbg_smem) {
  constexpr int MTT = MT == 0 ? 1 : MT;
  constexpr int TWW = TW == 0 ? 1 : TW;
  constexpr int MS = MTT / TWW;   // warps per unit
  constexpr int NU = NW / MS;     // units per round
  constexpr int BPAD = MT == 0 ? 8 : 16 * MT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int us = warp / MS, m0 = (warp % MS) * TWW;  // unit slot in the round, first token tile of this warp
  const uint32_t S = a.n_slots;
  const int KC = a.KC, RS = (KC + 8) * 2;  // RS: row stride of a streamed A chunk
  const uint32_t ring = smem_base + a.off_ring, act = smem_base + a.off_act;
  const bool stream = a_src != nullptr;
  const int AS = stream ? RS : (p.K + 8) * 2;  // activation row stride in shared memory (bytes)
  const int g = lane >> 2, c = lane & 3;
  const int n_c = p.n_c, nch = p.nch;
  RingPos rnd = base;  // first job of the current round

  auto a_chunk_load = [&](int ch) {  // stream mode: rows [0,B) x [ch*KC, +klen) -> buffer ch&1
    const int k0 = ch * KC, klen = min(KC, p.K - k0), per = klen >> 3;
    const uint32_t dst = act + (ch & 1) * (BPAD * RS);
    for (int i = threadIdx.x; i < a.B * per; i += NTC) {
      const int b = i / per, cc = i - b * per;
      cp_async16(dst + b * RS + cc * 16, a_src + (size_t)b * p.K + k0 + cc * 8);
    }
    cp_async_commit();
  };

  for (int r0 = 0; r0 < n_c; r0 += NU, rnd.advance(NU * nch, a.ph_round_slot[p.kind], a.ph_round_par[p.kind], S)) {
    const int nact = min(NU, n_c - r0);
    const bool has = us < nact;
    if (!stream && !has) break;  // resident A: idle warps need not take part
    int seg = 0, row0 = 0;
    if (has) {
      const int u = p.ufirst + (r0 + us) * p.ustep;
      while (seg + 1 < p.nseg && u >= p.ubeg[seg + 1]) ++seg;
      row0 = (u - p.ubeg[seg]) * p.upr;
    }
    float acc[TWW][4], acc2[TWW][4];
#pragma unroll
    for (int m = 0; m < TWW; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[m][i] = acc2[m][i] = 0.f;
    // <= 8 rows: the residual value of the epilogue is requested before the MMA chain, not after it
    uint32_t res_old = 0u;
    if (MT == 0 && has && (p.kind == PH_O || p.kind == PH_DOWN) && g < a.B && row0 + c * 2 < p.rows[seg])
      res_old = __ldcg(reinterpret_cast<const uint32_t*>(a.x + (size_t)g * a.H + row0 + c * 2));
    if (stream) a_chunk_load(0);
    // ring position of this warp's first job of the round; consecutive chunks are nact jobs apart
    RingPos me = rnd;
    me.step(us, S);
    for (int ch = 0; ch < nch; ++ch) {
      const int k0 = ch * KC;
      const int klen = min(KC, p.K - k0);
      if (stream) {
        if (ch + 1 < nch) {
          a_chunk_load(ch + 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        bar_consumers();
      }
      if (has) {
        const bool timed = a.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0;
        const long long tw0 = timed ? clock64() : 0;
        if (lds32_volatile(smem_base + OFF_ISSUED) <= me.job) {
          const long long t0 = clock64();
          while (lds32_volatile(smem_base + OFF_ISSUED) <= me.job)
            if (clock64() - t0 > SPIN_LIMIT) __trap();
        }
        const uint32_t slot = me.slot;
        mbar_wait(smem_base + OFF_FULL + slot * 8, me.par);
        const long long tw1 = timed ? clock64() : 0;
        const uint32_t a_addr = (stream ? act + (ch & 1) * (BPAD * RS) : act + k0 * 2) + m0 * 16 * AS;
        if (p.dual)
          mma_chunk<TW, true>(acc, acc2, ring + slot * a.slot_bytes, a_addr, AS, 8 * KC * 2, klen >> 4, a.B, lane);
        else
          mma_chunk<TW, false>(acc, acc2, ring + slot * a.slot_bytes, a_addr, AS, 0, klen >> 4, a.B, lane);
        __syncwarp();
        if (lane == 0) {
          if (MS == 1) {
            mbar_arrive(smem_base + OFF_EMPTY + slot * 8);
          } else {  // the last of the MS warps sharing the tile releases the slot
            uint32_t* rel = reinterpret_cast<uint32_t*>(dbg_smem - OFF_DBG + OFF_REL) + slot;
            if (atomicAdd(rel, 1u) == (uint32_t)(MS - 1)) {
              *rel = 0u;
              mbar_arrive(smem_base + OFF_EMPTY + slot * 8);
            }
          }
        }
        if (timed) {  // profiling: cycles this warp waited for weights / spent in the MMA loop
          unsigned long long* dbg = reinterpret_cast<unsigned long long*>(dbg_smem);
          dbg[p.kind * 2] += (unsigned long long)(tw1 - tw0);
          dbg[p.kind * 2 + 1] += (unsigned long long)(clock64() - tw1);
        }
        me.step(nact, S);
      }
      if (stream) bar_consumers();  // the buffer is refilled two chunks later
    }
    if (!has) continue;
    // epilogue: c0,c1 -> (token 16m + g, n = row0 + 2c + {0,1}); c2,c3 -> token 16m + g + 8
    const int n = row0 + c * 2;
    if (n >= p.rows[seg]) continue;  // rows are even everywhere (checked by the launcher)
    if (TWW >= 2 && p.kind != PH_LMHEAD) {
      // The epilogue of a layer phase runs ONCE per layer per warp, i.e. always out of a cold instruction cache (the
      // step walks ~8k instructions per layer), and cold straight-line code costs ~10 cycles per instruction: the
      // 8-fold unrolled SiLU / residual epilogue (8 x ~150 instructions) took 6 us per layer at batch 64 -- longer
      // than the phase's HMMAs.  The accumulators are parked in local memory (L1) and ONE copy of the epilogue code
      // loops over the (tile, half) pairs.
      float ev[TWW][8];
#pragma unroll
      for (int m = 0; m < TWW; ++m)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          ev[m][i] = acc[m][i];
          ev[m][4 + i] = acc2[m][i];
        }
#pragma unroll 1
      for (int mm = 0; mm < 2 * TWW; ++mm) {
        const int m = mm >> 1, hr = mm & 1;
        const int tok = (m0 + m) * 16 + g + hr * 8;
        if (tok >= a.B) continue;
        epilogue_store<MT>(a, p, seg, tok, n, ev[m][hr * 2], ev[m][hr * 2 + 1], ev[m][4 + hr * 2], ev[m][4 + hr * 2 + 1], res_old);
      }
      continue;
    }
#pragma unroll
    for (int m = 0; m < TWW; ++m)
#pragma unroll
      for (int hr = 0; hr < (MT == 0 ? 1 : 2); ++hr) {
        const int tok = (m0 + m) * 16 + g + hr * 8;
        if (tok >= a.B) continue;
        if (p.kind == PH_LMHEAD) {  // both tiles of the pair
          lm_store(a, tok, n, acc[m][hr * 2], acc[m][hr * 2 + 1], best[TW == MT ? m : 0][hr]);
          if (n + 8 < p.rows[seg]) lm_store(a, tok, n + 8, acc2[m][hr * 2], acc2[m][hr * 2 + 1], best[TW == MT ? m : 0][hr]);
        } else {
          epilogue_store<MT>(a, p, seg, tok, n, acc[m][hr * 2], acc[m][hr * 2 + 1], acc2[m][hr * 2], acc2[m][hr * 2 + 1],
                             res_old);
        }
      }
  }
  base.advance((uint32_t)(n_c * nch), a.ph_adv_slot[p.kind][p.cls], a.ph_adv_par[p.kind][p.cls], S);
}

// Tile-split GEMM phase (more than 16 rows, phases with at most one unit per CTA in the default mapping: QKV, O, DOWN of
// the 0.5B shape).  The default mapping gives such a phase ONE unit per CTA and spreads its token tiles over warps, so
// every CTA needs the activations of ALL rows: for down_proj that is the whole [64, I] operand (622 KB at I = 4864)
// copied from L2 by each of the 148 CTAs, 92 MB of L2 traffic per layer and 29 us (r01).  Here CTA c owns token tile
// c % T and the units c / T, c / T + G, ... (G = grid / T): warp w runs the HMMA chain of its unit for that ONE tile, the
// A operand is 16 rows -- resident for K = H (29 KB instead of 115 KB), streamed in KC-wide chunks through a 4-stage
// cp.async ring for down_proj (155 KB per CTA instead of 622 KB, one barrier per chunk) -- and the weights of a unit
// are fetched by the T CTAs of its group from L2 (HBM still sees them once).  Same arithmetic: one dependent chain per
// (unit, tile) with k ascending.
constexpr int TS_STAGES = 4;  // measured: 7 stages change nothing (the phase is not bound by the depth of the A stream)
template <int MT>
__device__ __forceinline__ void gemm_ts(const MegaArgs& a, const Phase& p, uint32_t smem_base, RingPos& base, const bf16* a_src,
                                        unsigned char* dbg_smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t S = a.n_slots;
  const int KC = a.KC, RS = (KC + 8) * 2;
  const uint32_t ring = smem_base + a.off_ring, act = smem_base + a.off_act;
  const bool stream = a_src != nullptr;
  const int AS = stream ? RS : (p.K + 8) * 2;
  const int g = lane >> 2, c = lane & 3;
  const int n_c = p.n_c, nch = p.nch;
  const int tok0 = p.tile * 16, nrows = min(16, a.B - tok0);
  RingPos rnd = base;
  // A stream by TMA (a.hmap: h as {64, rows, K / 64} in boxes of 16 rows x KC, 128-byte swizzle): ONE request per chunk
  // by one thread instead of ~4 cp.async per thread (measured: issuing those cost ~990 cycles per chunk, as much as
  // the chunk's HMMAs); the stage is a [k-block][16 rows][64] tile like the weight tiles (mma_chunk ASWZ).
  static_assert(TS_STAGES == 4, "stage = seq & 3");
  const bool tma = stream && a.hmap != nullptr;
  const uint32_t tsf = smem_base + OFF_TSF;
  const uint32_t stage0 = (act + 1023u) & ~1023u, STB = 16u * (uint32_t)KC * 2u;
  const bool issuer = threadIdx.x == (NW - 1) * 32;
  uint32_t aseq = tma ? lds32_volatile(smem_base + OFF_TSSEQ) : 0u;  // written behind a barrier of the previous tile-split phase
  auto a_issue = [&](int ch) {  // issuer thread: rows of this tile x [ch*KC, +KC) (zero-filled past K) -> stage (aseq + ch) & 3
    if (ch < nch) {
      const uint32_t st = (aseq + (uint32_t)ch) & 3u;
      mbar_expect_tx(tsf + st * 8, STB);
      tma_load_3d(stage0 + st * STB, a.hmap, tsf + st * 8, 0, tok0, ch * (KC / 64));
    }
  };
  // A stream by cp.async: piece i of a chunk is (row i / per, 16-byte column i % per); the threads from warp lw0 on walk their pieces
  // incrementally (one division per chunk).  All consumer threads copy (lw0 = 0).
  int lw0 = 0;  // first loader warp of the current round
  auto a_chunk_load = [&](int ch) {  // rows of this tile x [ch*KC, +klen) -> stage ch % TS_STAGES (an empty group past the end)
    const int lj = (int)threadIdx.x - lw0 * 32;
    if (ch < nch && lj >= 0) {
      const int NL = NTC - lw0 * 32;
      const int k0 = ch * KC, klen = min(KC, p.K - k0), per = klen >> 3;
      const uint32_t dst = act + (ch % TS_STAGES) * (16 * RS);
      int b = lj / per, cc = lj - b * per;
      while (b < nrows) {
        cp_async16(dst + b * RS + cc * 16, a_src + (size_t)(tok0 + b) * p.K + k0 + cc * 8);
        cc += NL;
        while (cc >= per) {
          cc -= per;
          ++b;
        }
      }
    }
    cp_async_commit();
  };
  for (int r0 = 0; r0 < n_c; r0 += NW, rnd.advance(NW * nch, a.ph_round_slot[p.kind], a.ph_round_par[p.kind], S)) {
    const int nact = min(NW, n_c - r0);
    const bool has = warp < nact;
    if (!stream && !has) break;
    int seg = 0, row0 = 0;
    if (has) {
      const int u = p.ufirst + (r0 + warp) * p.ustep;
      while (seg + 1 < p.nseg && u >= p.ubeg[seg + 1]) ++seg;
      row0 = (u - p.ubeg[seg]) * p.upr;
    }
    float acc[1][4] = {{0.f, 0.f, 0.f, 0.f}}, acc2[1][4] = {{0.f, 0.f, 0.f, 0.f}};
    // residual words of this lane's two outputs (O / down without tensor parallel): requested now, used in the epilogue
    uint32_t res_old[2] = {0u, 0u};
    const bool res_pre = has && a.tp_size <= 1 && (p.kind == PH_O || p.kind == PH_DOWN) && row0 + c * 2 < p.rows[seg];
    if (res_pre) {
#pragma unroll
      for (int hr = 0; hr < 2; ++hr) {
        const int tok = tok0 + g + hr * 8;
        if (tok < a.B) res_old[hr] = __ldcg(reinterpret_cast<const uint32_t*>(a.x + (size_t)tok * a.H + row0 + c * 2));
      }
    }
    lw0 = 0;  // measured (B200, batch 64): leaving the copies to the warps without a unit is slower (20.0 vs 17.0 us per layer)
    // ring mode (the CTA's units fit one round and leave the issuer's warp free -- the 0.5B / 1.5B shapes): no CTA
    // barrier per chunk; the issuer runs ahead, a stage is re-armed when the nact warps with a unit have read it
    // (fill f of a stage waits for phase f - 1 of its "read" barrier, whose count is nact: set at kernel start)
    const bool ring_mode = tma && n_c <= NW - 1;
    if (tma) {
      if (issuer) {
        // h was written by other SMs through the generic proxy (published by the grid barrier), the stages were
        // read / written through the generic proxy: order both before the async-proxy copies
        asm volatile("fence.proxy.async;" ::: "memory");
        if (ring_mode) {
          for (int ch = 0; ch < nch; ++ch) {
            const uint32_t q = aseq + (uint32_t)ch;
            if (q >= (uint32_t)TS_STAGES) mbar_wait(smem_base + OFF_TSE + (q & 3u) * 8, ((q >> 2) & 1u) ^ 1u);
            a_issue(ch);
          }
        } else {
          for (int s = 0; s < TS_STAGES - 1; ++s) a_issue(s);
        }
      }
    } else if (stream) {
      for (int s = 0; s < TS_STAGES - 1; ++s) a_chunk_load(s);
    }
    RingPos me = rnd;
    me.step(warp, S);
    for (int ch = 0; ch < nch; ++ch) {
      const int k0 = ch * KC, klen = min(KC, p.K - k0);
      if (tma) {
        if (!ring_mode) {
          if (ch > 0) bar_consumers();  // everybody is done with chunk ch - 1: its stage takes chunk ch + 3
          if (issuer) a_issue(ch + TS_STAGES - 1);
        }
        if (has) mbar_wait(tsf + ((aseq + (uint32_t)ch) & 3u) * 8, ((aseq + (uint32_t)ch) >> 2) & 1u);
      } else if (stream) {
        cp_async_wait<TS_STAGES - 2>();  // this thread's part of chunk ch has landed
        bar_consumers();                 // ... everybody's has, and everybody is done with chunk ch - 1
        // measured: issuing these ~4 cp.async per thread costs ~990 cycles per chunk (35 per LDGSTS) -- as much as the
        // chunk's HMMAs; a TMA box per chunk (as the weights use) is the fix that is still open
        a_chunk_load(ch + TS_STAGES - 1);  // into the stage chunk ch - 1 used
      }
      if (has) {
        const bool timed = a.prof != nullptr && threadIdx.x == 0 && blockIdx.x == 0;
        const long long tw0 = timed ? clock64() : 0;
        if (lds32_volatile(smem_base + OFF_ISSUED) <= me.job) {
          const long long t0 = clock64();
          while (lds32_volatile(smem_base + OFF_ISSUED) <= me.job)
            if (clock64() - t0 > SPIN_LIMIT) __trap();
        }
        const uint32_t slot = me.slot;
        mbar_wait(smem_base + OFF_FULL + slot * 8, me.par);
        const long long tw1 = timed ? clock64() : 0;
        const uint32_t a_addr = stream ? act + (ch % TS_STAGES) * (16 * RS) : act + k0 * 2;
        if (tma)
          mma_chunk<1, false, true>(acc, acc2, ring + slot * a.slot_bytes, stage0 + ((aseq + (uint32_t)ch) & 3u) * STB, AS, 0, klen >> 4, a.B, lane);
        else
          mma_chunk<1, false>(acc, acc2, ring + slot * a.slot_bytes, a_addr, AS, 0, klen >> 4, a.B, lane);
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(smem_base + OFF_EMPTY + slot * 8);
          if (ring_mode) mbar_arrive(smem_base + OFF_TSE + ((aseq + (uint32_t)ch) & 3u) * 8);
        }
        if (timed) {
          unsigned long long* dbg = reinterpret_cast<unsigned long long*>(dbg_smem);
          dbg[p.kind * 2] += (unsigned long long)(tw1 - tw0);
          dbg[p.kind * 2 + 1] += (unsigned long long)(clock64() - tw1);
        }
        me.step(nact, S);
      }
    }
    if (tma) {
      aseq += (uint32_t)nch;
      bar_consumers();  // the stages are refilled by the next round / the next phase's row load
    } else if (stream) {
      cp_async_wait<0>();
      bar_consumers();  // the stages are refilled by the next round / the next phase's row load
    }
    if (!has) continue;
    const int n = row0 + c * 2;
    if (n >= p.rows[seg]) continue;
#pragma unroll
    for (int hr = 0; hr < 2; ++hr) {
      const int tok = tok0 + g + hr * 8;
      if (tok < a.B) {
        if (res_pre)
          epilogue_store<0>(a, p, seg, tok, n, acc[0][hr * 2], acc[0][hr * 2 + 1], 0.f, 0.f, res_old[hr]);  // <0>: residual word passed in
        else
          epilogue_store<MT>(a, p, seg, tok, n, acc[0][hr * 2], acc[0][hr * 2 + 1], 0.f, 0.f, 0u);
      }
    }
  }
  if (tma && issuer) sts32_volatile(smem_base + OFF_TSSEQ, aseq);  // read by everybody behind the next grid barrier
  if (n_c > 0) base.advance((uint32_t)(n_c * nch), a.ph_adv_slot[p.kind][p.cls], a.ph_adv_par[p.kind][p.cls], S);
}

template <int MT>
__device__ __forceinline__ void gemm_phase(const MegaArgs& a, const Phase& p, uint32_t smem_base, RingPos& base,
                                           const bf16* a_src, Best (&best)[MT == 0 ? 1 : MT][2],
                                           unsigned char* dbg_smem) {
  if (MT >= 2 && p.ts)
    gemm_ts<MT>(a, p, smem_base, base, a_src, dbg_smem);  // one token tile per CTA
  else if (MT >= 2 && p.nu != NW)
    gemm_inner<MT, 1>(a, p, smem_base, base, a_src, best, dbg_smem);  // one token tile per warp
  else
    gemm_inner<MT, MT>(a, p, smem_base, base, a_src, best, dbg_smem);
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

// norm weight [H] -> shared memory buffer `which` (completes with the next cp.async wait)
__device__ __forceinline__ void prefetch_norm_w(const MegaArgs& a, uint32_t smem_base, int which, const bf16* w) {
  const uint32_t dst = smem_base + OFF_WNORM + which * a.H * 2;
  for (int i = threadIdx.x; i < (a.H >> 3); i += NTC) cp_async16(dst + i * 16, w + i * 8);
  cp_async_commit();
}

// RMSNorm of the rows in the activation region, in place (normalization.cu:5-26): the sum
// of squares is the reference's sequential FFMA chain, one thread per row.
__device__ __forceinline__ void rmsnorm_rows(const MegaArgs& a, unsigned char* smem, int B, int H, int which) {
  const int AS = (H + 8) * 2;
  unsigned char* act = smem + a.off_act;
  const bf16* w = reinterpret_cast<const bf16*>(smem + OFF_WNORM + which * H * 2);
  float* rms_s = reinterpret_cast<float*>(smem + OFF_RMS);
  if (threadIdx.x < B) {
    const uint4* row = reinterpret_cast<const uint4*>(act + threadIdx.x * AS);
    float sum = 0.f;
    const int nv = H >> 3;
    // the FFMA chain is 4 cycles per element; keep the next 4 vectors in flight so the
    // shared-memory latency never sits on the chain
    uint4 cur[4], nxt[4];  // H % 32 == 0 (launcher) -> nv % 4 == 0
#pragma unroll
    for (int j = 0; j < 4; ++j) cur[j] = row[j];
    for (int i = 0; i < nv; i += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) nxt[j] = row[min(i + 4 + j, nv - 1)];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint4 v = cur[j];
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
#pragma unroll
      for (int j = 0; j < 4; ++j) cur[j] = nxt[j];
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

// FAST numerics (tolerance, not bit-exact), <= 8 rows: every unit is shared by ALL consumer
// warps -- warp w takes the 64-wide k-groups g with g % NW == w of every weight tile -- so the
// dependent HMMA chain per warp is 1/NW of the reference's; partial sums meet in shared memory
// (fixed order, deterministic).  Units run in rounds of <= FAST_U so one reduction serves several.
constexpr int FAST_U = 4;
__device__ __forceinline__ void gemm_phase_fast(const MegaArgs& a, const Phase& p, uint32_t smem_base, RingPos& base,
                                                Best (&best)[1][2], unsigned char* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t S = a.n_slots;
  const int KC = a.KC;
  const uint32_t ring = smem_base + a.off_ring, act = smem_base + a.off_act;
  const int AS = (p.K + 8) * 2;
  const int g = lane >> 2, c = lane & 3;
  const int n_c = p.n_c, nch = p.nch, U = p.nu;
  float* red = reinterpret_cast<float*>(smem + a.off_red);  // [FAST_U][NW][32][8]
  RingPos me = base;
  for (int r0 = 0; r0 < n_c; r0 += U) {
    const int nact = min(U, n_c - r0);
    float acc[FAST_U][1][4], acc2[FAST_U][1][4];
#pragma unroll
    for (int u = 0; u < FAST_U; ++u)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[u][0][i] = acc2[u][0][i] = 0.f;
    for (int ch = 0; ch < nch; ++ch) {
      const int k0 = ch * KC;
      const int ngr = min(KC, p.K - k0) >> 6;
#pragma unroll
      for (int u = 0; u < FAST_U; ++u) {
        if (u < nact) {
          if (lds32_volatile(smem_base + OFF_ISSUED) <= me.job) {
            const long long t0 = clock64();
            while (lds32_volatile(smem_base + OFF_ISSUED) <= me.job)
              if (clock64() - t0 > SPIN_LIMIT) __trap();
          }
          mbar_wait(smem_base + OFF_FULL + me.slot * 8, me.par);
          const uint32_t slot_addr = ring + me.slot * a.slot_bytes;
          for (int gi = warp; gi < ngr; gi += NW) {
            if (p.dual)
              mma_chunk<0, true>(acc[u], acc2[u], slot_addr + gi * 1024, act + (k0 + gi * 64) * 2, AS, 8 * KC * 2, 4, a.B, lane);
            else
              mma_chunk<0, false>(acc[u], acc2[u], slot_addr + gi * 1024, act + (k0 + gi * 64) * 2, AS, 0, 4, a.B, lane);
          }
          __syncwarp();
          if (lane == 0) {  // the last warp through releases the slot
            uint32_t* rel = reinterpret_cast<uint32_t*>(smem + OFF_REL) + me.slot;
            if (atomicAdd(rel, 1u) == (uint32_t)(NW - 1)) {
              *rel = 0u;
              mbar_arrive(smem_base + OFF_EMPTY + me.slot * 8);
            }
          }
          me.step(1, S);
        }
      }
    }
    // partial sums -> shared memory; warp u finishes unit u of the round
#pragma unroll
    for (int u = 0; u < FAST_U; ++u) {
      if (u < nact) {
        float4* dst = reinterpret_cast<float4*>(red + ((size_t)(u * NW + warp) * 32 + lane) * 8);
        dst[0] = make_float4(acc[u][0][0], acc[u][0][1], acc[u][0][2], acc[u][0][3]);
        dst[1] = make_float4(acc2[u][0][0], acc2[u][0][1], acc2[u][0][2], acc2[u][0][3]);
      }
    }
    bar_consumers();
    if (warp < nact) {
      float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
      for (int w = 0; w < NW; ++w) {
        const float4* src = reinterpret_cast<const float4*>(red + ((size_t)(warp * NW + w) * 32 + lane) * 8);
        const float4 x = src[0], y = src[1];
        s0 += x.x;
        s1 += x.y;
        t0 += y.x;
        t1 += y.y;
      }
      const int u = blockIdx.x + (r0 + warp) * gridDim.x;
      int seg = 0;
      while (seg + 1 < p.nseg && u >= p.ubeg[seg + 1]) ++seg;
      const int row0 = (u - p.ubeg[seg]) * p.upr;
      const int n = row0 + c * 2;
      if (g < a.B && n < p.rows[seg]) {
        uint32_t res_old = 0u;
        if (p.kind == PH_O || p.kind == PH_DOWN) res_old = __ldcg(reinterpret_cast<const uint32_t*>(a.x + (size_t)g * a.H + n));
        if (p.kind == PH_LMHEAD) {
          lm_store(a, g, n, s0, s1, best[0][0]);
          if (n + 8 < p.rows[seg]) lm_store(a, g, n + 8, t0, t1, best[0][0]);
        } else {
          epilogue_store<0>(a, p, seg, g, n, s0, s1, t0, t1, res_old);
        }
      }
    }
    bar_consumers();  // red[] is rewritten by the next round
  }
  base = me;
}

// FAST numerics RMSNorm: the sum of squares is reduced in parallel (one warp per row)
__device__ __forceinline__ void rmsnorm_rows_fast(const MegaArgs& a, unsigned char* smem, int B, int H, int which) {
  const int AS = (H + 8) * 2;
  unsigned char* act = smem + a.off_act;
  const bf16* w = reinterpret_cast<const bf16*>(smem + OFF_WNORM + which * H * 2);
  float* rms_s = reinterpret_cast<float*>(smem + OFF_RMS);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < B; b += NW) {
    const uint4* row = reinterpret_cast<const uint4*>(act + b * AS);
    float sum = 0.f;
    for (int i = lane; i < (H >> 3); i += 32) {
      const uint4 v = row[i];
      float f;
      f = lo2f(v.x); sum += f * f;
      f = hi2f(v.x); sum += f * f;
      f = lo2f(v.y); sum += f * f;
      f = hi2f(v.y); sum += f * f;
      f = lo2f(v.z); sum += f * f;
      f = hi2f(v.z); sum += f * f;
      f = lo2f(v.w); sum += f * f;
      f = hi2f(v.w); sum += f * f;
    }
    sum = warp_sum(sum);
    if (lane == 0) rms_s[b] = __fsqrt_rn(__fadd_rn(__fdiv_rn(sum, (float)H), 1e-04f));
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

// the same tree over a K row that was widened to fp32 once (shared by all heads of the task),
// evaluated with packed fp32x2 instructions (sm_100 FMUL2 / FADD2: two IEEE fp32 operations per
// instruction, each lane bit-identical to the scalar op).  kf[j] = {k[2j], k[2j+1]}.
template <int HD>
__device__ __forceinline__ float dot_tree_f(const float* __restrict__ q, const float2 (&kf)[HD / 2]) {
  float2 d[HD / 4];
#pragma unroll
  for (int j = 0; j < HD / 4; j += 2) {  // stride HD/2: d[j] = p[2j..2j+1] + p[2j+HD/2..]
    const float4 qa = *reinterpret_cast<const float4*>(q + 2 * j);
    const float4 qb = *reinterpret_cast<const float4*>(q + 2 * j + HD / 2);
    d[j] = __fadd2_rn(__fmul2_rn(make_float2(qa.x, qa.y), kf[j]), __fmul2_rn(make_float2(qb.x, qb.y), kf[j + HD / 4]));
    d[j + 1] = __fadd2_rn(__fmul2_rn(make_float2(qa.z, qa.w), kf[j + 1]), __fmul2_rn(make_float2(qb.z, qb.w), kf[j + 1 + HD / 4]));
  }
#pragma unroll
  for (int s = HD / 8; s >= 1; s >>= 1)  // strides HD/4 ... 2 elements
#pragma unroll
    for (int j = 0; j < s; ++j) d[j] = __fadd2_rn(d[j], d[j + s]);
  return __fadd_rn(d[0].x, d[0].y);  // stride 1
}

// the same with the query operands travelling through a rolling register window: pair u of a head is the quads
// q[4u..4u+3] and q[HD/2+4u..]; when pair u has been consumed its slot is refilled with pair u+4 (of the NEXT head
// once this one runs out), so every broadcast LDS.128 is issued ~4 pairs (two dozen FP instructions) before its use.
// With 7 warps per SM nothing else hides the shared-memory latency, and holding a whole head (64 registers) next to
// the fp32 K row (64) spills.  (Measured dead end: the query head as packed bf16 -- half the shared-memory bytes, two
// ALU instructions per pair to widen -- makes the loop 20 % SLOWER: it is bound by issue slots, not by the LSU.)
template <int HD>
__device__ __forceinline__ float dot_tree_w(float4 (&wa)[4], float4 (&wb)[4], const float2 (&kf)[HD / 2], const float* __restrict__ q,
                                            const float* __restrict__ qnext) {
  constexpr int NPAIR = HD / 8;
  float2 d[HD / 4];
#pragma unroll
  for (int u = 0; u < NPAIR; ++u) {
    const float4 qa = wa[u & 3], qb = wb[u & 3];
    const float* src = u + 4 < NPAIR ? q + 4 * (u + 4) : qnext + 4 * (u + 4 - NPAIR);
    wa[u & 3] = *reinterpret_cast<const float4*>(src);
    wb[u & 3] = *reinterpret_cast<const float4*>(src + HD / 2);
    d[2 * u] = __fadd2_rn(__fmul2_rn(make_float2(qa.x, qa.y), kf[2 * u]), __fmul2_rn(make_float2(qb.x, qb.y), kf[2 * u + HD / 4]));
    d[2 * u + 1] = __fadd2_rn(__fmul2_rn(make_float2(qa.z, qa.w), kf[2 * u + 1]), __fmul2_rn(make_float2(qb.z, qb.w), kf[2 * u + 1 + HD / 4]));
  }
#pragma unroll
  for (int s = HD / 8; s >= 1; s >>= 1)
#pragma unroll
    for (int j = 0; j < s; ++j) d[j] = __fadd2_rn(d[j], d[j + s]);
  return __fadd_rn(d[0].x, d[0].y);
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

// Cached K/V of a layer do not depend on anything computed in this step (only the new position does,
// and that one travels through shared memory), so the chunks of this CTA's attention task(s) are
// requested into L2 long before the attention phase reads them: one cp.async.bulk.prefetch per
// (page, K|V) chunk, issued while the down_proj / QKV phases keep the SMs busy and HBM idle.  The
// scores loop and the V tile loader then see L2 latency instead of DRAM latency.
template <int NP>
__device__ __forceinline__ void attention_prefetch_l2(const MegaArgs& a, int layer) {
  constexpr int HD = 64 * NP;
  const int Gq = a.n_q / a.n_kv;
  const bool per_head = a.B * a.n_q <= (int)gridDim.x;
  const int ntask = per_head ? a.B * a.n_q : a.B * a.n_kv;
  const int psz = a.kv.page_size;
  const uint32_t bytes = (uint32_t)psz * HD * 2;
  if (bytes & 15) return;
  for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
    int b, kvh;
    if (per_head) {
      b = task / a.n_q;
      const int h0 = task - b * a.n_q;
      kvh = h0 / Gq;
      if (h0 != kvh * Gq) continue;  // one request per kv head
    } else {
      b = task / a.n_kv;
      kvh = task - b * a.n_kv;
    }
    const int ps = a.pos[b];
    if (ps <= a.attn_kstg) continue;  // short contexts are staged through cp.async next to the QKV GEMM
    const int n_pages = (ps + psz - 1) / psz;
    const int* bt = a.block_table + (size_t)a.slot[b] * a.max_pages;
    for (int i = threadIdx.x; i < 2 * n_pages; i += NTC) {  // K chunks first: the scores loop reads them first
      const int kv = i >= n_pages ? 1 : 0, pi = kv ? i - n_pages : i;
      const bf16* src = a.kv.chunk(bt[pi], layer, kv, kvh);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
    }
  }
}

// Tasks: (row, q head) while they fit one wave of CTAs, else (row, kv head) with the whole
// query group sharing the K/V stream.  q/k-norm + RoPE + the KV store of the new position
// are done here (replaces 2x qkNorm, 2x RoPE, kv_copy_layer_to_cache_decode).
// stage 0 (before the grid barrier that publishes this layer's q/k/v): everything of the CTA's
// first task that does not depend on them -- position, page ids, the first KSTG cached K rows,
// V tile 0, norm weights and the RoPE row -- is requested into shared memory, so that the
// dependent part after the barrier (stage 1) starts from on-chip data.
template <int NP>
__device__ __forceinline__ void attention_phase(const MegaArgs& a, int layer, unsigned char* smem, int stage) {
  constexpr int HD = 64 * NP;
  constexpr int KROW = HD * 2 + 16;  // staged K rows are padded: conflict-free 16-byte reads, one row per thread
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Gq = a.n_q / a.n_kv;
  const int Dq = a.n_q * HD, Dkv = a.n_kv * HD, QKV = Dq + 2 * Dkv;
  const bool per_head = a.B * a.n_q <= (int)gridDim.x;
  const int hs = per_head ? 1 : Gq;
  const int ntask = per_head ? a.B * a.n_q : a.B * a.n_kv;
  const int tmax = (a.max_kv_len + 3) & ~3;
  const int psz = a.kv.page_size;
  const int psz_shift = (psz & (psz - 1)) == 0 ? __ffs(psz) - 1 : -1;
  const MegaLayer& w = reinterpret_cast<const MegaLayer*>(smem + OFF_LAYERS)[layer];

  // staged mode: stage 0 runs next to the QKV GEMM, so the attention areas start behind its activation rows
  float* q_s = reinterpret_cast<float*>(smem + a.off_act + a.attn_off);  // [hs][HD] fp32
  bf16* knew = reinterpret_cast<bf16*>(q_s + hs * HD);              // [HD] bf16 (16-byte aligned)
  bf16* vnew = knew + HD;
  float* score = reinterpret_cast<float*>(vnew + HD);               // [hs][tmax]
  int* pages = reinterpret_cast<int*>(score + hs * tmax);
  const int n_pages_max = a.max_kv_len / psz + 1;
  unsigned char* vbuf = reinterpret_cast<unsigned char*>(pages) + ((n_pages_max * 4 + 15) & ~15);  // [2][VT][HD] bf16
  const uint32_t vbuf_u32 = smem_u32(vbuf);
  // stage-0 area (per-head tasks only): staged K rows, q/k norm weights, RoPE row, position
  unsigned char* kst = vbuf + 2 * VT * HD * 2;                       // [kstg][KROW]
  const int kstg = a.attn_kstg;
  bf16* wq_s = reinterpret_cast<bf16*>(kst + kstg * KROW);           // [HD]
  bf16* wk_s = wq_s + HD;
  float* cos_s = reinterpret_cast<float*>(wk_s + HD);                // [HD/2]
  float* sin_s = cos_s + HD / 2;
  int* pre = reinterpret_cast<int*>(sin_s + HD / 2);                 // [0] = position of the staged task
  const bool staged_mode = kstg > 0;

  if (stage == 0) {
    // executed by warps 1.. only (warp 0 owns the CTA's QKV unit and goes straight to its MMA chain)
    if (!staged_mode || (int)blockIdx.x >= ntask || warp == 0) return;
    const int tid = threadIdx.x - 32;
    constexpr int NT0 = NTC - 32;
    const int task = blockIdx.x;
    const int b = task / a.n_q, h0 = task - b * a.n_q, kvh = h0 / Gq;  // staged mode = one q head per task
    const int ps = a.pos[b];
    const int* bt = a.block_table + (size_t)a.slot[b] * a.max_pages;
    const int n_pages = ps / psz + 1;
    for (int i = tid; i < n_pages; i += NT0) pages[i] = bt[i];
    if (tid == 0) pre[0] = ps;
    const int nk = min(ps, kstg);
    const uint32_t kst_u32 = smem_u32(kst);
    for (int idx = tid; idx < nk * (HD / 8); idx += NT0) {
      const int r = idx / (HD / 8), cc = idx - r * (HD / 8);
      const int pi = psz_shift >= 0 ? (r >> psz_shift) : r / psz;
      cp_async16(kst_u32 + r * KROW + cc * 16, a.kv.chunk(bt[pi], layer, 0, kvh) + (size_t)(r - pi * psz) * HD + cc * 8);
    }
    if (tid < HD / 8) {
      if (w.q_norm) cp_async16(smem_u32(wq_s) + tid * 16, w.q_norm + tid * 8);
      if (w.k_norm) cp_async16(smem_u32(wk_s) + tid * 16, w.k_norm + tid * 8);
      cp_async16(smem_u32(cos_s) + tid * 16, a.cos_t + (size_t)ps * 32 * NP + tid * 4);
      cp_async16(smem_u32(sin_s) + tid * 16, a.sin_t + (size_t)ps * 32 * NP + tid * 4);
    }
    cp_async_commit();
    {  // V tile 0 (page ids straight from the block table: pages[] is still being written)
      const int tn = min(VT, ps);
      for (int idx = tid; idx < tn * (HD / 8); idx += NT0) {
        const int r = idx / (HD / 8), cc = idx - r * (HD / 8);
        const int pi = psz_shift >= 0 ? (r >> psz_shift) : r / psz;
        cp_async16(vbuf_u32 + r * (HD * 2) + cc * 16, a.kv.chunk(bt[pi], layer, 1, kvh) + (size_t)(r - pi * psz) * HD + cc * 8);
      }
      cp_async_commit();
    }
    return;
  }

  unsigned long long* dbg = reinterpret_cast<unsigned long long*>(smem + OFF_DBG);
  const bool timed = a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
    long long tq = timed ? clock64() : 0;
    auto lap = [&](int slot) {
      if (timed) {
        const long long t = clock64();
        dbg[slot] += (unsigned long long)(t - tq);
        tq = t;
      }
    };
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
    const bool staged = staged_mode && task == (int)blockIdx.x;  // stage 0 ran for this task
    const int ps = staged ? pre[0] : a.pos[b];
    const int n_pages = ps / psz + 1;
    const int* bt = nullptr;
    if (!staged) {
      bt = a.block_table + (size_t)a.slot[b] * a.max_pages;
      for (int i = threadIdx.x; i < n_pages; i += NTC) pages[i] = bt[i];
    }
    auto v_tile_load = [&](int it) {  // cached positions [it*VT, min(ps, it*VT+VT)) of V -> vbuf[it&1]
      const int t0 = it * VT, tn = min(VT, ps - t0);
      const uint32_t dst = vbuf_u32 + (it & 1) * (VT * HD * 2);
      const int* pg = it == 0 ? bt : pages;  // tile 0 is issued while pages[] is still being filled
      for (int idx = threadIdx.x; idx < tn * (HD / 8); idx += NTC) {
        const int r = idx / (HD / 8), cc = idx - r * (HD / 8);
        const int k = t0 + r;
        const int pi = psz_shift >= 0 ? (k >> psz_shift) : k / psz;
        const int po = k - pi * psz;
        cp_async16(dst + r * (HD * 2) + cc * 16, a.kv.chunk(pg[pi], layer, 1, kvh) + (size_t)po * HD + cc * 8);
      }
      cp_async_commit();
    };
    if (staged)
      cp_async_wait<1>();  // staged K rows / weights / RoPE row have landed (V tile 0 may still be in flight)
    else
      v_tile_load(0);  // overlaps q/k-norm, scores and softmax
    const bf16* row = a.qkv + (size_t)b * QKV;
    const float* cos_row = staged ? cos_s : a.cos_t + (size_t)ps * 32 * NP;
    const float* sin_row = staged ? sin_s : a.sin_t + (size_t)ps * 32 * NP;
    const bf16* wqn = staged ? wq_s : w.q_norm;
    const bf16* wkn = staged ? wk_s : w.k_norm;
    const int new_page = staged ? pages[ps / psz] : bt[ps / psz], new_off = ps % psz;
    if (staged) bar_consumers();  // cp.async data of other threads (weights, RoPE row) is visible after this
    for (int r = warp; r < hs + 2; r += NW) {
      float v[NP][2];
      if (r < hs) {
        head_load_cg<NP>(v, row + (size_t)(h0 + r) * HD, lane);
        if (w.q_norm) head_norm<NP>(v, wqn, lane);
        head_rope<NP>(v, cos_row, sin_row, lane);
#pragma unroll
        for (int p = 0; p < NP; ++p)
          *reinterpret_cast<float2*>(q_s + r * HD + 64 * p + 2 * lane) = make_float2(v[p][0], v[p][1]);
      } else if (r == hs) {
        head_load_cg<NP>(v, row + Dq + (size_t)kvh * HD, lane);
        if (w.k_norm) head_norm<NP>(v, wkn, lane);
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
    lap(10);

    // scores: one thread per cache position, all heads of the task
    const float den = __fsqrt_rn((float)HD);
    auto k_row = [&](int k) -> const uint4* {
      return (k == ps) ? reinterpret_cast<const uint4*>(knew)
             : (staged && k < kstg) ? reinterpret_cast<const uint4*>(kst + k * KROW)
                       : reinterpret_cast<const uint4*>(
                             a.kv.chunk(pages[psz_shift >= 0 ? (k >> psz_shift) : k / psz], layer, 0, kvh) +
                             (size_t)(psz_shift >= 0 ? (k & (psz - 1)) : k % psz) * HD);
    };
    // (measured: requesting the thread's next K row into registers ahead of the product trees does not
    // shorten this loop -- it is bound by the broadcast loads of q and FP32 issue -- and costs spills)
    for (int k = threadIdx.x; k <= ps; k += NTC) {
      const uint4* kp = k_row(k);
      float2 kf[HD / 2];
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        const uint4 t = kp[i];
        kf[4 * i] = make_float2(lo2f(t.x), hi2f(t.x));
        kf[4 * i + 1] = make_float2(lo2f(t.y), hi2f(t.y));
        kf[4 * i + 2] = make_float2(lo2f(t.z), hi2f(t.z));
        kf[4 * i + 3] = make_float2(lo2f(t.w), hi2f(t.w));
      }
      // dot / sqrtf(hd): for hd = 64 the divisor is exactly 8, and x / 8 == x * 0.125 bit for bit
      for (int i = 0; i < hs; ++i) {
        const float dot = dot_tree_f<HD>(q_s + i * HD, kf);
        score[i * tmax + k] = HD == 64 ? __fmul_rn(dot, 0.125f) : __fdiv_rn(dot, den);
      }
    }
    bar_consumers();
    lap(11);

    // softmax: one warp per head (self_attension.cu:94-107)
    for (int i = warp; i < hs; i += NW) {
      float* s = score + i * tmax;
      float m = -1e9f;
      for (int k = lane; k <= ps; k += 32) m = fmaxf(m, s[k]);
      m = warp_max(m);
      for (int k = lane; k <= ps; k += 32) s[k] = expf(__fsub_rn(s[k], m));
      __syncwarp();
      float sum = 0.f;
      if (lane == 0) {  // the reference's sequential sum: keep 8 values in flight ahead of the FADD chain
        const int n = ps + 1, n8 = n & ~7;
        float4 c0, c1, n0, n1;
        if (n8) {
          c0 = *reinterpret_cast<const float4*>(s);
          c1 = *reinterpret_cast<const float4*>(s + 4);
        }
        for (int k = 0; k < n8; k += 8) {
          const int kn = k + 8 < n8 ? k + 8 : k;
          n0 = *reinterpret_cast<const float4*>(s + kn);
          n1 = *reinterpret_cast<const float4*>(s + kn + 4);
          sum = __fadd_rn(sum, c0.x);
          sum = __fadd_rn(sum, c0.y);
          sum = __fadd_rn(sum, c0.z);
          sum = __fadd_rn(sum, c0.w);
          sum = __fadd_rn(sum, c1.x);
          sum = __fadd_rn(sum, c1.y);
          sum = __fadd_rn(sum, c1.z);
          sum = __fadd_rn(sum, c1.w);
          c0 = n0;
          c1 = n1;
        }
        for (int k = n8; k < n; ++k) sum = __fadd_rn(sum, s[k]);
      }
      sum = __shfl_sync(0xffffffffu, sum, 0);
      for (int k = lane; k <= ps; k += 32) s[k] = __fdiv_rn(s[k], sum);
    }
    lap(12);
    // PV (self_attension.cu:112-137): V streams through shared memory in tiles of VT cached
    // positions (all consumer threads copy, double buffered); warp i keeps the sequential
    // fma chain of head i, lane l owning dims 2l, 2l+1 (+64p).
    float o[2][NP][2];  // heads i = warp and warp + NW of the task (hs <= 2*NW)
#pragma unroll
    for (int q2 = 0; q2 < 2; ++q2)
#pragma unroll
      for (int p = 0; p < NP; ++p) o[q2][p][0] = o[q2][p][1] = 0.f;
    const int nt = (ps + VT - 1) / VT;
    for (int it = 0; it < nt; ++it) {
      if (it + 1 < nt) {
        v_tile_load(it + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      bar_consumers();
      const unsigned char* vb = vbuf + (it & 1) * (VT * HD * 2);
      const int t0 = it * VT, tn = min(VT, ps - t0);
#pragma unroll
      for (int q2 = 0; q2 < 2; ++q2) {
        const int i = warp + q2 * NW;
        if (i < hs) {
          const float* s = score + i * tmax + t0;
          int k = 0;
#if QIE_PV_PIPE
          if (tn >= 4) {
            // groups of 4 positions, the next group's probabilities and V words requested before the
            // current group's FFMA chain runs (same ascending order: the chain is the spec, its
            // operands just arrive earlier)
            const unsigned char* vl = vb + 4 * lane;
            float4 pr = *reinterpret_cast<const float4*>(s);
            uint32_t vv[4][NP];
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int p = 0; p < NP; ++p) vv[j][p] = *reinterpret_cast<const uint32_t*>(vl + j * (HD * 2) + 128 * p);
            for (; k + 4 <= tn; k += 4) {
              const bool more = k + 8 <= tn;
              const int kn = more ? k + 4 : k;
              const float4 prn = *reinterpret_cast<const float4*>(s + kn);
              uint32_t vn[4][NP];
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int p = 0; p < NP; ++p) vn[j][p] = *reinterpret_cast<const uint32_t*>(vl + (kn + j) * (HD * 2) + 128 * p);
              const float prj[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                  o[q2][p][0] = __fmaf_rn(prj[j], lo2f(vv[j][p]), o[q2][p][0]);
                  o[q2][p][1] = __fmaf_rn(prj[j], hi2f(vv[j][p]), o[q2][p][1]);
                }
              pr = prn;
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int p = 0; p < NP; ++p) vv[j][p] = vn[j][p];
            }
          }
#else
          for (; k + 4 <= tn; k += 4) {
            const float4 pr = *reinterpret_cast<const float4*>(s + k);
            const float prj[4] = {pr.x, pr.y, pr.z, pr.w};
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int p = 0; p < NP; ++p) {
                const uint32_t vv = *reinterpret_cast<const uint32_t*>(vb + (k + j) * (HD * 2) + (64 * p + 2 * lane) * 2);
                o[q2][p][0] = __fmaf_rn(prj[j], lo2f(vv), o[q2][p][0]);
                o[q2][p][1] = __fmaf_rn(prj[j], hi2f(vv), o[q2][p][1]);
              }
          }
#endif
          for (; k < tn; ++k) {
            const float pk = s[k];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              const uint32_t vv = *reinterpret_cast<const uint32_t*>(vb + k * (HD * 2) + (64 * p + 2 * lane) * 2);
              o[q2][p][0] = __fmaf_rn(pk, lo2f(vv), o[q2][p][0]);
              o[q2][p][1] = __fmaf_rn(pk, hi2f(vv), o[q2][p][1]);
            }
          }
        }
      }
      bar_consumers();  // the buffer is refilled two tiles later
    }
    __syncwarp();
#pragma unroll
    for (int q2 = 0; q2 < 2; ++q2) {
      const int i = warp + q2 * NW;
      if (i < hs) {
        const float pk = score[i * tmax + ps];  // the new position: V from this step's projection
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const uint32_t vv = *reinterpret_cast<const uint32_t*>(vnew + 64 * p + 2 * lane);
          o[q2][p][0] = __fmaf_rn(pk, lo2f(vv), o[q2][p][0]);
          o[q2][p][1] = __fmaf_rn(pk, hi2f(vv), o[q2][p][1]);
        }
        head_store<NP>(o[q2], a.att + (size_t)b * Dq + (size_t)(h0 + i) * HD, lane);
      }
    }
    bar_consumers();  // shared memory is reused by the next task / phase
    lap(13);
  }
}

// ---------------------------------------------------------------- attention, query-group tasks (head_dim 64)
// Many rows (B * n_q > grid): a task is (row, kv head) and the n_q / n_kv query heads of the group share ONE stream
// of the cached K and V rows.  Both travel through shared memory in tiles of AT positions (one per consumer
// thread), copied with coalesced 16-byte cp.async into 128-byte rows whose 16-byte chunk c sits at c ^ (row & 7)
// (a thread reading ITS row with LDS.128 is conflict free), two buffers: K tile j+1 / V tile 0 are in flight while
// tile j is used.  What the reference fixes is kept (self_attension.cu:47-137): products then the 2^k tree per
// score, max, expf, the sequential sum, p = e / sum, the sequential fma chain over positions per output element.
//  * scores: thread = cached position, the query group's heads in turn (q as fp32 in shared memory);
//  * softmax: one warp per head; the sum chain runs in lane 0 with its operands requested 16 positions ahead;
//  * PV: warps 0..3 own heads w, w+4, w+8, w+12 with lane l on output dims 2l, 2l+1: one LDS.32 of a V row feeds
//    the fma chains of all heads of the warp (packed FFMA2: two IEEE fp32 fmas per instruction), the four
//    probabilities of a head arrive with one broadcast LDS.128.  (r01: one warp per head re-read every V row
//    7 times and the phase was bound by shared-memory wavefronts: 27 cycles per position, now ~8.)
constexpr int AT = NTC;            // positions per K/V tile
constexpr int AT_BYTES = AT * 128;  // head_dim 64, bf16
__host__ __device__ inline int attn_score_pitch(int max_kv_len) { return ((max_kv_len + 1 + 31) & ~31) + 4; }  // pitch % 32 == 4

// PV of one V tile (tn cached positions, rows of 128 bytes, lane l = output dims 2l, 2l+1) for the NH heads of a warp:
// o[h] = fma(p[h][k], v[k], o[h]) for k ascending -- the reference's chain (self_attension.cu:112-137), two IEEE fmas
// per packed FFMA2.  Groups of 4 positions in two operand sets; a set is reloaded (for 8 positions ahead) right
// after it has been consumed, so every load has the other set's chains to land.
template <int NH>
__device__ __forceinline__ void pv_tile(const unsigned char* __restrict__ vl, const float* __restrict__ sb, int hstride, int tn,
                                        float2 (&o)[2]) {
  // bf16 pair -> two floats on the ALU pipe only (PRMT / LOP3): the shift form becomes IMAD.U32 every other time, which
  // shares the FMA pipe with the FFMA2 chain (math-pipe throttle was 12 % of this loop's stall samples)
  auto cvt2 = [](uint32_t v) { return make_float2(__uint_as_float(__byte_perm(v, 0u, 0x1044)), __uint_as_float(v & 0xffff0000u)); };
  auto load4 = [&](uint32_t (&vv)[4], float4 (&pr)[NH], int kk) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) vv[jj] = *reinterpret_cast<const uint32_t*>(vl + (kk + jj) * 128);
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) pr[hh] = *reinterpret_cast<const float4*>(sb + hh * hstride + kk);
  };
  auto fma4 = [&](const uint32_t (&vv)[4], const float4 (&pr)[NH]) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float2 vf = cvt2(vv[jj]);
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        const float pj = jj == 0 ? pr[hh].x : (jj == 1 ? pr[hh].y : (jj == 2 ? pr[hh].z : pr[hh].w));
        o[hh] = __ffma2_rn(make_float2(pj, pj), vf, o[hh]);
      }
    }
  };
  int k = 0;
  // 16 positions per iteration in four operand sets of 4; a set is reloaded (16 positions ahead) right behind its use,
  // so every load has the other three sets' chains (~50 cycles) to land and the loop overhead is paid once per 16
  const int n16 = tn & ~15;
  if (n16) {
    uint32_t va[4], vb[4], vc[4], vd[4];
    float4 pa[NH], pb[NH], pc[NH], pd[NH];
    load4(va, pa, 0);
    load4(vb, pb, 4);
    load4(vc, pc, 8);
    load4(vd, pd, 12);
#pragma unroll 1
    for (; k + 16 < n16; k += 16) {
      fma4(va, pa);
      load4(va, pa, k + 16);
      fma4(vb, pb);
      load4(vb, pb, k + 20);
      fma4(vc, pc);
      load4(vc, pc, k + 24);
      fma4(vd, pd);
      load4(vd, pd, k + 28);
    }
    fma4(va, pa);
    fma4(vb, pb);
    fma4(vc, pc);
    fma4(vd, pd);
    k += 16;
  }
  for (; k < tn; ++k) {
    const uint32_t v1 = *reinterpret_cast<const uint32_t*>(vl + k * 128);
    const float2 vf = cvt2(v1);
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      const float pj = sb[hh * hstride + k];
      o[hh] = __ffma2_rn(make_float2(pj, pj), vf, o[hh]);
    }
  }
}

__device__ __noinline__ void attention_group_phase(const MegaArgs& a, int layer, unsigned char* smem) {
  constexpr int HD = 64, NP = 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a task is (row, kv head, part): the query group of a kv head (hsk heads) is cut into parts of hp heads while all
  // tasks still fit one wave of CTAs (batch 64: one part of 7 heads, 128 tasks; batch 32: 4 + 3 heads, 128 tasks; batch
  // 16: 2 + 2 + 2 + 1, 128 tasks) -- r02 first version: one task per (row, kv head) left most SMs idle below 64 rows
  const int hsk = a.n_q / a.n_kv, hp = a.attn_hp, nparts = (hsk + hp - 1) / hp;
  const int Dq = a.n_q * HD, Dkv = a.n_kv * HD, QKV = Dq + 2 * Dkv;
  const int ntask = a.B * a.n_kv * nparts;
  const int SP = attn_score_pitch(a.max_kv_len);
  const int psz = a.kv.page_size;             // a power of two on this path (mega_geometry)
  const int psz_shift = __ffs(psz) - 1;
  const uint32_t psz_mask = (uint32_t)psz - 1u;
  const uint32_t pstride_b = (uint32_t)(a.kv.page_stride() * sizeof(bf16));
  const MegaLayer& w = reinterpret_cast<const MegaLayer*>(smem + OFF_LAYERS)[layer];

  float* q_s = reinterpret_cast<float*>(smem + a.off_act);   // [hp][HD] fp32
  bf16* knew = reinterpret_cast<bf16*>(q_s + hp * HD);        // [HD]
  bf16* vnew = knew + HD;
  float* score = reinterpret_cast<float*>(vnew + HD);         // [hs][SP]
  unsigned char* kvb = reinterpret_cast<unsigned char*>(score + hp * SP);  // [2][AT][128]
  const uint32_t kvb_u32 = smem_u32(kvb);
  int* pages = reinterpret_cast<int*>(kvb + 2 * AT_BYTES);    // [max_kv_len / page_size + 1]

  unsigned long long* dbg = reinterpret_cast<unsigned long long*>(smem + OFF_DBG);
  const bool timed = a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
    long long tq = timed ? clock64() : 0;
    auto lap = [&](int slot) {
      if (timed) {
        const long long t = clock64();
        dbg[slot] += (unsigned long long)(t - tq);
        tq = t;
      }
    };
    const int bk = task / nparts, part = task - bk * nparts;
    const int b = bk / a.n_kv, kvh = bk - b * a.n_kv, h0 = kvh * hsk + part * hp;
    const int hs = min(hp, hsk - part * hp);  // query heads of this task
    const bool writer = part == 0;           // one task per (row, kv head) stores the new K / V row
    const int ps = a.pos[b];                      // cached positions 0..ps-1, the new one is ps
    const int* bt = a.block_table + (size_t)a.slot[b] * a.max_pages;
    const int ns = ps / AT + 1;                   // score tiles (the last one holds position ps)
    const int nt = (ps + AT - 1) / AT;            // tiles with cached rows
    // job j < ns: K tile j; job ns + i: V tile i.  Buffer j & 1.  K rows are stored with the chunk swizzle (a thread
    // reads ITS row), V rows linearly (a warp reads ONE row, lane l its word l).  The 8 copies of a thread are
    // independent (page ids come from shared memory), so they are issued back to back.
    auto tile_load = [&](int j) {
      const int kv = j >= ns ? 1 : 0, it = kv ? j - ns : j;
      const int t0 = it * AT, tn = max(0, min(AT, ps - t0));
      const uint32_t dst = kvb_u32 + (j & 1) * AT_BYTES + (threadIdx.x >> 3) * 128;
      const uint32_t cc = threadIdx.x & 7;
      // byte address of (page 0, this layer, K|V, this kv head, slot 0, chunk cc); a page id adds pstride_b, a slot 128
      const unsigned char* base = reinterpret_cast<const unsigned char*>(a.kv.chunk(0, layer, kv, kvh)) + cc * 16;
#pragma unroll
      for (int i = 0; i < AT * 8 / NTC; ++i) {
        const int r = (int)(threadIdx.x >> 3) + i * (NTC / 8);
        if (r < tn) {
          const uint32_t k = (uint32_t)(t0 + r);
          const uint32_t col = kv ? cc : (cc ^ (uint32_t)(r & 7));
          const size_t off = (size_t)(uint32_t)pages[k >> psz_shift] * pstride_b + (k & psz_mask) * 128u;  // one IMAD.WIDE
          cp_async16(dst + i * (NTC / 8) * 128 + (col << 4), base + off);
        }
      }
      cp_async_commit();
    };
    const int njobs = ns + nt;
    for (int i = threadIdx.x; i <= ps / psz; i += NTC) pages[i] = bt[i];
    bar_consumers();
    tile_load(0);  // cached rows do not depend on this step's q/k/v: in flight during the set-up
    // q/k-norm + RoPE of the group's query heads and of the new K row; KV store of the new position
    {
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
          *reinterpret_cast<float2*>(q_s + r * HD + 2 * lane) = make_float2(v[0][0], v[0][1]);
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
    }
    lap(10);

    // ---- scores (self_attension.cu:47-74)
    for (int j = 0; j < ns; ++j) {
      const long long tw0 = timed ? clock64() : 0;
      if (j + 1 < njobs) {
        tile_load(j + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      bar_consumers();  // tile j (and, at j = 0, q_s / knew) visible
      if (timed) dbg[14] += (unsigned long long)(clock64() - tw0);
      const int k = j * AT + threadIdx.x;
      if (k <= ps) {
        const unsigned char* base = k == ps ? reinterpret_cast<const unsigned char*>(knew) : kvb + (j & 1) * AT_BYTES + threadIdx.x * 128;
        const int swz = k == ps ? 0 : (threadIdx.x & 7);
        float2 kf[HD / 2];
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
          const uint4 t = *reinterpret_cast<const uint4*>(base + ((i ^ swz) << 4));
          kf[4 * i] = make_float2(lo2f(t.x), hi2f(t.x));
          kf[4 * i + 1] = make_float2(lo2f(t.y), hi2f(t.y));
          kf[4 * i + 2] = make_float2(lo2f(t.z), hi2f(t.z));
          kf[4 * i + 3] = make_float2(lo2f(t.w), hi2f(t.w));
        }
        // dot / sqrtf(64): the divisor is exactly 8, and x / 8 == x * 0.125 bit for bit
        float4 wa[4], wb[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          wa[u] = *reinterpret_cast<const float4*>(q_s + 4 * u);
          wb[u] = *reinterpret_cast<const float4*>(q_s + HD / 2 + 4 * u);
        }
        for (int i = 0; i < hs; ++i)
          score[i * SP + k] = __fmul_rn(dot_tree_w<HD>(wa, wb, kf, q_s + i * HD, q_s + (i + 1 < hs ? i + 1 : i) * HD), 0.125f);
      }
      bar_consumers();  // buffer j & 1 is refilled by job j + 2
    }
    lap(11);

    // ---- softmax (self_attension.cu:94-107): one warp per head
    for (int i = warp; i < hs; i += NW) {
      float* s = score + i * SP;
      const int n = ps + 1;
      float m = -1e9f;
      for (int k = lane; k < n; k += 32) m = fmaxf(m, s[k]);
      m = warp_max(m);
      for (int k = lane; k < n; k += 32) s[k] = expf(__fsub_rn(s[k], m));
      __syncwarp();
      float sum = 0.f;
      if (lane == 0) {
        // the reference's sequential sum: blocks of 16 values, the next block requested before the current FADD chain
        const int n16 = n & ~15;
        float4 c[4], nx[4];
        if (n16) {
#pragma unroll
          for (int u = 0; u < 4; ++u) c[u] = *reinterpret_cast<const float4*>(s + 4 * u);
        }
        for (int k = 0; k < n16; k += 16) {
          const int kn = k + 16 < n16 ? k + 16 : k;
#pragma unroll
          for (int u = 0; u < 4; ++u) nx[u] = *reinterpret_cast<const float4*>(s + kn + 4 * u);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            sum = __fadd_rn(sum, c[u].x);
            sum = __fadd_rn(sum, c[u].y);
            sum = __fadd_rn(sum, c[u].z);
            sum = __fadd_rn(sum, c[u].w);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) c[u] = nx[u];
        }
        for (int k = n16; k < n; ++k) sum = __fadd_rn(sum, s[k]);
      }
      sum = __shfl_sync(0xffffffffu, sum, 0);
      for (int k = lane; k < n; k += 32) s[k] = __fdiv_rn(s[k], sum);
    }
    lap(12);

    // ---- PV (self_attension.cu:112-137)
    float2 o[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    const int nh = warp < hs ? (warp + NW < hs ? 2 : 1) : 0;  // heads warp and warp + NW
    for (int it = 0; it < nt; ++it) {
      const int j = ns + it;
      const long long tw0 = timed ? clock64() : 0;
      if (j + 1 < njobs) {
        tile_load(j + 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      bar_consumers();  // V tile `it` visible (and, at it = 0, the probabilities)
      if (timed) dbg[15] += (unsigned long long)(clock64() - tw0);
      const unsigned char* vl = kvb + (j & 1) * AT_BYTES + 4 * lane;  // V rows are stored linearly: lane l reads word l of a row
      const int t0 = it * AT, tn = min(AT, ps - t0);
      if (nh == 1)
        pv_tile<1>(vl, score + warp * SP + t0, NW * SP, tn, o);
      else if (nh == 2)
        pv_tile<2>(vl, score + warp * SP + t0, NW * SP, tn, o);
      bar_consumers();  // the buffer is refilled two tiles later
    }
    if (nt == 0) bar_consumers();  // the probabilities of other warps are read below
    {
      const uint32_t v1 = *reinterpret_cast<const uint32_t*>(vnew + 2 * lane);  // the new position: V of this step's projection
      const float2 vf = make_float2(lo2f(v1), hi2f(v1));
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
        if (hh < nh) {
          const float pj = score[(warp + NW * hh) * SP + ps];
          o[hh] = __ffma2_rn(make_float2(pj, pj), vf, o[hh]);
          *reinterpret_cast<uint32_t*>(a.att + (size_t)b * Dq + (size_t)(h0 + warp + NW * hh) * HD + 2 * lane) = pack2(f2bf(o[hh].x), f2bf(o[hh].y));
        }
    }
    bar_consumers();  // shared memory is reused by the next task / phase
    lap(13);
  }
}

// ---------------------------------------------------------------- attention: query-group tasks, K/V streamed by TMA
// Same tasks and the same arithmetic as attention_group_phase (the reference's order, self_attension.cu:47-137), but
// the cached rows arrive through the KV tensor map: one cp.async.bulk.tensor request per page chunk (16 positions =
// 2 KB) instead of 128 cp.async instructions, written with the 128-byte swizzle that both readers want.
//  * scores: groups of 64 positions; warp w owns ring slot w and the groups w, w + 7, ...: it waits for its slot,
//    pulls TWO rows per lane (positions l and l + 32 of the group) into fp32 registers, re-arms the slot with its next
//    group and evaluates the product tree of both rows for every head of the task.  One broadcast LDS.128 of q now
//    feeds two positions (the loop was bound by those loads: 16 per head and position), the two trees give the
//    scheduler independent work, and no CTA-wide barrier separates the groups.  q is stored in the order the tree
//    consumes it: quad u = {q[2m], q[2m+1], q[2m+32], q[2m+33]}, m = bitrev4(u), so that partial sums meet as early as
//    possible (depth-first, 4 live partial sums per row instead of 16).
//    products of two bf16 values are exact in fp32, so  p[i] + p[i+32]  ==  fma(q[i+32], k[i+32], p[i])  bit for bit
//    (one rounding of the same exact sum) unless a product underflows below 2^-126 -- |q k| < 1e-38 does not occur
//    for normalised heads.
//  * softmax: one warp per head; exponentials of block j + 2 are computed by all lanes while every lane walks the
//    reference's sequential sum over block j (broadcast LDS.128, no divergence), so the FADD chain (4 cycles per
//    position) is the only cost.
//  * PV: 4 warps own heads w and w + 4 (two FFMA2 chains per lane, one LDS.32 + one widening per V row for both); a
//    fifth warp keeps the V ring (7 slots of 64 positions) filled; mbarriers per slot, no CTA-wide barrier per tile.
constexpr int AG = 64;              // cached positions per K slot (one slot per warp)
constexpr int AG_BYTES = AG * 128;  // head_dim 64, bf16
constexpr int AVG = 128;            // cached positions per V slot: the PV warps pay the slot hand-over half as often
constexpr int AVG_BYTES = AVG * 128;
constexpr int NVS = NW * AG / AVG;  // V slots in the same ring memory (3)
constexpr int OFF_AKF = 656;        // NW mbarriers: K slot of warp w filled      (behind OFF_DBG)
constexpr int OFF_AVF = 896;        // 3 mbarriers: V slot s filled               (behind OFF_REL)
constexpr int OFF_AVE = 920;        // 3 mbarriers: V slot s read by all PV warps
constexpr int PVW = 4;              // PV warps; warp PVW feeds the V ring
constexpr int AQ = 20;              // q quads per head in shared memory: 8 (lanes with even pairs) + 1 pad + 8 (odd pairs) + 3 pad --
                                    // the two quads a lane pair reads together are 144 bytes apart: different banks
static_assert(OFF_DBG + 16 * 8 <= OFF_AKF && OFF_AKF + NW * 8 <= OFF_REL && OFF_REL + MAX_SLOTS * 4 <= OFF_AVF &&
                  OFF_AVF + NVS * 8 <= OFF_AVE && OFF_AVE + NVS * 8 <= OFF_TSF && OFF_TSF + 4 * 8 <= OFF_TSE && OFF_TSE + 4 * 8 <= OFF_RMS,
              "shared-memory header layout");

// ring state kept across tasks and layers, in the shared-memory header (registers of the caller stay untouched):
// u32 kuse[NW] fills of warp w's K slot so far | u32 vseq: V groups streamed so far (slot = vseq % NVS, fill number =
// vseq / NVS) | u32 pre: the K groups of the CTA's first task were requested in front of the grid barrier
constexpr int OFF_APIPE = OFF_AKF + NW * 8;
static_assert(OFF_APIPE + (NW + 2) * 4 <= OFF_TSSEQ && OFF_TSSEQ + 4 <= OFF_REL, "shared-memory header layout");

struct AttnTmaLayout {
  int q, knew, sums, score, ring, total;  // byte offsets from the activation area
};
__host__ __device__ inline AttnTmaLayout attn_tma_layout(int off_act, int hp, int max_kv_len) {
  AttnTmaLayout l;
  l.q = 0;
  l.knew = hp * AQ * 16;
  l.sums = l.knew + 2 * 64 * 2;
  l.score = l.sums + 64;
  l.ring = ((off_act + l.score + hp * attn_score_pitch(max_kv_len) * 4 + 1023) & ~1023) - off_act;
  l.total = l.ring + NW * AG_BYTES;
  return l;
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* map, uint32_t mbar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(map), "r"(mbar), "r"(c0), "r"(c1)
               : "memory");
}

// group g (cached positions GS g .. min(GS g + GS, ps) - 1) of (layer, K|V, kv head) -> ring slot at dst: lane i requests
// box i (min(page, 64) slots of one page chunk); whole boxes are copied, rows behind ps are never read
template <int GS>
__device__ __forceinline__ void attn_issue_group(const MegaArgs& a, uint32_t dst, uint32_t mbar, const int* __restrict__ bt, int layer,
                                                 int kv, int kvh, int g, int ps, int lane) {
  const int psz = a.kv.page_size, sh = __ffs(psz) - 1;
  const int br = min(psz, AG), bsh = min(sh, 6);  // rows per box (the tensor map's box: min(page, 64) slots)
  const int nbox = (min(GS, ps - GS * g) + br - 1) >> bsh;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slot was read / written through the generic proxy
  if (lane == 0) mbar_expect_tx(mbar, (uint32_t)(nbox * br * 128));
  if (lane < nbox) {
    const int p0 = GS * g + lane * br;
    const int page = __ldg(bt + (p0 >> sh));
    const int row = (((page * a.kv.n_layers + layer) * 2 + kv) * a.kv.n_kv + kvh) * psz + (p0 & (psz - 1));
    tma_load_2d(dst + (uint32_t)(lane * br * 128), a.kvmap, mbar, 0, row);
  }
}

__host__ __device__ constexpr int trailing_ones(int u) { return (u & 1) ? ((u & 2) ? ((u & 4) ? ((u & 8) ? 4 : 3) : 2) : 1) : 0; }

// Half of the product tree for TWO cached rows against one query head.  The last combine of the reference's tree
// but one (stride 2 elements) adds the sums over the even and the odd element pairs, so a lane pair splits a row by
// pair parity: lane `half` evaluates strides 32 ... 4 over its 16 pairs (8 q quads, tree order) for both rows and the
// two lanes swap one partial sum each.  kw[c][0|1]: pairs 4c + half and 4c + 2 + half of the row, as fp32.
// w: rolling window of four q quads, refilled four steps ahead (from the next head at the end).
__device__ __forceinline__ void dot_half2(float4 (&w)[4], const float4* __restrict__ qc, const float4* __restrict__ qn,
                                          const float2 (&ka)[8][2], const float2 (&kb)[8][2], float2& ya, float2& yb) {
  float2 st0[3], st1[3], x0, x1;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int c = ((u & 1) << 1) | ((u & 2) >> 1), hl = u >> 2;  // pair bitrev4(u) + half and its partner 16 pairs on
    const float4 q = w[u & 3];
    w[u & 3] = u + 4 < 8 ? qc[u + 4] : qn[u + 4 - 8];
    const float2 qa = make_float2(q.x, q.y), qb = make_float2(q.z, q.w);
    x0 = __ffma2_rn(qb, ka[c + 4][hl], __fmul2_rn(qa, ka[c][hl]));  // stride 32 (exact products)
    x1 = __ffma2_rn(qb, kb[c + 4][hl], __fmul2_rn(qa, kb[c][hl]));
    const int t = trailing_ones(u);
#pragma unroll
    for (int b = 0; b < 3; ++b)
      if (b < t) {  // strides 16, 8, 4 as soon as both operands exist
        x0 = __fadd2_rn(st0[b], x0);
        x1 = __fadd2_rn(st1[b], x1);
      }
    if (t < 3) {
      st0[t] = x0;
      st1[t] = x1;
    }
  }
  ya = x0;
  yb = x1;
}

__device__ __forceinline__ float2 bf2x_to_f2(uint32_t v) {  // ALU pipe only (PRMT / LOP3), see pv_tile
  return make_float2(__uint_as_float(__byte_perm(v, 0u, 0x1044)), __uint_as_float(v & 0xffff0000u));
}

// PV over one whole ring slot (AVG = 128 cached rows of V, swizzled: word `lane` of row r sits at (lane << 2) ^ ((r & 7) << 4)) for
// the two heads of a warp: o[h] = fma(p[h][k], v[k], o[h]) for k ascending -- the reference's chain
// (self_attension.cu:112-137), two IEEE fmas per packed FFMA2, lane l = output dims 2l, 2l + 1.  The body is PVU rows
// unrolled (every address is one of 8 row registers + an immediate: a position costs one LDS.32, the widening and two
// FFMA2, + a quarter of a broadcast LDS.128 per head); chunk c + 1 is loaded before the chains of chunk c.
// ONE copy of the body serves every PV warp: a warp with a single head runs the second chain on a dummy probability
// row (its scheduler has the slots to spare) -- the step executes ~120 KB of code per layer, right at the size of the
// instruction cache, and every KB less shows up in ALL phases (cold code costs ~10 cycles per instruction).
constexpr int PVU = 64;
__device__ __forceinline__ void pv_group(const unsigned char* __restrict__ slot, const uint32_t (&xo)[8], const float* __restrict__ p0,
                                         const float* __restrict__ p1, float2 (&o)[2]) {
  const unsigned char* rowp[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) rowp[j] = slot + xo[j];
  uint32_t v[2][8];
  float4 pr[2][2][2];
  auto load = [&](int c) {  // c: chunk of 8 rows within the current block of PVU (c = PVU / 8: first chunk of the next block)
#pragma unroll
    for (int j = 0; j < 8; ++j) v[c & 1][j] = *reinterpret_cast<const uint32_t*>(rowp[j] + c * 1024 + j * 128);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      pr[c & 1][0][u] = *reinterpret_cast<const float4*>(p0 + 8 * c + 4 * u);
      pr[c & 1][1][u] = *reinterpret_cast<const float4*>(p1 + 8 * c + 4 * u);
    }
  };
  load(0);
#pragma unroll 1
  for (int blk = 0; blk < AVG / PVU; ++blk) {
#pragma unroll
    for (int c = 0; c < PVU / 8; ++c) {
      if (c + 1 < PVU / 8 || blk + 1 < AVG / PVU) load(c + 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // low half through the FMA pipe (IMAD.SHL), high half through the ALU pipe (LOP3)
        const float2 vf = make_float2(__uint_as_float(v[c & 1][j] * 65536u), __uint_as_float(v[c & 1][j] & 0xffff0000u));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          const float4 p4 = pr[c & 1][hh][j >> 2];
          const float pj = (j & 3) == 0 ? p4.x : ((j & 3) == 1 ? p4.y : ((j & 3) == 2 ? p4.z : p4.w));
          o[hh] = __ffma2_rn(make_float2(pj, pj), vf, o[hh]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) rowp[j] += PVU * 128;
    p0 += PVU;
    p1 += PVU;
  }
}
// the rows of a partial last group, one at a time
template <int NH>
__device__ __forceinline__ void pv_tail(const unsigned char* __restrict__ slot, int lane, const float* __restrict__ p0,
                                        const float* __restrict__ p1, int nrows, float2 (&o)[2]) {
  for (int r = 0; r < nrows; ++r) {
    const float2 vf = bf2x_to_f2(*reinterpret_cast<const uint32_t*>(slot + r * 128 + (uint32_t)((lane << 2) ^ ((r & 7) << 4))));
    o[0] = __ffma2_rn(make_float2(p0[r], p0[r]), vf, o[0]);
    if (NH == 2) o[1] = __ffma2_rn(make_float2(p1[r], p1[r]), vf, o[1]);
  }
}

// requests for the K groups of the CTA's first attention task, issued behind the QKV GEMM and in front of the grid
// barrier that publishes q/k/v (cached rows do not depend on this step)
__device__ __noinline__ void attention_tma_prefetch(const MegaArgs& a, int layer, unsigned char* smem) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hsk = a.n_q / a.n_kv, hp = a.attn_hp, nparts = (hsk + hp - 1) / hp;
  const int ntask = a.B * a.n_kv * nparts;
  const int task = blockIdx.x;
  if (task >= ntask) return;  // (pre stays 0: no task, nobody reads it)
  const int bk = task / nparts, b = bk / a.n_kv, kvh = bk - b * a.n_kv;
  const int ps = a.pos[b];
  const AttnTmaLayout lay = attn_tma_layout(a.off_act, hp, a.max_kv_len);
  const uint32_t ring = smem_u32(smem + a.off_act + lay.ring);
  if (AG * warp < ps)
    attn_issue_group<AG>(a, ring + warp * AG_BYTES, smem_u32(smem + OFF_AKF + warp * 8), a.block_table + (size_t)a.slot[b] * a.max_pages, layer,
                     0, kvh, warp, ps, lane);
  if (threadIdx.x == 0) reinterpret_cast<uint32_t*>(smem + OFF_APIPE)[NW + 1] = 1u;  // read behind the grid barrier's bar.sync
}

__device__ __noinline__ void attention_tma_phase(const MegaArgs& a, int layer, unsigned char* smem) {
  constexpr int HD = 64, NP = 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hsk = a.n_q / a.n_kv, hp = a.attn_hp, nparts = (hsk + hp - 1) / hp;
  const int Dq = a.n_q * HD, Dkv = a.n_kv * HD, QKV = Dq + 2 * Dkv;
  const int ntask = a.B * a.n_kv * nparts;
  const int SP = attn_score_pitch(a.max_kv_len);
  const MegaLayer& w = reinterpret_cast<const MegaLayer*>(smem + OFF_LAYERS)[layer];
  const AttnTmaLayout lay = attn_tma_layout(a.off_act, hp, a.max_kv_len);
  unsigned char* area = smem + a.off_act;
  float* q_s = reinterpret_cast<float*>(area + lay.q);  // [hp][AQ quads], tree order
  bf16* knew = reinterpret_cast<bf16*>(area + lay.knew);
  bf16* vnew = knew + HD;
  float* score = reinterpret_cast<float*>(area + lay.score);  // [hp][SP]
  unsigned char* ringp = area + lay.ring;
  const uint32_t ring = smem_u32(ringp);
  const uint32_t kfull = smem_u32(smem + OFF_AKF + warp * 8);
  const uint32_t vfull = smem_u32(smem + OFF_AVF), vempty = smem_u32(smem + OFF_AVE);

  volatile uint32_t* pipe = reinterpret_cast<volatile uint32_t*>(smem + OFF_APIPE);
  uint32_t kuse = pipe[warp], vseq = pipe[NW];
  bool pre = pipe[NW + 1] != 0u;

  unsigned long long* dbg = reinterpret_cast<unsigned long long*>(smem + OFF_DBG);
  const bool timed = a.prof != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
    long long tq = timed ? clock64() : 0;
    auto lap = [&](int slot) {
      if (timed) {
        const long long t = clock64();
        dbg[slot] += (unsigned long long)(t - tq);
        tq = t;
      }
    };
    const int bk = task / nparts, part = task - bk * nparts;
    const int b = bk / a.n_kv, kvh = bk - b * a.n_kv, h0 = kvh * hsk + part * hp;
    const int hs = min(hp, hsk - part * hp);  // query heads of this task
    const bool writer = part == 0;           // one task per (row, kv head) stores the new K / V row
    const int ps = a.pos[b];                 // cached positions 0..ps-1, the new one is ps
    const int* bt = a.block_table + (size_t)a.slot[b] * a.max_pages;
    const int ng = (ps + AG - 1) / AG;       // groups with cached rows
    const int nsg = ps / AG + 1;             // score groups (the last one holds position ps)
    if (!pre && AG * warp < ps) attn_issue_group<AG>(a, ring + warp * AG_BYTES, kfull, bt, layer, 0, kvh, warp, ps, lane);
    pre = false;
    // q/k-norm + RoPE of the group's query heads and of the new K row; KV store of the new position.  Rows
    // warp and warp + NW: both loads are in flight before the first norm
    {
      const bf16* row = a.qkv + (size_t)b * QKV;
      const float* cos_row = a.cos_t + (size_t)ps * 32 * NP;
      const float* sin_row = a.sin_t + (size_t)ps * 32 * NP;
      const int psz = a.kv.page_size;
      const int new_page = bt[ps / psz], new_off = ps % psz;
      float v[2][NP][2];
      auto src_of = [&](int r) { return r < hs ? row + (size_t)(h0 + r) * HD : (r == hs ? row + Dq + (size_t)kvh * HD : row + Dq + Dkv + (size_t)kvh * HD); };
#pragma unroll
      for (int t = 0; t < 2; ++t)
        if (warp + NW * t < hs + 2) head_load_cg<NP>(v[t], src_of(warp + NW * t), lane);
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int r = warp + NW * t;
        if (r < hs) {
          if (w.q_norm) head_norm<NP>(v[t], w.q_norm, lane);
          head_rope<NP>(v[t], cos_row, sin_row, lane);
          // pair `lane` = elements 2 lane, 2 lane + 1: first (lane < 16) or second half of quad bitrev4(lane & 15)
          const int u = (int)(__brev((unsigned)(lane & 15)) >> 28);
          *reinterpret_cast<float2*>(q_s + r * (4 * AQ) + 4 * (u + (u >> 3)) + 2 * (lane >> 4)) = make_float2(v[t][0][0], v[t][0][1]);
        } else if (r == hs) {
          if (w.k_norm) head_norm<NP>(v[t], w.k_norm, lane);
          head_rope<NP>(v[t], cos_row, sin_row, lane);
          head_store<NP>(v[t], knew, lane);
          if (writer) head_store<NP>(v[t], a.kv.chunk(new_page, layer, 0, kvh) + (size_t)new_off * HD, lane);
        } else if (r == hs + 1) {
          head_store<NP>(v[t], vnew, lane);
          if (writer) head_store<NP>(v[t], a.kv.chunk(new_page, layer, 1, kvh) + (size_t)new_off * HD, lane);
        }
      }
    }
    bar_consumers();  // q_s / knew / vnew visible
    lap(10);

    // ---- scores (self_attension.cu:47-74)
    {
      const int half = lane & 1, tp = lane >> 1;
      const float4* q4 = reinterpret_cast<const float4*>(q_s) + 9 * half;  // this lane's 8 quads of every head
      // (measured dead end: dealing the groups 8 ways with a double share for warp 3, which has its scheduler almost
      // to itself -- 33 k -> 38 k cycles: a single warp is bound by the latency of its own trees, not by the FP pipe)
      constexpr int gstep = NW;
      for (int g = warp; g < nsg; g += gstep) {
        if (g < ng) {
          const long long tw0 = timed ? clock64() : 0;
          mbar_wait(kfull, kuse & 1u);
          ++kuse;
          if (timed) dbg[14] += (unsigned long long)(clock64() - tw0);
        }
#pragma unroll 1
        for (int sp = 0; sp < 2; ++sp) {  // 32 positions per pass: lane pair tp owns rows 32 sp + tp and 32 sp + 16 + tp
          float2 kw[2][8][2];
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const int r = 32 * sp + 16 * p + tp, k = AG * g + r;
            const bool cached = k < ps;  // k == ps: the new row; k > ps: nothing (any readable address)
            const unsigned char* rowp = cached ? ringp + warp * AG_BYTES + r * 128 : reinterpret_cast<const unsigned char*>(knew);
            const int swz = cached ? (r & 7) : 0;
#pragma unroll
            for (int c = 0; c < HD / 8; ++c) {
              const uint4 t = *reinterpret_cast<const uint4*>(rowp + ((c ^ swz) << 4));
              kw[p][c][0] = bf2x_to_f2(half ? t.y : t.x);
              kw[p][c][1] = bf2x_to_f2(half ? t.w : t.z);
            }
          }
          if (sp == 1) {  // the slot has been read: re-arm it with this warp's next group
            __syncwarp();
            if (g + gstep < ng) attn_issue_group<AG>(a, ring + warp * AG_BYTES, kfull, bt, layer, 0, kvh, g + gstep, ps, lane);
          }
          float4 wq[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) wq[u] = q4[u];
          const int kmine = AG * g + 32 * sp + 16 * half + tp;  // the row this lane finishes
          for (int i = 0; i < hs; ++i) {
            float2 ya, yb;
            dot_half2(wq, q4 + AQ * i, q4 + AQ * (i + 1 < hs ? i + 1 : i), kw[0], kw[1], ya, yb);
            // stride 2: sum over the even pairs + sum over the odd pairs (the other lane's half), then stride 1
            const float2 give = half ? ya : yb, keep = half ? yb : ya;
            const float2 got = make_float2(__shfl_xor_sync(0xffffffffu, give.x, 1), __shfl_xor_sync(0xffffffffu, give.y, 1));
            const float2 d = __fadd2_rn(keep, got);
            // dot / sqrtf(64): the divisor is exactly 8, and x / 8 == x * 0.125 bit for bit
            if (kmine <= ps) score[i * SP + kmine] = __fmul_rn(__fadd_rn(d.x, d.y), 0.125f);
          }
        }
      }
    }
    bar_consumers();  // all scores visible, every K slot is free
    lap(11);

    // ---- V ring: the first NVS groups (the PV warps free a slot per group from then on)
    const int nvg = (ps + AVG - 1) / AVG;  // V groups
    const uint32_t vs0 = vseq % NVS, vpar0 = (vseq / NVS) & 1u;
    if (warp == PVW) {
      uint32_t s = vs0;
      for (int j = 0; j < min(nvg, NVS); ++j) {
        attn_issue_group<AVG>(a, ring + s * AVG_BYTES, vfull + s * 8, bt, layer, 1, kvh, j, ps, lane);
        if (++s == NVS) s = 0;
      }
    }

    // ---- softmax (self_attension.cu:94-107): one warp per head
    for (int i = warp; i < hs; i += NW) {
      float* s = score + i * SP;
      const int n = ps + 1;
      float m = -1e9f;
      {
        int k = 4 * lane;
        for (; k + 3 < n; k += 128) {
          const float4 t = *reinterpret_cast<const float4*>(s + k);
          m = fmaxf(fmaxf(m, fmaxf(t.x, t.y)), fmaxf(t.z, t.w));
        }
        for (; k < n; ++k) m = fmaxf(m, s[k]);  // (at most one lane has a partial quad)
      }
      m = warp_max(m);
      // exponentials two blocks of 32 ahead of the sequential sum; every lane walks the same sum
      const int nfull = n >> 5;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int k = 32 * j + lane;
        if (k < n) s[k] = expf(__fsub_rn(s[k], m));
      }
      __syncwarp();
      float sum = 0.f;
      float4 cur[4], nxt[4];
      auto chain16 = [&](const float4 (&v)[4]) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          sum = __fadd_rn(sum, v[u].x);
          sum = __fadd_rn(sum, v[u].y);
          sum = __fadd_rn(sum, v[u].z);
          sum = __fadd_rn(sum, v[u].w);
        }
      };
#pragma unroll
      for (int u = 0; u < 4; ++u) cur[u] = *reinterpret_cast<const float4*>(s + 4 * u);
      for (int j = 0; j < nfull; ++j) {
#pragma unroll
        for (int u = 0; u < 4; ++u) nxt[u] = *reinterpret_cast<const float4*>(s + 32 * j + 16 + 4 * u);
        // (no branch around the exponential: its dependent chain is scheduled between the FADDs of the sum)
        const int kn = 32 * (j + 2) + lane;
        float en = expf(__fsub_rn(s[min(kn, n - 1)], m));
        asm volatile("" : "+f"(en));  // computed here, for every lane (the compiler would sink it into the conditional store)
        chain16(cur);
#pragma unroll
        for (int u = 0; u < 4; ++u) cur[u] = *reinterpret_cast<const float4*>(s + 32 * (j + 1) + 4 * u);  // (pitch covers one block behind n)
        chain16(nxt);
        if (kn < n) s[kn] = en;
        __syncwarp();
      }
      for (int k = 32 * nfull; k < n; ++k) sum = __fadd_rn(sum, s[k]);
      // (measured dead end: leaving these divisions to the two warps that idle during PV, group by group behind the V
      // ring's "filled" barriers -- softmax 22 k -> 17 k cycles, PV 33 k -> 47 k: the PV warps' schedulers are saturated,
      // whatever else runs on them lengthens the chains)
      for (int k = lane; k < n; k += 32) s[k] = __fdiv_rn(s[k], sum);
    }
    bar_consumers();  // probabilities visible
    lap(12);

    // ---- PV (self_attension.cu:112-137)
    if (warp < PVW) {
      const int nh = warp < hs ? (warp + PVW < hs ? 2 : 1) : 0;  // heads warp and warp + PVW
      float2 o[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      uint32_t ls = vs0, lpar = vpar0;  // slot / fill parity of the current group
      uint32_t xo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) xo[j] = (uint32_t)((lane << 2) ^ (j << 4));  // word `lane` of a swizzled row r: r & 7 == j
      const float* sb = score + warp * SP;
      const int hstride = PVW * SP;
      // every PV warp walks all V groups (also a warp without a head: the "drained" barrier counts PVW arrivals)
      const float* p0 = sb;
      const float* p1 = nh == 2 ? sb + hstride : sb;  // one head: the second chain repeats the first (never stored)
      for (int j = 0; j < nvg; ++j) {
        {
          const long long tw0 = timed ? clock64() : 0;
          mbar_wait(vfull + ls * 8, lpar);
          if (timed) dbg[15] += (unsigned long long)(clock64() - tw0);
        }
        const unsigned char* slot = ringp + ls * AVG_BYTES;
        const int nrows = min(AVG, ps - AVG * j);
        if (nrows == AVG) {
          if (nh >= 1) pv_group(slot, xo, p0, p1, o);
        } else {
          if (nh == 2)
            pv_tail<2>(slot, lane, p0, p1, nrows, o);
          else if (nh == 1)
            pv_tail<1>(slot, lane, p0, p1, nrows, o);
        }
        p0 += AVG;
        p1 += AVG;
        __syncwarp();
        if (lane == 0) mbar_arrive(vempty + ls * 8);
        if (++ls == NVS) {
          ls = 0;
          lpar ^= 1u;
        }
      }
      {
        const uint32_t v1 = *reinterpret_cast<const uint32_t*>(vnew + 2 * lane);  // the new position: V of this step's projection
        const float2 vf = make_float2(lo2f(v1), hi2f(v1));
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
          if (hh < nh) {
            const float pj = sb[hh * hstride + ps];
            o[hh] = __ffma2_rn(make_float2(pj, pj), vf, o[hh]);
            *reinterpret_cast<uint32_t*>(a.att + (size_t)b * Dq + (size_t)(h0 + warp + PVW * hh) * HD + 2 * lane) = pack2(f2bf(o[hh].x), f2bf(o[hh].y));
          }
      }
    } else if (warp == PVW) {
      // V groups NVS.. : a slot is refilled when the PV warps have read its previous group
      // (fill number f of a slot waits for the phase f - 1 of its "drained" barrier: one phase per group read)
      uint32_t s = vs0, fpar = vpar0 ^ 1u;  // group NVS: the slot of group 0, one fill later
      for (int j = NVS; j < nvg; ++j) {
        mbar_wait(vempty + s * 8, fpar ^ 1u);
        attn_issue_group<AVG>(a, ring + s * AVG_BYTES, vfull + s * 8, bt, layer, 1, kvh, j, ps, lane);
        if (++s == NVS) {
          s = 0;
          fpar ^= 1u;
        }
      }
    }
    vseq += (uint32_t)nvg;
    bar_consumers();  // shared memory is reused by the next task / phase
    lap(13);
  }
  if (lane == 0) pipe[warp] = kuse;
  if (threadIdx.x == 0) {
    pipe[NW] = vseq;
    pipe[NW + 1] = 0u;
  }
}

// ---------------------------------------------------------------- the kernel
template <int NP, int MT, bool FAST>
__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_mega_kernel(const __grid_constant__ MegaArgs a) {
  constexpr int MTT = MT == 0 ? 1 : MT;
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t smem_base = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.n_layers_run > 0 ? min(a.n_layers_run, a.L) : a.L;
  const bool with_head = a.n_layers_run <= 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < a.n_slots; ++s) {
      mbar_init(smem_base + OFF_FULL + s * 8, 1);
      mbar_init(smem_base + OFF_EMPTY + s * 8, 1);
    }
    for (int s = 0; s < NW; ++s) mbar_init(smem_base + OFF_AKF + s * 8, 1);  // K/V rings of the attention phase (attention_tma_phase)
    for (int s = 0; s < NVS; ++s) {
      mbar_init(smem_base + OFF_AVF + s * 8, 1);
      mbar_init(smem_base + OFF_AVE + s * 8, PVW);
    }
    {  // A stream of the tile-split down_proj (gemm_ts): "read" barriers count the warps of this CTA that own a unit
      const int G = a.ph_G[PH_DOWN];
      const int n_c = G > 0 ? a.ph_q[PH_DOWN] + (((int)blockIdx.x / a.mtt) < a.ph_r[PH_DOWN] ? 1 : 0) : 0;
      for (int s = 0; s < 4; ++s) {
        mbar_init(smem_base + OFF_TSF + s * 8, 1);
        mbar_init(smem_base + OFF_TSE + s * 8, (uint32_t)max(1, min(NW, n_c)));
      }
    }
    *reinterpret_cast<volatile uint32_t*>(smem + OFF_ISSUED) = 0u;
    for (int i = 0; i < 16; ++i) reinterpret_cast<unsigned long long*>(smem + OFF_DBG)[i] = 0ull;
    for (int i = 0; i < MAX_SLOTS; ++i) reinterpret_cast<uint32_t*>(smem + OFF_REL)[i] = 0u;
    for (int i = 0; i < NW + 2; ++i) reinterpret_cast<uint32_t*>(smem + OFF_APIPE)[i] = 0u;
    *reinterpret_cast<uint32_t*>(smem + OFF_TSSEQ) = 0u;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  {  // layer table -> shared memory (pointer chasing through global memory costs ~1 us per phase)
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.layers);
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + OFF_LAYERS);
    for (int i = threadIdx.x; i < a.L * (int)(sizeof(MegaLayer) / 4); i += MEGA_THREADS) dst[i] = src[i];
  }
  __syncthreads();

  if (warp == NW) {  // producer warp
    producer_loop(a, smem_base, with_head ? 4 * a.L + 1 : 4 * L);
    return;
  }

  // ---- consumers
  const MegaLayer* layers = reinterpret_cast<const MegaLayer*>(smem + OFF_LAYERS);
  unsigned epoch = 0;
  RingPos ring_pos{0u, 0u, 0u};
  int prof_i = 0;
  auto stamp = [&]() {
    if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) {
      a.prof[prof_i] = globaltimer();
      a.prof[a.prof_stride + prof_i] = (unsigned long long)clock64();
    }
    ++prof_i;
  };
  const uint32_t act = smem_base + a.off_act;
  const int H = a.H, Dq = a.n_q * a.hd;
  // tensor parallel: exchanges done in this launch, residual ping-pong, generation base of the flag words
  const bool tp = a.tp_size > 1;
  unsigned xch = 0;
  const unsigned xbase = tp ? __ldcg(a.tp_epoch) : 0u;
  const bf16* x_cur = a.x;
  bf16* x_next = a.x2;
  Best best[MTT][2];
#pragma unroll
  for (int m = 0; m < MTT; ++m) best[m][0] = best[m][1] = Best{-CUDART_INF_F, -1};

  prefetch_norm_w(a, smem_base, 0, layers[0].in_ln);
  // embedding rows (embedded_matrix.cu:5-17): x[b] = E[ids[b]]; layer 0 reads E directly
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    const uint4* src = reinterpret_cast<const uint4*>(a.embed + (size_t)max(a.ids[b], 0) * H);
    uint4* dst = reinterpret_cast<uint4*>(a.x + (size_t)b * H);
    for (int i = threadIdx.x; i < (H >> 3); i += NTC) dst[i] = src[i];
  }
  if (a.kv_l2_prefetch) attention_prefetch_l2<NP>(a, 0);
  stamp();
  // One loop over the GEMM phases (4 per layer, then lm_head) with a single inlined copy of
  // every stage, so kernel arguments stay in the constant bank and the code stays small.
  const int n_idx = with_head ? 4 * a.L + 1 : 4 * L;
  for (int idx = 0; idx < n_idx; ++idx) {
    const int kind = idx < 4 * a.L ? (idx & 3) : PH_LMHEAD;
    const int l = idx >> 2;
    const bf16* a_src = nullptr;
    if (kind == PH_O) {
      // ---- attention, then O + residual
      if (NP == 1 && a.attn_tma)
        attention_tma_phase(a, l, smem);
      else if (NP == 1 && a.attn_group)
        attention_group_phase(a, l, smem);
      else
        attention_phase<NP>(a, l, smem, 1);
      stamp();
      grid_sync(a.bar, epoch);
      stamp();
      if (a.ph_ts[PH_O]) {  // tile-split: this CTA's 16-row token tile only
        const int tok0 = ((int)blockIdx.x % a.mtt) * 16;
        load_rows(act, min(16, a.B - tok0), Dq, [&](int b) { return a.att + (size_t)(tok0 + b) * Dq; });
      } else {
        load_rows(act, a.B, Dq, [&](int b) { return a.att + (size_t)b * Dq; });
      }
      stamp();
    } else if (kind == PH_DOWN) {
      // ---- down + residual
      if (a.kv_l2_prefetch && l + 1 < L) attention_prefetch_l2<NP>(a, l + 1);
      if (a.stream_down || a.ph_ts[PH_DOWN])
        a_src = a.h;
      else
        load_rows(act, a.B, a.I, [&](int b) { return a.h + (size_t)b * a.I; });
      stamp();
    } else {
      // ---- QKV / gate+up / lm_head: RMSNorm of the residual stream in front
      const int which = kind == PH_GATEUP ? 1 : 0;
      if (kind == PH_QKV)
        prefetch_norm_w(a, smem_base, 1, layers[l].post_ln);
      else if (kind == PH_GATEUP)
        prefetch_norm_w(a, smem_base, 0, l + 1 < a.L ? layers[l + 1].in_ln : a.final_norm);
      const bool from_embed = idx == 0;
      if (a.dist_norm) {
        // many rows: every CTA would repeat the whole batch's normalisation (IEEE divisions), so
        // CTA b normalises row b once, publishes it, and everybody copies the result
        const int b = blockIdx.x;
        if (b < a.B) {  // B <= gridDim.x (launcher)
          if (tp && !from_embed)
            load_row_tp(a, act, b, H, x_cur, x_next, (int)((xch - 1) & 1));
          else
            load_rows(act, 1, H, [&](int) { return from_embed ? a.embed + (size_t)max(a.ids[b], 0) * H : a.x + (size_t)b * H; });
          rmsnorm_rows(a, smem, 1, H, which);
          const uint4* src = reinterpret_cast<const uint4*>(smem + a.off_act);
          uint4* dst = reinterpret_cast<uint4*>(a.xn + (size_t)b * H);
          for (int i = threadIdx.x; i < (H >> 3); i += NTC) dst[i] = src[i];
        }
        if (tp && !from_embed) {  // every row was rewritten by its CTA
          const bf16* t = x_cur;
          x_cur = x_next;
          x_next = const_cast<bf16*>(t);
        }
        grid_sync(a.bar, epoch);
        stamp();
        if (a.ph_ts[kind == PH_LMHEAD ? PH_LMHEAD : kind]) {  // tile-split (QKV): this CTA's token tile only
          const int tok0 = ((int)blockIdx.x % a.mtt) * 16;
          load_rows(act, min(16, a.B - tok0), H, [&](int r) { return a.xn + (size_t)(tok0 + r) * H; });
        } else {
          load_rows(act, a.B, H, [&](int r) { return a.xn + (size_t)r * H; });
        }
        stamp();
      } else {
        if (tp && !from_embed) {
          if (a.B > 4) {
            combine_rows_tp(a, a.B, H, x_cur, x_next, (int)((xch - 1) & 1));
            grid_sync(a.bar, epoch);
            const bf16* xs = x_next;
            load_rows(act, a.B, H, [&](int b) { return xs + (size_t)b * H; });
          } else {
            load_rows_tp(a, act, a.B, H, x_cur, x_next, (int)((xch - 1) & 1));
          }
          const bf16* t = x_cur;
          x_cur = x_next;
          x_next = const_cast<bf16*>(t);
        } else {
          load_rows(act, a.B, H, [&](int b) { return from_embed ? a.embed + (size_t)max(a.ids[b], 0) * H : a.x + (size_t)b * H; });
        }
        stamp();
        if (FAST)
          rmsnorm_rows_fast(a, smem, a.B, H, which);
        else
          rmsnorm_rows(a, smem, a.B, H, which);
        stamp();
      }
    }
    Phase p;
    make_phase(a, idx, p);
    p.xpar = (int)(xch & 1);
    if (kind == PH_QKV && a.attn_kstg > 0)
      attention_phase<NP>(a, l, smem, 0);  // warps 1..: request what attention needs that does not depend on q/k/v
    bool done = false;
    if constexpr (FAST) {  // fast numerics: K of every unit split over the warps of the CTA
      gemm_phase_fast(a, p, smem_base, ring_pos, best, smem);
      done = true;
    }
    if (!done) gemm_phase<MT>(a, p, smem_base, ring_pos, a_src, best, smem + OFF_DBG);
    if (NP == 1 && kind == PH_QKV && a.attn_tma) {
      bar_consumers();  // the activation rows of the QKV GEMM are dead: the K ring takes their place
      attention_tma_prefetch(a, l, smem);
    }
    stamp();
    if (kind == PH_LMHEAD) {
      // candidates: lanes (g, c) of a warp hold tokens 16m + g (+8); reduce over c, then over warps
      MegaCand* cs = reinterpret_cast<MegaCand*>(smem + OFF_CAND);
#pragma unroll
      for (int m = 0; m < MTT; ++m)
#pragma unroll
        for (int hr = 0; hr < (MT == 0 ? 1 : 2); ++hr) {
          Best bb = best[m][hr];
#pragma unroll
          for (int o = 1; o <= 2; o <<= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bb.v, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bb.i, o);
            if (cand_better(ov, oi, bb.v, bb.i)) {
              bb.v = ov;
              bb.i = oi;
            }
          }
          const int tok = m * 16 + (lane >> 2) + hr * 8;
          if ((lane & 3) == 0 && tok < MAX_ROWS) cs[warp * MAX_ROWS + tok] = MegaCand{bb.v, bb.i};
        }
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
    }
    if (tp && (kind == PH_O || kind == PH_DOWN)) {
      ++xch;
      tp_exchange_sync(a, epoch, xbase + xch);
    } else {
      grid_sync(a.bar, epoch);
    }
    stamp();
  }
  if (!with_head) {
    if (tp && blockIdx.x == 0 && threadIdx.x == 0) *a.tp_epoch = xbase + xch;  // every CTA read it before its first barrier
    return;
  }
  // ---- arg-max over the CTAs' candidates + step bookkeeping (advance_kernel)
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
        if (tp) {
          // this rank's best of its vocabulary range (the low byte of the index, the tie-break key, is the same
          // in the shard and in the full row: shards start at multiples of 256) -> every rank's candidate table
          const MegaCand mine{bv, bi >= 0 ? bi + a.tp_vocab0 : -1};
          for (int r = 0; r < a.tp_size; ++r) a.tp_cand[r][a.tp_rank * MEGA_TP_ROWS + b] = mine;
        } else {
          a.sampled[b] = bi;
          if (a.advance) {
            a.pos[b] += 1;
            a.ids[b] = bi;
            if (a.rowstep) a.rowstep[b] += 1;
          }
        }
      }
    }
  }
  if (tp) {
    if (a.greedy) {
      ++xch;
      tp_exchange_sync(a, epoch, xbase + xch);
      if (threadIdx.x == 0) {
        for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
          float bv = -CUDART_INF_F;
          int bi = -1;
          for (int r = 0; r < a.tp_size; ++r) {  // rank order; cand_better is a total order, so every rank picks the same token
            const MegaCand* cp = a.tp_cand[a.tp_rank] + r * MEGA_TP_ROWS + b;
            const float ov = __ldcg(&cp->val);
            const int oi = __ldcg(&cp->idx);
            if (cand_better(ov, oi, bv, bi)) {
              bv = ov;
              bi = oi;
            }
          }
          a.sampled[b] = bi;
          if (a.advance) {
            a.pos[b] += 1;
            a.ids[b] = bi;
            if (a.rowstep) a.rowstep[b] += 1;
          }
        }
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.tp_epoch = xbase + xch;  // every CTA read it before its first barrier
  }
  stamp();
  if (a.prof && blockIdx.x == 0 && threadIdx.x == 0)
    for (int i = 0; i < 16; ++i) a.prof[2 * a.prof_stride + i] = reinterpret_cast<unsigned long long*>(smem + OFF_DBG)[i];
}

struct Geom {
  int KC, slot_bytes, act_bytes, n_slots, off_act, off_ring, mt, stream_down, kstg, attn_off, off_red, group;
  size_t smem;
};

int kc_for(int H, int big) {
  const int lim = big ? 1024 : 512;  // small batches: fewer, larger weight tiles
  const int nch = (H + lim - 1) / lim;
  int KC = ((H + nch - 1) / nch + 63) & ~63;
  if (const char* kv = getenv("QIE_MEGA_KC")) {  // tuning knob: k elements per weight tile
    const int v = atoi(kv);
    if (v >= 64 && v <= 4096 && v % 64 == 0) KC = v;
  }
  return KC;
}

bool mega_geometry(int H, int I, int L, int n_q, int n_kv, int hd, int B, int max_kv_len, int grid, int KC, int fast, int psz, Geom* g) {
  const int Dq = n_q * hd;
  if ((H % 64) || (I % 64) || (Dq % 64) || (hd != 64 && hd != 128) || KC < 64 || (KC % 64)) return false;
  if (B < 1 || B > MAX_ROWS || n_q % n_kv || n_q / n_kv > 2 * NW || L > MAX_LAYERS) return false;
  const int mt = B <= 8 ? 0 : (B <= 16 ? 1 : (B <= 32 ? 2 : 4));
  const int bpad = mt == 0 ? 8 : 16 * mt;
  const int rows_a = mt == 0 ? B : bpad;  // rows the A fragments may touch
  const int res_h = rows_a * (std::max(H, Dq) + 8) * 2;
  const int res_i = rows_a * (I + 8) * 2;
  const int strm = 2 * bpad * (KC + 8) * 2;
  if (fast && B > 8) return false;
  const int stream_down = !fast && res_i > 48 * 1024;  // fast numerics keeps [B, I] resident (<= 8 rows)
  int act = std::max(res_h, stream_down ? strm : res_i);
  if (mt >= 2) act = std::max(act, TS_STAGES * 16 * (KC + 8) * 2);  // tile-split down_proj: stages of the 16-row A stream (gemm_ts)
  const int hs = (B * n_q <= grid) ? 1 : n_q / n_kv;
  const int group = hs > 1 && hd == 64 && psz > 0 && (psz & (psz - 1)) == 0;  // query-group tasks with K and V tiles in shared memory (attention_group_phase)
  const int tmax = (max_kv_len + 3) & ~3;
  const int act_gemm = act;
  const int off_act = (OFF_WNORM + 2 * H * 2 + 127) & ~127;
  const int red = fast ? FAST_U * NW * 32 * 8 * 4 : 0;
  const int slot = 16 * KC * 2;                          // gate + up boxes of 8 rows x KC
  int kstg = hs == 1 ? std::min(256, (max_kv_len + 15) & ~15) : 0;  // K rows staged before the barrier
  int attn_off = 0, off_red = 0, off_ring = 0, S = 0;
  for (;;) {
    const int attn = group ? hs * hd * 4 + 2 * hd * 2 + hs * attn_score_pitch(max_kv_len) * 4 + 2 * AT_BYTES + (max_kv_len / std::max(1, psz) + 2) * 4 + 64
                           : hs * hd * 4 + 2 * hd * 2 + hs * tmax * 4 + (max_kv_len + 32) * 4 /* pages: <= one per position */ +
                                 2 * VT * hd * 2 + kstg * (hd * 2 + 16) + 2 * hd * 2 + hd * 4 + 64;
    attn_off = kstg > 0 ? ((res_h + 127) & ~127) : 0;  // staged attention areas live behind the resident rows
    act = std::max(act_gemm, attn + attn_off);
    if (group) act = std::max(act, attn_tma_layout(off_act, hs, max_kv_len).total);  // K/V rings fed by TMA (attention_tma_phase)
    off_red = off_act + ((act + 127) & ~127);     // fast numerics: split-K partial sums [FAST_U][NW][32][8] fp32
    off_ring = (off_red + red + 1023) & ~1023;    // swizzled TMA tiles: 1024-byte aligned slots
    const int budget = 227 * 1024 - off_ring;
    S = budget / slot;
    if (S > MAX_SLOTS) S = MAX_SLOTS;
    if (S >= 3 || kstg == 0) break;
    kstg = 0;  // the staged layout (K rows + attention areas behind the resident rows) does not fit: per-head tasks without stage 0
  }
  if (S < 3) return false;
  g->KC = KC;
  g->slot_bytes = slot;
  g->act_bytes = off_ring - off_act;
  g->n_slots = S;
  g->off_act = off_act;
  g->off_ring = off_ring;
  g->mt = mt;
  g->stream_down = stream_down;
  g->kstg = kstg;
  g->attn_off = attn_off;
  g->off_red = off_red;
  g->group = group;
  g->smem = (size_t)off_ring + (size_t)S * slot;
  return true;
}

}  // namespace

int decode_mega_kc(int H, int big) { return kc_for(H, big); }

bool decode_mega_supports(int H, int I, int L, int n_q, int n_kv, int hd, int B, int max_kv_len, int num_sms, int KC, int fast, int page_size) {
  Geom g;
  return mega_geometry(H, I, L, n_q, n_kv, hd, B, max_kv_len, num_sms, KC, fast, page_size, &g);
}

int decode_mega_prof_slots(int L) { return 2 * (16 * L + 8) + 16; }

template <int NP>
static void (*pick_kernel(int mt, bool fast))(MegaArgs) {
  if (fast) return decode_mega_kernel<NP, 0, true>;  // fast numerics: <= 8 rows only
  switch (mt) {
    case 0: return decode_mega_kernel<NP, 0, false>;
    case 1: return decode_mega_kernel<NP, 1, false>;
    case 2: return decode_mega_kernel<NP, 2, false>;
    default: return decode_mega_kernel<NP, 4, false>;
  }
}

cudaError_t launch_decode_mega(MegaArgs a, int num_sms, cudaStream_t st) {
  Geom g;
  if (!mega_geometry(a.H, a.I, a.L, a.n_q, a.n_kv, a.hd, a.B, a.max_kv_len, num_sms, a.KC, a.fast, a.kv.page_size, &g)) return cudaErrorInvalidValue;
  a.slot_bytes = g.slot_bytes;
  a.act_bytes = g.act_bytes;
  a.n_slots = g.n_slots;
  a.off_act = g.off_act;
  a.off_ring = g.off_ring;
  a.stream_down = g.stream_down;
  a.attn_kstg = g.kstg;
  a.attn_off = g.attn_off;
  a.off_red = g.off_red;
  a.attn_group = g.group;
  a.attn_hp = 1;
  a.attn_tma = 0;
  if (g.group) {  // heads per attention task: the smallest part size whose tasks fit one wave
    const int hsk = a.n_q / a.n_kv;
    int hp = 1;
    while (hp < hsk && a.B * a.n_kv * ((hsk + hp - 1) / hp) > num_sms) ++hp;
    a.attn_hp = hp;
    // K/V streamed by TMA: needs the pool's tensor map (head_dim 64, pages of >= 8 slots) and <= 2 heads per PV warp
    a.attn_tma = a.kvmap != nullptr && a.kv.page_size >= 8 && hp <= 2 * PVW;
    static const bool tma_env = [] { const char* v = getenv("QIE_MEGA_ATMA"); return !(v && v[0] == '0'); }();  // A/B knob
    if (!tma_env) a.attn_tma = 0;
  }
  a.dist_norm = a.B > 16 && a.B <= num_sms && a.xn != nullptr;
  a.mtt = g.mt == 0 ? 1 : g.mt;
  // tile split needs the tiles to cover the rows it is given (mtt = 2: 17..32 rows, 4: 33..64) and no tensor parallel
  // exchange in the epilogue (kept on the default mapping); QIE_MEGA_TS=0 switches it off (A/B knob)
  static const bool ts_env = [] { const char* v = getenv("QIE_MEGA_TS"); return !(v && v[0] == '0'); }();
  const bool ts_on = ts_env && a.tp_size <= 1 && !a.fast;
  {  // per phase kind: units of a CTA (q, +1 for the first r CTAs), chunks, and ring advances modulo the slot count
    const int Dq = a.n_q * a.hd, Dkv = a.n_kv * a.hd;
    const int units[5] = {(Dq + 7) / 8 + 2 * ((Dkv + 7) / 8), (a.H + 7) / 8, (a.I + 7) / 8, (a.H + 7) / 8, (a.V + 15) / 16};
    const int Ks[5] = {a.H, Dq, a.H, a.I, a.H};
    const int S = g.n_slots;
    for (int k = 0; k < 5; ++k) {
      a.ph_q[k] = units[k] / num_sms;
      a.ph_r[k] = units[k] % num_sms;
      a.ph_nch[k] = (Ks[k] + a.KC - 1) / a.KC;
      for (int cls = 0; cls < 2; ++cls) {
        const int jobs = (a.ph_q[k] + cls) * a.ph_nch[k];
        a.ph_adv_slot[k][cls] = jobs % S;
        a.ph_adv_par[k][cls] = (jobs / S) & 1;
      }
      // phases with at most one unit per CTA spread the token tiles of that unit over MT warps -- or, where the
      // normalised rows come from the row-per-CTA path (dist_norm) / plain buffers, over CTAs (tile split, gemm_ts)
      const int mtt = g.mt == 0 ? 1 : g.mt;
      const bool split = mtt >= 2 && units[k] <= num_sms && k != 4;
      const bool ts = split && ts_on && (k == PH_O || k == PH_DOWN || (k == PH_QKV && a.dist_norm));
      a.ph_ts[k] = ts ? 1 : 0;
      a.ph_G[k] = ts ? num_sms / mtt : 0;
      if (ts) {
        const int G = a.ph_G[k];
        a.ph_q[k] = units[k] / G;
        a.ph_r[k] = units[k] % G;
        for (int cls = 0; cls < 2; ++cls) {
          const int jobs = (a.ph_q[k] + cls) * a.ph_nch[k];
          a.ph_adv_slot[k][cls] = jobs % S;
          a.ph_adv_par[k][cls] = (jobs / S) & 1;
        }
      }
      a.ph_nu[k] = ts ? NW : (split ? NW / mtt : NW);
      if (a.fast)  // split-K rounds of <= FAST_U units
        a.ph_nu[k] = std::max(1, std::min(FAST_U, (units[k] + num_sms - 1) / num_sms));
      const int rj = a.ph_nu[k] * a.ph_nch[k];
      a.ph_round_slot[k] = rj % S;
      a.ph_round_par[k] = (rj / S) & 1;
    }
  }
  void (*kern)(MegaArgs) = a.hd == 64 ? pick_kernel<1>(g.mt, a.fast != 0) : pick_kernel<2>(g.mt, a.fast != 0);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  void* params[] = {&a};
  return cudaLaunchCooperativeKernel((const void*)kern, dim3(num_sms), dim3(MEGA_THREADS), params, g.smem, st);
}

}  // namespace qie
