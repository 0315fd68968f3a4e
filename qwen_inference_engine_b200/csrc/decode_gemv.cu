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

// consumer warps.  7 + the producer = 8 warps = two per SM sub-partition: 255 registers per thread are available and
// ptxas uses 242 without spills.  With 8 + 1 one sub-partition hosts three warps, which caps EVERY thread at 168
// registers (256-512 bytes of spills in the GEMV loops): batch 1 2248 -> 2332 tok/s, batch 4 / ctx 1000 3748 -> 3952.
constexpr int GV_CW = 7;
constexpr int GV_CT = GV_CW * 32;       // consumer threads
constexpr int GV_THREADS = GV_CT + 32;  // + the producer warp
constexpr int GV_MAX_SLOTS = 16;
#ifndef GV_R_SMALL
#define GV_R_SMALL 2  // weight rows per item in the QKV / o / down phases (measured: 1 = more warps busy but ctx 2048 15 % slower)
#endif
constexpr int GV_PT = 2048;             // cached positions one attention task takes
constexpr int GV_MAX_LAYERS = 64;
constexpr int GV_NVJ = 2;               // 16-byte vectors of a residual row per thread: H <= 8 * 256 * 2
constexpr int GV_NPJ = 4 * GV_NVJ;      // = bf16 pairs per thread (words 4 v .. 4 v + 3 of vector v = thread + 256 jv)
constexpr int GV_SMEM_MAX = 227 * 1024;
constexpr long long GV_SPIN_LIMIT = 4000000000ll;  // ~2 s of SM clocks: trap instead of hanging the GPU

struct GemvArgs {
  int n_slots, slot_bytes, off_act, off_pf, off_kv, off_ring;  // off_kv: 2 x GV_KVS bytes, the task's first K / V rows
  int n_split;   // KV splits per (row, q head)
  int dataflow;  // 1 (default): no grid barrier between the phases, consumers poll the per-layer buffers (batch 1: 1780 ->
                 // 2150 tok/s); 0 (QIE_GEMV_DATAFLOW=0): grid barriers between the phases, every layer reuses one block
  bf16* act;     // per-layer activation buffers, every element 0xFFFF until its value is stored (GvAct)
};

// ---------------------------------------------------------------- data-flow synchronisation
// The layers are NOT separated by grid barriers.  Every activation a phase hands to the next one has its own buffer PER
// LAYER (x_in, qkv, attention output / split partials, x_mid, h: ~0.4 MB per sequence for the 0.5B shape), filled
// with the bf16 pattern 0xFFFF (a NaN no operator stores) before the kernel starts.  A consumer polls exactly the
// elements it needs until they stop being 0xFFFF: one L2 round trip behind the producer's store instead of
// store -> fence -> atomic -> poll -> CTA barrier (measured ~1.9 us per grid barrier, 5 per layer, with every CTA
// waiting for the slowest).  One real grid barrier remains (arg-max candidates), then all CTAs put the pattern back.
struct GvAct {
  bf16* base;
  size_t ls;          // elements per layer
  int oq, oa, om, oh, op;
  __device__ __forceinline__ bf16* xi(int l) const { return base + (size_t)l * ls; }           // [B][H] input of layer l (l = L: final)
  __device__ __forceinline__ bf16* qkv(int l) const { return xi(l) + oq; }                       // [B][Dq + 2 Dkv]
  __device__ __forceinline__ bf16* att(int l) const { return xi(l) + oa; }                       // [B][Dq]
  __device__ __forceinline__ bf16* xm(int l) const { return xi(l) + om; }                        // [B][H] behind o_proj
  __device__ __forceinline__ bf16* hh(int l) const { return xi(l) + oh; }                        // [B][I]
  __device__ __forceinline__ float* part(int l) const { return reinterpret_cast<float*>(xi(l) + op); }  // [tasks][hd + GV_PART]
};
__host__ __device__ inline size_t gv_act_layer_elems(int B, int H, int I, int Dq, int Dkv, int hd, int ntask_part) {
  return (size_t)B * (2 * H + (Dq + 2 * Dkv) + Dq + I) + (((size_t)ntask_part * (hd + 4) * 2 + 7) & ~(size_t)7);
}
__device__ __forceinline__ GvAct gv_act(const MegaArgs& a, const GemvArgs& g) {
  const int Dq = a.n_q * a.hd, Dkv = a.n_kv * a.hd;
  GvAct A;
  A.base = g.act;
  A.oq = a.B * a.H;
  A.oa = A.oq + a.B * (Dq + 2 * Dkv);
  A.om = A.oa + a.B * Dq;
  A.oh = A.om + a.B * a.H;
  A.op = A.oh + a.B * a.I;
  // with grid barriers every layer reuses the first block (hot in L2; a store into a line that L2 no longer holds costs a
  // DRAM fill in front of the reader); the pattern is not needed then, the polls succeed on the first request
  A.ls = g.dataflow ? gv_act_layer_elems(a.B, a.H, a.I, Dq, Dkv, a.hd, g.n_split > 1 ? a.B * a.n_q * g.n_split : 0) : 0;
  return A;
}
__device__ __forceinline__ bool gv_word_ready(uint32_t w) { return (w & 0xffffu) != 0xffffu && (w >> 16) != 0xffffu; }
__device__ __forceinline__ uint32_t gv_ld_relaxed(const void* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 gv_ld_relaxed_v4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// wait until both bf16 halves of a word have been stored
__device__ __forceinline__ uint32_t gv_poll_u32(const void* p) {
  uint32_t v = gv_ld_relaxed(p);
  if (gv_word_ready(v)) return v;
  const long long t0 = clock64();
  do {
    // (no pause: __nanosleep costs ~1 us per call here)
    v = gv_ld_relaxed(p);
    if (clock64() - t0 > GV_SPIN_LIMIT) __trap();
  } while (!gv_word_ready(v));
  return v;
}

// shared-memory header
constexpr int GO_FULL = 0, GO_EMPTY = 128, GO_RED = 256, GO_DBG = 512, GO_CAND = 640, GO_RANGE = 1152, GO_ZERO = 1248, GO_LAYERS = 1280;  // GO_DBG: 16 x u64 cycle counters (profiled launches)
constexpr int GO_XS = GO_LAYERS + GV_MAX_LAYERS * (int)sizeof(MegaLayer);  // residual rows [B][H] bf16
// attention scratch inside the activation area
constexpr int GA_Q = 0, GA_K = 256, GA_V = 512, GA_SCORE = 768, GA_PV = GA_SCORE + GV_PT * 4, GA_END = GA_PV + GV_CW * 128 * 4;
// staged in front of the QKV phase (outside the activation area): cos + sin row, q/k-norm weights, the task's page list
constexpr int GP_COS = 0, GP_NW = 512, GP_PAGES = 1024, GP_END = GP_PAGES + (GV_PT + 16) * 4;
constexpr int GV_MAX_SPLIT = 12;
constexpr int GV_PART = 4;  // floats in front of a split's partial output row: max, sum, 2 pad (the row is read as 16-byte vectors)
#ifndef GV_STAGE_MAX_KV_
#define GV_STAGE_MAX_KV_ 256
#endif
constexpr int GV_STAGE_MAX_KV = GV_STAGE_MAX_KV_;  // contexts (kv bucket) up to which the first K / V rows are staged in shared memory
constexpr int GV_KVS = 4 * GV_CW * 4 * 64 * 2;  // bytes of staged K (and of V) rows per task: 4 * NG positions = 112 at head_dim 64, 56 at 128

__device__ __forceinline__ void mbar_init(uint32_t addr, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(addr) : "memory");
}
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
// The weights are read exactly once per step: marked evict-first in L2, so that the ~1 GB stream does not push out what
// IS re-read (activation buffers, norm weights, block tables, the KV rows of the previous tokens).
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar, uint64_t policy) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(GV_CT) : "memory"); }
__device__ __forceinline__ unsigned long long globaltimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// grid barrier of the consumer warps (see decode_mega.cu grid_sync: release add + acquire poll)
// split in two: loads that the next phases need and that do not depend on other CTAs are issued BETWEEN the arrival and
// the wait -- in front of the arrival they would sit in front of the release fence (measured: +1 us per barrier)
__device__ __forceinline__ void grid_arrive(unsigned* ctr, unsigned& epoch) {
  bar_consumers();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
  }
}
__device__ __forceinline__ void grid_wait(unsigned* ctr, unsigned epoch) {
  if (threadIdx.x == 0) {
    unsigned v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (clock64() - t0 > GV_SPIN_LIMIT) __trap();
    } while (v < epoch);
  }
  bar_consumers();
}
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
__device__ __forceinline__ void gv_phase(const MegaArgs& a, const unsigned char* smem, const MegaLayer* layers, int kind, int l, GPhase& p) {
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
  p.ch = reinterpret_cast<const int*>(smem + GO_RANGE)[10 + kind];
}
// rows [r0, r1) of a phase kind that CTA c owns (contiguous, sizes differ by at most one).  The same for every layer:
// computed once at kernel start (two 64-bit divisions per call cost ~1 us per phase on the token's critical path).
__device__ __forceinline__ void gv_range_init(const MegaArgs& a, const GemvArgs& g, unsigned char* smem) {
  if (threadIdx.x < 5) {
    const int Dq = a.n_q * a.hd, Dkv = a.n_kv * a.hd;
    const int k = threadIdx.x;
    const long long rows = k == PH_QKV ? Dq + 2 * Dkv : (k == PH_O || k == PH_DOWN ? a.H : (k == PH_GATEUP ? a.I : a.V));
    int* rg = reinterpret_cast<int*>(smem + GO_RANGE);
    rg[2 * k] = (int)((long long)blockIdx.x * rows / gridDim.x);
    rg[2 * k + 1] = (int)((long long)(blockIdx.x + 1) * rows / gridDim.x);
    const int K = k == PH_O ? Dq : (k == PH_DOWN ? a.I : a.H);
    rg[10 + k] = max(1, g.slot_bytes / (K * 2 * (k == PH_GATEUP ? 2 : 1)));  // rows (row pairs) per ring slot
  }
}
__device__ __forceinline__ void gv_range(const unsigned char* smem, int kind, int& r0, int& r1) {
  const int* rg = reinterpret_cast<const int*>(smem + GO_RANGE);
  r0 = rg[2 * kind];
  r1 = rg[2 * kind + 1];
}

// ---------------------------------------------------------------- producer: the weight stream of this CTA
__device__ __forceinline__ void gv_producer(const MegaArgs& a, const GemvArgs& g, const unsigned char* smem, uint32_t smem_base,
                                            const MegaLayer* layers) {
  if ((threadIdx.x & 31) != 0) return;
  uint32_t slot = 0, round = 0;
  const uint64_t pol = l2_policy_evict_first();
  const int nph = 4 * a.L + 1;
  for (int idx = 0; idx < nph; ++idx) {
    const int kind = idx < 4 * a.L ? (idx & 3) : PH_LMHEAD;
    GPhase p;
    gv_phase(a, smem, layers, kind, idx >> 2, p);
    int r0, r1;
    gv_range(smem, kind, r0, r1);
    const uint32_t row_bytes = (uint32_t)p.K * 2u;
    for (int r = r0; r < r1; r += p.ch) {
      const int n = min(p.ch, r1 - r);
      if (round > 0) mbar_wait(smem_base + GO_EMPTY + slot * 8, (round - 1) & 1);
      const uint32_t dst = smem_base + g.off_ring + slot * g.slot_bytes;
      const uint32_t full = smem_base + GO_FULL + slot * 8;
      mbar_expect_tx(full, (uint32_t)n * row_bytes * (p.w2 ? 2u : 1u));
      if (p.w2) {
        bulk_g2s(dst, p.seg[0] + (size_t)r * p.K, (uint32_t)n * row_bytes, full, pol);
        bulk_g2s(dst + (uint32_t)n * row_bytes, p.w2 + (size_t)r * p.K, (uint32_t)n * row_bytes, full, pol);
      } else {
        int s0 = 0;
        for (int s = 0; s < p.nseg; ++s) {
          const int s1 = s0 + p.seg_rows[s];
          const int lo = max(r, s0), hi = min(r + n, s1);
          if (lo < hi)
            bulk_g2s(dst + (uint32_t)(lo - r) * row_bytes, p.seg[s] + (size_t)(lo - s0) * p.K, (uint32_t)(hi - lo) * row_bytes, full, pol);
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
  int rot;  // items dealt so far modulo the consumer warps: item i of a chunk goes to warp (rot + i) % GV_CW
};

// ---------------------------------------------------------------- consumers: one GEMV phase
// act: [NB][K] bf16 in shared memory.  KIND selects the epilogue.
// A warp owns an ITEM = R weight rows that share the activation loads (gate/up: two gate rows and their two up rows;
// lm_head: four neighbouring rows; else two): 32 lanes x 16 bytes per step, the loads of G steps issued in one batch
// in front of their FMAs.  The step is bound by the length of a warp's dependent instruction stream, not by issue
// slots or shared-memory bandwidth, so the per-row overhead is what counts: lanes beyond the row read the
// weights of step 0 against a block of zeros instead of being predicated (no register clearing), and the R sums of a
// lane are reduced together -- every exchange halves the values a lane carries; in the end lanes 8 g + i hold row g.
template <int NB, int KIND>
__device__ __forceinline__ void gv_gemv(const MegaArgs& a, const GemvArgs& g, const GPhase& p, unsigned char* smem, uint32_t smem_base,
                                        GvRing& ring, const unsigned char* act, GvBest& best, bf16* out) {
  constexpr bool PAIR = KIND == PH_GATEUP;
  constexpr int R = (PAIR || KIND == PH_LMHEAD) ? 4 : GV_R_SMALL;  // weight rows per item
  constexpr int RI = PAIR ? 2 : R;                         // output rows (row pairs) per item
  constexpr int G = R * NB > 4 ? 2 : 4;
  static_assert(R == 1 || R == 2 || R == 4, "reduction below is written for 1, 2 or 4 rows per item");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.K, nv = K >> 3;
  const uint32_t row_bytes = (uint32_t)K * 2u;
  const bf16* xs = reinterpret_cast<const bf16*>(smem + GO_XS);
  const uint4* zero4 = reinterpret_cast<const uint4*>(smem + GO_ZERO);
  int r0, r1;
  gv_range(smem, KIND, r0, r1);
#ifdef GV_PROFILE_DETAIL  // cycles a warp waits for weights / spends in the phase (CTA 0): costs code on the hot path
  const bool dbg_on = a.prof && blockIdx.x == 0 && threadIdx.x == 0;
  long long t_in = 0, t_wait = 0;
  if (dbg_on) t_in = clock64();
#endif
  for (int r = r0; r < r1; r += p.ch) {
    const int n = min(p.ch, r1 - r);
#ifdef GV_PROFILE_DETAIL
    if (a.prof) {  // warp-uniform: a lane on its own path would be measured apart from the lanes that really wait
      const long long t0 = clock64();
      mbar_wait(smem_base + GO_FULL + ring.slot * 8, ring.round & 1);
      t_wait += clock64() - t0;
    } else
#endif
      mbar_wait(smem_base + GO_FULL + ring.slot * 8, ring.round & 1);
    const unsigned char* base = smem + g.off_ring + ring.slot * g.slot_bytes;
    const int items = (n + RI - 1) / RI;
    int it = warp - ring.rot;
    if (it < 0) it += GV_CW;
#ifdef GV_EXP_SKIP
    if (KIND == PH_GATEUP) it = items;
#endif
    for (; it < items; it += GV_CW) {
      // rows of the item inside the chunk (a short tail repeats the last row, results dropped)
      const uint4* wr[R];
#pragma unroll
      for (int q = 0; q < R; ++q) {
        const int row = PAIR ? ((q & 1) ? n : 0) + min(RI * it + (q >> 1), n - 1) : min(R * it + q, n - 1);
        wr[q] = reinterpret_cast<const uint4*>(base + (size_t)row * row_bytes);
      }
      float c[R][NB][4];
#pragma unroll
      for (int q = 0; q < R; ++q)
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int k = 0; k < 4; ++k) c[q][b][k] = 0.f;
      for (int v0 = lane; v0 < nv; v0 += 32 * G) {
        uint4 w[R][G];
        int vv[G];
        bool ok[G];
#pragma unroll
        for (int u = 0; u < G; ++u) {
          ok[u] = v0 + 32 * u < nv;
          vv[u] = ok[u] ? v0 + 32 * u : lane;
#pragma unroll
          for (int q = 0; q < R; ++q) w[q][u] = wr[q][vv[u]];
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          uint4 x[G];
#pragma unroll
          for (int u = 0; u < G; ++u) x[u] = *(ok[u] ? reinterpret_cast<const uint4*>(act + (size_t)b * row_bytes) + vv[u] : zero4);
#pragma unroll
          for (int u = 0; u < G; ++u)
#pragma unroll
            for (int q = 0; q < R; ++q) dot8(c[q][b], w[q][u], x[u]);
        }
      }
      // reduce R sums per lane: exchanges at distance 16 (and 8 for R = 4) halve the values carried, then a butterfly
      float s[NB];
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float t[R];
#pragma unroll
        for (int q = 0; q < R; ++q) t[q] = (c[q][b][0] + c[q][b][1]) + (c[q][b][2] + c[q][b][3]);
        const bool h16 = (lane & 16) != 0, h8 = (lane & 8) != 0;
        if (R == 4) {
          const float k0 = h16 ? t[2] : t[0], k1 = h16 ? t[3] : t[1];
          const float g0 = h16 ? t[0] : t[2], g1 = h16 ? t[1] : t[3];
          const float u0 = k0 + __shfl_xor_sync(0xffffffffu, g0, 16), u1 = k1 + __shfl_xor_sync(0xffffffffu, g1, 16);
          s[b] = (h8 ? u1 : u0) + __shfl_xor_sync(0xffffffffu, h8 ? u0 : u1, 8);
        } else if (R == 2) {
          const float u0 = (h16 ? t[R - 1] : t[0]) + __shfl_xor_sync(0xffffffffu, h16 ? t[0] : t[R - 1], 16);
          s[b] = u0 + __shfl_xor_sync(0xffffffffu, u0, 8);
        } else {
          const float u0 = t[0] + __shfl_xor_sync(0xffffffffu, t[0], 16);
          s[b] = u0 + __shfl_xor_sync(0xffffffffu, u0, 8);
        }
      }
#pragma unroll
      for (int o = 4; o > 0; o >>= 1)
#pragma unroll
        for (int b = 0; b < NB; ++b) s[b] += __shfl_xor_sync(0xffffffffu, s[b], o);
      // R = 4: lanes 8 q + i hold weight row q; R = 1: every lane holds the sum.  Lane i < B of a group finishes batch row i.
      float v1 = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b)
        if ((lane & 7) == b) v1 = s[b];
      float v2 = 0.f;
      if (PAIR) v2 = __shfl_xor_sync(0xffffffffu, v1, 8);  // the up_proj sum next to the gate_proj sum
      const int q = R == 4 ? lane >> 3 : (R == 2 ? lane >> 4 : 0);  // weight row of the item this lane finishes
      const int orow = PAIR ? RI * it + (q >> 1) : R * it + q;  // output row (pair) inside the chunk
      const bool mine = (R == 4 ? (lane & 7) : (R == 2 ? (lane & 15) : lane)) < a.B && orow < n && (!PAIR || (q & 1) == 0);
      if (mine) {
        const int b = lane & 7, gr = r + orow;
        const float y = bf2f(f2bf(v1));  // the projection output as the reference stores it (matrix_mul.cu: bf16)
        if (KIND == PH_QKV) {
          out[(size_t)b * p.rows + gr] = f2bf(v1);
        } else if (KIND == PH_O || KIND == PH_DOWN) {
          out[(size_t)b * a.H + gr] = f2bf(bf2f(xs[b * a.H + gr]) + y);  // residual_add.cu:7
        } else if (KIND == PH_GATEUP) {
          const float sg = bf2f(f2bf(y * (1.0f / (1.0f + expf(-y)))));     // SiLU.cu:6-8, stored as bf16
          out[(size_t)b * a.I + gr] = f2bf(sg * bf2f(f2bf(v2)));            // element_add.cu (element-wise product)
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
    ring.rot = (ring.rot + items) % GV_CW;
    if (++ring.slot == (uint32_t)g.n_slots) {
      ring.slot = 0;
      ++ring.round;
    }
  }
#ifdef GV_PROFILE_DETAIL
  if (dbg_on) {
    unsigned long long* dbg = reinterpret_cast<unsigned long long*>(smem + GO_DBG);
    dbg[2 * KIND] += (unsigned long long)t_wait;
    dbg[2 * KIND + 1] += (unsigned long long)(clock64() - t_in);
  }
#endif
}

// norm weights of the next RMSNorm into registers (requested in front of the grid barrier that precedes their use)
__device__ __forceinline__ void gv_norm_w(const bf16* w, int H, uint32_t (&wr)[GV_NPJ]) {
  const uint4* wp = reinterpret_cast<const uint4*>(w);
#pragma unroll
  for (int jv = 0; jv < GV_NVJ; ++jv) {
    const int v = threadIdx.x + jv * GV_CT;
    const uint4 t = v < (H >> 3) ? __ldg(wp + v) : make_uint4(0u, 0u, 0u, 0u);
    wr[4 * jv] = t.x, wr[4 * jv + 1] = t.y, wr[4 * jv + 2] = t.z, wr[4 * jv + 3] = t.w;
  }
}
// residual rows -> shared memory (raw, for the residual epilogues) + their RMSNorm (normalization.cu:9-21 rounding:
// bf16((x / rms) * w)) into the activation area; the sum of squares is a tree, not the reference's chain
template <int NB>
__device__ __forceinline__ void gv_load_norm(const MegaArgs& a, unsigned char* smem, const GemvArgs& g, const bf16* rows,
                                             const uint32_t (&wr)[GV_NPJ]) {  // rows == nullptr: the embedding rows of a.ids
  const int H = a.H, hv = H >> 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4 xv[NB][GV_NVJ];
  float ss[NB];
  // Polling discipline: a thread owns 16-byte vectors of a row (every CTA polls every row: with 4-byte polls the 148 x
  // B x H / 2 requests on the row's few L2 lines were served one after the other -- the wait grew with the row count);
  // every vector of a thread is requested in one round (the checks sit behind ALL requests: a check next to its load
  // serialises the round trips) and the round is repeated while a word still shows the pattern.
  // Measured alternatives: every word polled in turn (a round trip per word: 4 rows cost 7 us), one lane per warp
  // polling a representative word first (+1 round trip, no gain), a pause between rounds (__nanosleep costs ~1 us).
  {
    const long long t0 = clock64();
    bool ok;
    do {
      ok = true;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        ss[b] = 0.f;
        const bool live = b < a.B;
        const uint4* src =
            reinterpret_cast<const uint4*>(rows ? rows + (size_t)(live ? b : 0) * H : a.embed + (size_t)(live ? max(a.ids[b], 0) : 0) * H);
#pragma unroll
        for (int jv = 0; jv < GV_NVJ; ++jv) {
          const int v = threadIdx.x + jv * GV_CT;
          xv[b][jv] = make_uint4(0u, 0u, 0u, 0u);
          if (live && v < hv) xv[b][jv] = rows ? gv_ld_relaxed_v4(src + v) : __ldcg(src + v);
        }
      }
      if (rows) {  // checked behind ALL requests of the round
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int jv = 0; jv < GV_NVJ; ++jv)
            if (b < a.B && threadIdx.x + jv * GV_CT < hv)
              ok = ok && gv_word_ready(xv[b][jv].x) && gv_word_ready(xv[b][jv].y) && gv_word_ready(xv[b][jv].z) && gv_word_ready(xv[b][jv].w);
      }
      if (!ok && clock64() - t0 > GV_SPIN_LIMIT) __trap();
    } while (!ok);
  }
  uint4* xs = reinterpret_cast<uint4*>(smem + GO_XS);
  float* red = reinterpret_cast<float*>(smem + GO_RED);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
#pragma unroll
    for (int jv = 0; jv < GV_NVJ; ++jv) {
      const int v = threadIdx.x + jv * GV_CT;
      if (v < hv) {
        xs[b * hv + v] = xv[b][jv];
        const uint32_t w4[4] = {xv[b][jv].x, xv[b][jv].y, xv[b][jv].z, xv[b][jv].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float lo = lo2f(w4[c]), hi = hi2f(w4[c]);
          ss[b] += lo * lo + hi * hi;
        }
      }
    }
    ss[b] = warp_sum(ss[b]);
    if (lane == 0) red[b * GV_CW + warp] = ss[b];
  }
  bar_consumers();
  uint4* xn = reinterpret_cast<uint4*>(smem + g.off_act);
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float tot = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < GV_CW; ++w8) tot += red[b * GV_CW + w8];
    const float rms = sqrtf(tot / (float)H + 1e-04f);
#pragma unroll
    for (int jv = 0; jv < GV_NVJ; ++jv) {
      const int v = threadIdx.x + jv * GV_CT;
      if (v < hv) {
        const uint32_t w4[4] = {xv[b][jv].x, xv[b][jv].y, xv[b][jv].z, xv[b][jv].w};
        uint32_t o4[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
          o4[c] = pack2(f2bf(__fdividef(lo2f(w4[c]), rms) * lo2f(wr[4 * jv + c])), f2bf(__fdividef(hi2f(w4[c]), rms) * hi2f(wr[4 * jv + c])));
        xn[b * hv + v] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
  bar_consumers();
}

// ---------------------------------------------------------------- attention: split-KV flash decoding
struct GvTask {
  int b, h, sp, kvh, ps, p0, p1, pc1, pg0;
  int page_new;  // pool page of the position this step appends (set for the CTA's staged task at the start of the step)
  bool has_new;
};
template <int NP>
__device__ __forceinline__ bool gv_task(const MegaArgs& a, const GemvArgs& g, int task, int psh, GvTask& t) {
  const int S = g.n_split, Gq = a.n_q / a.n_kv, psz = a.kv.page_size;
  t.b = task / (a.n_q * S);
  const int rem = task - t.b * a.n_q * S;
  t.h = rem / S;
  t.sp = rem - t.h * S;
  t.kvh = t.h / Gq;
  t.ps = a.pos[t.b];
  const int npos = t.ps + 1, len = (npos + S - 1) / S;
  t.p0 = min(npos, t.sp * len);
  t.p1 = min(npos, t.p0 + len);
  t.has_new = t.p1 == npos;               // the split holds the position this step appends
  t.pc1 = t.has_new ? t.p1 - 1 : t.p1;    // cached positions [p0, pc1)
  t.pg0 = psh >= 0 ? (t.p0 >> psh) : t.p0 / psz;
  return t.p0 < t.p1;
}
// Everything of a task that does not depend on this step's q / k / v, staged outside the activation area:
//   ST_STEP  (once per kernel for the CTA's task when every CTA has at most one): page list, cos / sin row
//   ST_LAYER (in front of the QKV phase of every layer): q/k-norm weights, and the cached K / V rows requested into L2 --
//            they were written a token ago and ~1 GB of weights has passed through L2 since: a plain load behind the
//            grid barrier would be a DRAM + TLB miss on the token's critical path
enum { ST_STEP = 1, ST_LAYER = 2, ST_PREFETCH = 4 };
template <int NP, bool KS>
__device__ __forceinline__ void gv_attn_stage(const MegaArgs& a, const GemvArgs& g, int layer, unsigned char* smem, const MegaLayer& w,
                                              const GvTask& t, int psh, int what) {
  constexpr int HD = 64 * NP;
  const int psz = a.kv.page_size;
  unsigned char* pf = smem + g.off_pf;
  float* cs = reinterpret_cast<float*>(pf + GP_COS);
  uint32_t* nw = reinterpret_cast<uint32_t*>(pf + GP_NW);
  int* pages = reinterpret_cast<int*>(pf + GP_PAGES);
  const int tid = threadIdx.x;
  const int pgn = t.pc1 > t.p0 ? (psh >= 0 ? ((t.pc1 - 1) >> psh) : (t.pc1 - 1) / psz) - t.pg0 + 1 : 0;
  if (what & ST_STEP) {
    if (tid < 64 * NP) {  // cos row then sin row: 32 * NP floats each
      const int k = tid < 32 * NP ? tid : tid - 32 * NP;
      cs[tid] = __ldg((tid < 32 * NP ? a.cos_t : a.sin_t) + (size_t)t.ps * 32 * NP + k);
    }
    const int* bt = a.block_table + (size_t)a.slot[t.b] * a.max_pages;
    for (int i = tid; i < pgn; i += GV_CT) pages[i] = bt[t.pg0 + i];
  }
  if (what & ST_LAYER) {
    // q_norm (warp 0) / k_norm (warp 1) weights as bf16 pairs, 32 * NP words each: asynchronous copies (a register
    // round trip would stall this thread for a DRAM latency in front of the CTA barrier that follows); lane i copies
    // exactly the words lane i reads in head_norm, so a cp.async.wait_all of the reader is all the ordering needed
    if (tid < 64) {
      const int wq = tid >> 5, ln = tid & 31;
      const bf16* src = wq == 0 ? w.q_norm : w.k_norm;
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        uint32_t* dst = nw + wq * 32 * NP + 32 * p + ln;
        if (src)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(reinterpret_cast<const uint32_t*>(src) + 32 * p + ln) : "memory");
        else
          *dst = 0u;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  if (what & ST_PREFETCH) {  // needs the page list in shared memory (ST_STEP + a barrier before)
    const int rows = t.pc1 - t.p0, per = HD / 64;  // 128-byte lines per row; the first GV_KVS bytes are staged (gv_pre_stage)
    for (int i = tid + (KS ? (GV_KVS / 128) * 2 : 0); i < rows * per * 2; i += GV_CT) {
      const int kv = i & 1, j = i >> 1, rr = j / per, ln = j - rr * per, pp = t.p0 + rr;
      const int pi = psh >= 0 ? (pp >> psh) : pp / psz;
      const bf16* src = a.kv.chunk(pages[pi - t.pg0], layer, kv, t.kvh) + (size_t)(pp - pi * psz) * HD + ln * 64;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
    }
  }
}

// the first 4 * NG cached positions of a task's K and V rows, 16 bytes per lane and position: element offsets inside
// (layer 0, K) computed once per step.  The rows are copied into shared memory with cp.async a phase ahead (behind the
// down_proj phase of the layer before): they were written a token ago and ~1 GB of weights has passed through L2 since,
// so a load at the start of the task -- even behind an L2 prefetch -- was still on the task's critical path.  Every lane
// copies exactly the 16-byte pieces it reads itself: cp.async.wait_all of the reader is all the ordering needed, and
// nothing is held in registers across the phases in between.
template <int NP>
struct GvPre {
  size_t off[4];   // element offset of this lane's piece of position base + u * NG + grp; ~0 = beyond the task
  uint4 k[4], v[4];  // long contexts (no staging area, GemvArgs::off_kv == 0): the rows in registers
};
template <int NP>
__device__ __forceinline__ void gv_pre_offsets(const MegaArgs& a, const int* pages, const GvTask& t, int psh, GvPre<NP>& pre) {
  constexpr int HD = 64 * NP, LPP = HD / 8, GPW = 32 / LPP, NG = GV_CW * GPW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp * GPW + lane / LPP, j = lane % LPP, psz = a.kv.page_size;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int p = t.p0 + u * NG + grp;
    pre.off[u] = ~(size_t)0;
    if (p < t.pc1) {
      const int pi = psh >= 0 ? (p >> psh) : p / psz;
      pre.off[u] = (size_t)(a.kv.chunk(pages[pi - t.pg0], 0, 0, t.kvh) - a.kv.pool) + (size_t)(p - pi * psz) * HD + j * 8;
    }
  }
}
template <int NP>
__device__ __forceinline__ void gv_pre_load(const MegaArgs& a, int layer, GvPre<NP>& pre) {
  const bf16* kb = a.kv.pool + (size_t)layer * a.kv.layer_stride();
  const bf16* vb = kb + a.kv.kv_stride();
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    pre.k[u] = pre.v[u] = make_uint4(0u, 0u, 0u, 0u);
    if (pre.off[u] != ~(size_t)0) {
      pre.k[u] = ld_nc_v4(kb + pre.off[u]);
      pre.v[u] = ld_nc_v4(vb + pre.off[u]);
    }
  }
}
template <int NP>
__device__ __forceinline__ void gv_pre_stage(const MegaArgs& a, const GemvArgs& g, int layer, unsigned char* smem, const GvPre<NP>& pre) {
  constexpr int HD = 64 * NP, LPP = HD / 8, GPW = 32 / LPP, NG = GV_CW * GPW;
  static_assert(4 * NG * HD * 2 == GV_KVS, "staged rows");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = warp * GPW + lane / LPP, j = lane % LPP;
  const bf16* kb = a.kv.pool + (size_t)layer * a.kv.layer_stride();
  const bf16* vb = kb + a.kv.kv_stride();
  const uint32_t dk = smem_u32(smem + g.off_kv) + (uint32_t)(grp * HD + j * 8) * 2u;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    if (pre.off[u] != ~(size_t)0) {
      const uint32_t d = dk + (uint32_t)(u * NG * HD) * 2u;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(kb + pre.off[u]) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + (uint32_t)GV_KVS), "l"(vb + pre.off[u]) : "memory");
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int NP, bool KS>
__device__ __forceinline__ void gv_attention(const MegaArgs& a, const GemvArgs& g, const GvAct& A, int layer, unsigned char* smem,
                                             const MegaLayer& w, GvPre<NP>& pre, bool preloaded, const GvTask& task0) {
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
  float* pv = reinterpret_cast<float*>(sc + GA_PV);
  float* red = reinterpret_cast<float*>(smem + GO_RED);
  unsigned char* pf = smem + g.off_pf;
  const float* cos_s = reinterpret_cast<const float*>(pf + GP_COS);
  const float* sin_s = cos_s + 32 * NP;
  const bf16* qnw = reinterpret_cast<const bf16*>(pf + GP_NW);
  const bf16* knw = qnw + 64 * NP;
  const int* pages = reinterpret_cast<const int*>(pf + GP_PAGES);
  const float rs = 1.0f / sqrtf((float)HD);
  for (int task = blockIdx.x; task < ntask; task += gridDim.x) {
    // the CTA's staged task was worked out at the start of the step: a.pos / a.slot / block-table loads here would be
    // two to three dependent L2 round trips in front of everything else, once per layer
    GvTask t;
    bool any = true;
    if (preloaded) {
      t = task0;
    } else {
      any = gv_task<NP>(a, g, task, psh, t);
      t.page_new = -1;
    }
    const int b = t.b, h = t.h, kvh = t.kvh, ps = t.ps, p0 = t.p0, p1 = t.p1, pc1 = t.pc1, pg0 = t.pg0;
    float* part = A.part(layer) + (size_t)((b * a.n_q + h) * S + t.sp) * (HD + GV_PART);
    if (!any) {  // empty split (S > 1 only)
      if (threadIdx.x < HD) part[GV_PART + threadIdx.x] = 0.f;
      if (threadIdx.x == 0) {
        part[0] = -CUDART_INF_F;
        part[1] = 0.f;
      }
      continue;
    }
#ifdef GV_PROFILE_DETAIL
    const bool dbg_on = a.prof && blockIdx.x == 0 && threadIdx.x == 0;
    unsigned long long* dbg = reinterpret_cast<unsigned long long*>(smem + GO_DBG);
    long long tq = 0;
    if (dbg_on) tq = clock64();
    auto lap = [&](int k) {
      if (dbg_on) {
        const long long tt = clock64();
        dbg[10 + k] += (unsigned long long)(tt - tq);
        tq = tt;
      }
    };
#else
    auto lap = [](int) {};
#endif
    if (ntask > (int)gridDim.x) {  // several tasks per CTA: nothing was staged ahead
      gv_attn_stage<NP, KS>(a, g, layer, smem, w, t, psh, ST_STEP | ST_LAYER);
      bar_consumers();
    }
    const bool has_new = t.has_new;
    const bool writer = has_new && h == kvh * Gq;
    int page_new = t.page_new;
    if (!preloaded && writer) page_new = a.block_table[(size_t)a.slot[b] * a.max_pages + ps / psz];
    const bf16* row = A.qkv(layer) + (size_t)b * QKV;  // polled: the QKV phase of this layer stores it
    // the first 4 * NG cached positions of K and V: requested before anything that depends on this step's q
    auto load4 = [&](int base, int kv, uint4 (&r4)[4]) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = base + u * NG + grp;
        r4[u] = make_uint4(0u, 0u, 0u, 0u);
        if (p < pc1) {
          const int pi = psh >= 0 ? (p >> psh) : p / psz;
          r4[u] = ld_nc_v4(a.kv.chunk(pages[pi - pg0], layer, kv, kvh) + (size_t)(p - pi * psz) * HD + j * 8);
        }
      }
    };
    uint4 kr[4], vr[4];
    const bool staged = KS && preloaded;  // KS: a build of the kernel for short contexts (GV_STAGE_MAX_KV)
    if (!preloaded) {
      load4(p0, 0, kr);
      load4(p0, 1, vr);
    } else if (!staged) {  // long contexts: element offsets from the start of the step, the rows were requested into L2 a
                           // phase ago.  (Requested in front of the CTA barrier behind the QKV phase the values were
                           // spilled across it: ~1 us per layer; ctx 2048 1731 -> 1791 tok/s with the loads here.)
      gv_pre_load<NP>(a, layer, pre);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        kr[u] = pre.k[u];
        vr[u] = pre.v[u];
      }
    }
    lap(4);
    if (warp == 0) {
      float v[NP][2];
#pragma unroll
      for (int p = 0; p < NP; ++p) {
        const uint32_t u = gv_poll_u32(row + (size_t)h * HD + 64 * p + 2 * lane);
        v[p][0] = lo2f(u);
        v[p][1] = hi2f(u);
      }
      asm volatile("cp.async.wait_all;" ::: "memory");  // this lane's q_norm words (gv_attn_stage)
      if (w.q_norm) head_norm<NP>(v, qnw, lane);
      head_rope<NP>(v, cos_s, sin_s, lane);
      head_store<NP>(v, q_s, lane);
    } else if (warp == 1) {
      if (has_new) {
        float v[NP][2];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const uint32_t u = gv_poll_u32(row + Dq + (size_t)kvh * HD + 64 * p + 2 * lane);
          v[p][0] = lo2f(u);
          v[p][1] = hi2f(u);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");  // this lane's k_norm words (gv_attn_stage)
        if (w.k_norm) head_norm<NP>(v, knw, lane);
        head_rope<NP>(v, cos_s, sin_s, lane);
        head_store<NP>(v, knew, lane);
        if (writer) head_store<NP>(v, a.kv.chunk(page_new, layer, 0, kvh) + (size_t)(ps % psz) * HD, lane);
      }
    } else if (warp == 2) {
      if (has_new) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          const uint32_t u = gv_poll_u32(row + Dq + Dkv + (size_t)kvh * HD + 64 * p + 2 * lane);
          reinterpret_cast<uint32_t*>(vnew)[32 * p + lane] = u;
          if (writer)
            reinterpret_cast<uint32_t*>(a.kv.chunk(page_new, layer, 1, kvh) + (size_t)(ps % psz) * HD)[32 * p + lane] = u;
        }
      }
    }
    lap(5);
    bar_consumers();
    lap(0);
    // ---- scores: LPP lanes per cached position, 16 bytes of the row each
    const uint4* kv_s = reinterpret_cast<const uint4*>(smem + g.off_kv) + grp * LPP + j;  // staged rows (gv_pre_stage)
    if (staged) {
      asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
      for (int u = 0; u < 4; ++u)
        kr[u] = pre.off[u] != ~(size_t)0 ? kv_s[u * NG * LPP] : make_uint4(0u, 0u, 0u, 0u);
    }
    const uint4 qv = reinterpret_cast<const uint4*>(q_s)[j];
    float mx = -CUDART_INF_F;
    for (int base = p0; base < pc1; base += 4 * NG) {
      if (base > p0) load4(base, 0, kr);
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
    lap(1);
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
    lap(2);
    // ---- PV: the same position groups, 8 output columns per lane
    float o8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o8[k] = 0.f;
    auto fold = [&](const uint4& vv, float e) {
      o8[0] = fmaf(e, lo2f(vv.x), o8[0]);
      o8[1] = fmaf(e, hi2f(vv.x), o8[1]);
      o8[2] = fmaf(e, lo2f(vv.y), o8[2]);
      o8[3] = fmaf(e, hi2f(vv.y), o8[3]);
      o8[4] = fmaf(e, lo2f(vv.z), o8[4]);
      o8[5] = fmaf(e, hi2f(vv.z), o8[5]);
      o8[6] = fmaf(e, lo2f(vv.w), o8[6]);
      o8[7] = fmaf(e, hi2f(vv.w), o8[7]);
    };
    if (staged) {
#pragma unroll
      for (int u = 0; u < 4; ++u)
        vr[u] = pre.off[u] != ~(size_t)0 ? kv_s[GV_KVS / 16 + u * NG * LPP] : make_uint4(0u, 0u, 0u, 0u);
    }
    for (int base = p0; base < pc1; base += 4 * NG) {
      if (base > p0) load4(base, 1, vr);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = base + u * NG + grp;
        fold(vr[u], p < pc1 ? score[p - p0] : 0.f);
      }
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
      float tt = 0.f, l = 0.f;
#pragma unroll
      for (int w8 = 0; w8 < GV_CW; ++w8) {
        tt += pv[w8 * HD + threadIdx.x];
        l += red[GV_CW + w8];
      }
      if (S == 1) {
        A.att(layer)[(size_t)b * Dq + (size_t)h * HD + threadIdx.x] = f2bf(tt / l);  // self_attension.cu:137: rounded once
      } else {
        part[GV_PART + threadIdx.x] = tt;
        if (threadIdx.x == 0) {
          part[0] = M;
          part[1] = l;
        }
      }
    }
    bar_consumers();  // the scratch is reused by the next task
    lap(3);
  }
}

// attention output rows [B][Dq] bf16 (self_attension.cu:137: rounded once) from the tasks' partial results (S > 1).
// A thread owns 4 outputs of one head: per split one 8-byte load of (max, sum) and one 16-byte load of the partial
// outputs, all S of them in flight together, repeated until none shows the pattern.  Every CTA reads the same few
// lines at the same time (tasks x 272 bytes): with one 4-byte load per value -- 3 S loads per output, 148 x 5 k
// requests on ~60 lines -- the L2 slices that hold them served the requests for ~4 us per layer (ctx 300, 2 splits).
__device__ __forceinline__ void gv_load_att(const MegaArgs& a, const GemvArgs& g, unsigned char* smem, const float* part) {
  const int HD = a.hd, Dq = a.n_q * HD, S = g.n_split, stride = HD + GV_PART;
  bf16* att = reinterpret_cast<bf16*>(smem + g.off_act);
  for (int v = threadIdx.x; v < a.B * Dq / 4; v += GV_CT) {
    const int e = v * 4, b = e / Dq, r = e - b * Dq, h = r / HD, d = r - h * HD;
    const float* pp = part + (size_t)((b * a.n_q + h) * S) * stride;
    uint2 ml[GV_MAX_SPLIT];
    uint4 o[GV_MAX_SPLIT];
    {
      const long long t0 = clock64();
      bool ready;
      do {
        ready = true;
#pragma unroll
        for (int s = 0; s < GV_MAX_SPLIT; ++s) {
          ml[s] = make_uint2(0xff800000u, 0u);  // max = -inf, sum = 0
          o[s] = make_uint4(0u, 0u, 0u, 0u);
          if (s < S) {
            asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(ml[s].x), "=r"(ml[s].y) : "l"(pp + s * stride) : "memory");
            o[s] = gv_ld_relaxed_v4(pp + s * stride + GV_PART + d);
          }
        }
#pragma unroll
        for (int s = 0; s < GV_MAX_SPLIT; ++s)
          if (s < S)
            ready = ready && ml[s].x != 0xffffffffu && ml[s].y != 0xffffffffu && o[s].x != 0xffffffffu && o[s].y != 0xffffffffu &&
                    o[s].z != 0xffffffffu && o[s].w != 0xffffffffu;
        if (!ready && clock64() - t0 > GV_SPIN_LIMIT) __trap();
      } while (!ready);
    }
    float M = -CUDART_INF_F;
#pragma unroll
    for (int s = 0; s < GV_MAX_SPLIT; ++s) M = fmaxf(M, __uint_as_float(ml[s].x));
    float L = 0.f, acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s = 0; s < GV_MAX_SPLIT; ++s) {
      const float m = __uint_as_float(ml[s].x);
      const float wgt = m == -CUDART_INF_F ? 0.f : __expf(m - M);
      L += __uint_as_float(ml[s].y) * wgt;
      acc[0] += __uint_as_float(o[s].x) * wgt;
      acc[1] += __uint_as_float(o[s].y) * wgt;
      acc[2] += __uint_as_float(o[s].z) * wgt;
      acc[3] += __uint_as_float(o[s].w) * wgt;
    }
    uint2 out;
    out.x = pack2(f2bf(acc[0] / L), f2bf(acc[1] / L));
    out.y = pack2(f2bf(acc[2] / L), f2bf(acc[3] / L));
    *reinterpret_cast<uint2*>(att + e) = out;
  }
  bar_consumers();
}
__device__ __forceinline__ void gv_load_rows(unsigned char* dst, const bf16* src, int n_elems) {
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  const int n4 = n_elems >> 3;
  // polling discipline as in gv_load_norm: four requests in flight per lane, the round repeated for stragglers
  for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * GV_CT) {
    uint4 v[4];
    const long long t0 = clock64();
    bool ok;
    do {
      ok = true;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * GV_CT;
        if (i < n4) v[u] = gv_ld_relaxed_v4(s4 + i);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (i0 + u * GV_CT < n4) ok = ok && gv_word_ready(v[u].x) && gv_word_ready(v[u].y) && gv_word_ready(v[u].z) && gv_word_ready(v[u].w);
      if (!ok && clock64() - t0 > GV_SPIN_LIMIT) __trap();
    } while (!ok);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * GV_CT;
      if (i < n4) d4[i] = v[u];
    }
  }
  bar_consumers();
}

// phase stamp of a profiled launch (CTA 0, thread 0).  One out-of-line copy: the kernel's straight-line code is several
// times the 32 KB instruction cache, so everything that is not on the hot path is kept out of it.
__device__ __noinline__ void gv_stamp(const MegaArgs& a, int prof_i) {
  if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) {
    a.prof[prof_i] = globaltimer();
    if (prof_i < 8 || prof_i >= 8 + 2 * 148) a.prof[a.prof_stride + prof_i] = (unsigned long long)clock64();  // (the gap: per-CTA stamps)
  }
}

template <int NP, int NB, bool DF, bool KS>
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
  if (threadIdx.x < 16) reinterpret_cast<unsigned long long*>(smem + GO_DBG)[threadIdx.x] = 0ull;
  if (threadIdx.x < 4) reinterpret_cast<uint32_t*>(smem + GO_ZERO)[threadIdx.x] = 0u;
  gv_range_init(a, g, smem);
  __syncthreads();
  if (warp == GV_CW) {
    gv_producer(a, g, smem, smem_base, layers);
    return;
  }
  unsigned epoch = 0;
  GvRing ring{0u, 0u, 0};
  int prof_i = 0;
  auto stamp = [&]() {
    if (a.prof) gv_stamp(a, prof_i);
    ++prof_i;
  };
  const unsigned char* act = smem + g.off_act;
  GvBest best{-CUDART_INF_F, -1};
  uint32_t wr[GV_NPJ];
  gv_norm_w(layers[0].in_ln, a.H, wr);
  // this CTA's attention task (the same in every layer) when every CTA has at most one: page list + cos / sin row now
  const int psz = a.kv.page_size;
  const int psh = (psz & (psz - 1)) == 0 ? __ffs(psz) - 1 : -1;
  const int ntask = a.B * a.n_q * g.n_split;
  GvTask task0{};
  GvPre<NP> pre;
  bool my_task = ntask <= (int)gridDim.x && (int)blockIdx.x < ntask;
  if (my_task) my_task = gv_task<NP>(a, g, blockIdx.x, psh, task0);
  task0.page_new = my_task ? a.block_table[(size_t)a.slot[task0.b] * a.max_pages + task0.ps / psz] : -1;
  if (my_task) gv_attn_stage<NP, KS>(a, g, 0, smem, layers[0], task0, psh, ST_STEP | ST_LAYER);
  bar_consumers();
  if (my_task) {
    gv_pre_offsets<NP>(a, reinterpret_cast<const int*>(smem + g.off_pf + GP_PAGES), task0, psh, pre);
    if (KS) gv_pre_stage<NP>(a, g, 0, smem, pre);
    gv_attn_stage<NP, KS>(a, g, 0, smem, layers[0], task0, psh, ST_PREFETCH);
  }
  stamp();
  GPhase p;
  const GvAct A = gv_act(a, g);
  // No grid barrier inside the loop: every load of another CTA's result polls the per-layer buffer (GvAct).  The
  // bar_consumers() behind a GEMV phase only protects the shared-memory activation area / residual rows of this CTA.
  for (int l = 0; l < a.L; ++l) {
    // ---- RMSNorm + QKV
    stamp();  // (qkv.load: the row load is part of the norm here)
    gv_load_norm<NB>(a, smem, g, l == 0 ? nullptr : A.xi(l), wr);
    stamp();
    gv_phase(a, smem, layers, PH_QKV, l, p);
    gv_gemv<NB, PH_QKV>(a, g, p, smem, smem_base, ring, act, best, A.qkv(l));
    stamp();
    if (!DF) grid_arrive(a.bar, epoch);
    gv_norm_w(layers[l].post_ln, a.H, wr);
    if (!DF) grid_wait(a.bar, epoch); else bar_consumers();
    stamp();
    // ---- q/k-norm + RoPE + KV store + attention
    gv_attention<NP, KS>(a, g, A, l, smem, layers[l], pre, my_task, task0);
    stamp();
    if (!DF) grid_sync(a.bar, epoch);
    stamp();
    // ---- O + residual
    if (g.n_split == 1)
      gv_load_rows(smem + g.off_act, A.att(l), a.B * a.n_q * a.hd);
    else
      gv_load_att(a, g, smem, A.part(l));
    stamp();
    gv_phase(a, smem, layers, PH_O, l, p);
    gv_gemv<NB, PH_O>(a, g, p, smem, smem_base, ring, act, best, A.xm(l));
    stamp();
    if (!DF) grid_sync(a.bar, epoch); else bar_consumers();
    stamp();
    // ---- RMSNorm + gate/up + SiLU * up
    stamp();
    gv_load_norm<NB>(a, smem, g, A.xm(l), wr);
    stamp();
    gv_phase(a, smem, layers, PH_GATEUP, l, p);
#ifdef GV_PROFILE_DETAIL
    long long xt0 = 0;
    if (a.prof && l == 5) xt0 = (long long)globaltimer();
#endif
    gv_gemv<NB, PH_GATEUP>(a, g, p, smem, smem_base, ring, act, best, A.hh(l));
#ifdef GV_PROFILE_DETAIL  // start / end of this phase on every CTA (tools/mega_probe.py)
    if (a.prof && l == 5 && threadIdx.x == 0 && gridDim.x <= 148) {
      a.prof[a.prof_stride + 8 + blockIdx.x] = globaltimer();
      a.prof[a.prof_stride + 8 + 148 + blockIdx.x] = (unsigned long long)xt0;
    }
#endif
    stamp();
    if (!DF) grid_arrive(a.bar, epoch);
    gv_norm_w(l + 1 < a.L ? layers[l + 1].in_ln : a.final_norm, a.H, wr);
    if (!DF) grid_wait(a.bar, epoch); else bar_consumers();
    stamp();
    // ---- down + residual
    gv_load_rows(smem + g.off_act, A.hh(l), a.B * a.I);
    stamp();
    gv_phase(a, smem, layers, PH_DOWN, l, p);
    gv_gemv<NB, PH_DOWN>(a, g, p, smem, smem_base, ring, act, best, A.xi(l + 1));
    stamp();
    if (!DF) grid_arrive(a.bar, epoch);
    // the next layer's attention task: q/k-norm weights into shared memory, its cached K / V rows requested into L2
    if (my_task && l + 1 < a.L) {
      gv_attn_stage<NP, KS>(a, g, l + 1, smem, layers[l + 1], task0, psh, ST_LAYER | ST_PREFETCH);
      if (KS) gv_pre_stage<NP>(a, g, l + 1, smem, pre);
    }
    if (!DF) grid_wait(a.bar, epoch); else bar_consumers();
    stamp();
  }
  // ---- final norm + lm_head + arg-max candidates
  stamp();
  gv_load_norm<NB>(a, smem, g, A.xi(a.L), wr);
  stamp();
  gv_phase(a, smem, layers, PH_LMHEAD, 0, p);
  gv_gemv<NB, PH_LMHEAD>(a, g, p, smem, smem_base, ring, act, best, nullptr);
  stamp();
  {
    MegaCand* cs = reinterpret_cast<MegaCand*>(smem + GO_CAND);  // [GV_CW][4]
#pragma unroll
    for (int o = 16; o >= 8; o >>= 1) {  // lanes 8 q + b of a warp carry candidates of batch row b (weight row q of their items)
      const float ov = __shfl_xor_sync(0xffffffffu, best.v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best.i, o);
      if (cand_better(ov, oi, best.v, best.i)) {
        best.v = ov;
        best.i = oi;
      }
    }
    if (lane < 4) cs[warp * 4 + lane] = MegaCand{best.v, best.i};
    bar_consumers();
    if (threadIdx.x < a.B) {
      float bv = cs[threadIdx.x].val;
      int bi = cs[threadIdx.x].idx;
      for (int k = 1; k < GV_CW; ++k) {
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
  {  // every CTA is behind the grid barrier above, i.e. done with all layer buffers: put the "not stored yet" pattern back
    const size_t one = gv_act_layer_elems(a.B, a.H, a.I, a.n_q * a.hd, a.n_kv * a.hd, a.hd, g.n_split > 1 ? a.B * a.n_q * g.n_split : 0);
    const size_t n16 = ((g.dataflow ? one * (size_t)a.L : one) + (size_t)a.B * a.H + 7) >> 3;
    uint4* dst = reinterpret_cast<uint4*>(A.base);
    uint64_t keep;  // these lines are stored to and read again during the next step: keep them in L2 across the weight stream
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(keep));
    for (size_t i = (size_t)blockIdx.x * GV_CT + threadIdx.x; i < n16; i += (size_t)gridDim.x * GV_CT)
      asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1, %1, %1, %1}, %2;" ::"l"(dst + i), "r"(0xffffffffu), "l"(keep) : "memory");
  }
  if (a.prof && blockIdx.x == 0 && threadIdx.x == 0)
    for (int i = 0; i < 16; ++i) a.prof[2 * a.prof_stride + i] = reinterpret_cast<unsigned long long*>(smem + GO_DBG)[i];
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
  if (H % 128 || I % 128 || Dq % 128 || H > 8 * GV_CT * GV_NVJ) return false;
  const int nb = B <= 1 ? 1 : (B <= 2 ? 2 : 4);
  const int maxk = std::max(std::max(H, I), Dq);
  const int act_bytes = std::max(GA_END, nb * maxk * 2);
  GemvArgs g{};
  g.off_act = (GO_XS + nb * H * 2 + 127) & ~127;
  g.off_pf = (g.off_act + act_bytes + 127) & ~127;
  // short contexts (one task holds all positions, most of them staged): the task's first K / V rows in shared memory a
  // phase ahead (batch 1 / ctx 32-96: attention task 5.5 -> 3.1 us per layer, 2110 -> 2290 tok/s).  Long contexts keep
  // the 32 KB for the weight ring: the wait for the slowest attention task is then long enough to fill the ring, and
  // what it holds at that point is what the O / gate-up phases do not have to wait for (ctx 2048 staged: -3 %)
  const bool stage = max_kv_len <= GV_STAGE_MAX_KV;
  g.off_kv = stage ? (g.off_pf + GP_END + 127) & ~127 : 0;
  g.off_ring = stage ? g.off_kv + 2 * GV_KVS : (g.off_pf + GP_END + 127) & ~127;
  int rd = std::max(1, 20480 / (I * 2));
  int slot = rd * I * 2;
  slot = std::max(slot, std::max(4 * H, 2 * Dq));
  slot = (slot + 127) & ~127;
  g.slot_bytes = slot;
  g.n_slots = std::min(GV_MAX_SLOTS, (GV_SMEM_MAX - g.off_ring) / slot);
  if (g.n_slots < 3) return false;
  // KV splits: ~256 positions per task while the tasks fit one wave (kv buckets are multiples of 128: one task up to 256)
  int S = std::max(1, std::min(std::min((max_kv_len + 255) / 256, GV_MAX_SPLIT), grid / std::max(1, B * n_q)));
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
size_t decode_gemv_scratch_bytes(int H, int I, int L, int n_q, int n_kv, int hd, int num_sms) {
  // per-layer activation buffers of the data-flow synchronisation (GvAct): the largest layout over the row counts,
  // split partials for at most one task per SM.  The caller fills it with 0xFF bytes once.
  const size_t ls = gv_act_layer_elems(DECODE_GEMV_MAX_ROWS, H, I, n_q * hd, n_kv * hd, hd, num_sms);
  return (ls * (size_t)L + (size_t)DECODE_GEMV_MAX_ROWS * H + 64) * sizeof(bf16);
}

cudaError_t launch_decode_gemv(MegaArgs a, void* scratch, int num_sms, cudaStream_t st, int dataflow) {
  GvGeom gg;
  if (!gv_geometry(a.H, a.I, a.L, a.n_q, a.n_kv, a.hd, a.B, a.max_kv_len, num_sms, &gg)) return cudaErrorInvalidValue;
  gg.g.act = reinterpret_cast<bf16*>(scratch);
  {
    static const int df = [] { const char* v = getenv("QIE_GEMV_DATAFLOW"); return v ? atoi(v) : -1; }();
    gg.g.dataflow = dataflow >= 0 ? (dataflow != 0) : (df >= 0 ? (df != 0) : 1);
  }
  const int nb = a.B <= 1 ? 1 : (a.B <= 2 ? 2 : 4);
  void (*kern)(MegaArgs, GemvArgs);
  const bool ks = gg.g.off_kv != 0;
#define GV_PICK(NP, DF) \
  (ks ? (nb == 1 ? decode_gemv_kernel<NP, 1, DF, true> : (nb == 2 ? decode_gemv_kernel<NP, 2, DF, true> : decode_gemv_kernel<NP, 4, DF, true>)) \
      : (nb == 1 ? decode_gemv_kernel<NP, 1, DF, false> : (nb == 2 ? decode_gemv_kernel<NP, 2, DF, false> : decode_gemv_kernel<NP, 4, DF, false>)))
  if (gg.g.dataflow)
    kern = a.hd == 64 ? GV_PICK(1, true) : GV_PICK(2, true);
  else
    kern = a.hd == 64 ? GV_PICK(1, false) : GV_PICK(2, false);
#undef GV_PICK
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gg.smem);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(a.bar, 0, sizeof(unsigned), st);
  if (e != cudaSuccess) return e;
  void* params[] = {&a, &gg.g};
  return cudaLaunchCooperativeKernel((const void*)kern, dim3(num_sms), dim3(GV_THREADS), params, gg.smem, st);
}

}  // namespace qie
