// gemm_tcgen05.cu -- C[M,Nout] = A[M,K] * W[Nout,K]^T on the 5th-gen tensor cores
// (tcgen05.mma, accumulator in TMEM, operands staged by TMA), for M > 8: batched decode
// (M = sequences) and prefill (M = prompt tokens).  Fast numerics: fp32 accumulation in
// the tensor core's own order (+ a fixed-order split-K sum), so results match the
// reference's wmma kernel (/root/reference/layers/src/matrix_mul.cu:165-288) within the
// bf16 tolerance, not bit for bit; gemm_ref_order.cu stays the bit-exact path for M <= 8.
//
// Swap-AB: the weight tile is the UMMA "A" operand (M = 128 weight rows = 128 TMEM lanes),
// the activations are the "B" operand (N = BN tokens, 16..256), both K-major, 128-byte
// swizzle, BLOCK_K = 64 bf16.  D[w_row, token] lives in TMEM: lane = weight row, column =
// token.  Decode is HBM-bound on the weights: each CTA streams its [128 x Kslice] weight
// slab exactly once through a TMA/mbarrier ring.
//   warp 0 : TMA producer (one elected lane); weight tiles are requested BEFORE the PDL
//            wait on the previous kernel, activation tiles after it
//   warp 1 : TMEM alloc/dealloc + MMA issuer (one elected lane, tcgen05.mma / commit)
//   warps 2-5 : epilogue, TMEM -> registers (tcgen05.ld 32x32b) -> ...
// Split-K without a workspace: the S K-slices of one output tile are the S CTAs of a
// thread-block CLUSTER (grid.y = cluster dim).  Each CTA parks its fp32 partial tile in its
// own shared memory, the cluster synchronises, and CTA r sums rows [r*128/S, (r+1)*128/S)
// of all S partials through distributed shared memory (ld.shared::cluster) in rank order
// -- deterministic -- and applies the epilogue.  No partials in HBM/L2, no second launch.
// Epilogues keep the reference's rounding points (R3, R8, R9 of SURVEY.md 8a):
//   EPI_STORE     C = bf16(acc)
//   EPI_RESIDUAL  x = bf16(float(x) + float(bf16(acc)))
//   EPI_SILU_MUL  h = bf16(bf16(up) * bf16(silu(bf16(gate))));  the CTA's 128-lane tile is
//                 64 gate rows (lanes 0-63) + the same 64 up rows (lanes 64-127)
#include <cuda.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace qie {

static constexpr int BM = 128;  // weight rows per CTA (UMMA M)
static constexpr int BK = 64;   // k elements per stage (128 B rows, SWIZZLE_128B)
template <int BN>
struct TcCfg {
  static constexpr int STAGES = BN <= 128 ? 6 : 4;  // 6 x (16+16) KiB at BN=128; 4 x 48 KiB at BN=256
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug must trap (loud failure), never hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t remote;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta_rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(remote) : "memory");
  return v;
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) =
// 1024 B (8 rows x 128 B) | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16 instruction descriptor: D=F32 (bits 4-5 = 1), A=B=BF16 (bits 7-9, 10-12 = 1),
// A,B K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ float silu_ref_f(float x) {
  float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
  return __fmul_rn(x, sg);
}

// tolerance-mode SiLU for the prefill epilogue: ex2.approx + rcp.approx (relative error ~1e-6, far inside the bf16 rounding
// that follows) instead of expf + an IEEE division (~35 instructions per element on the epilogue warps that pace the
// K = 1536 gate/up tiles)
__device__ __forceinline__ float silu_fast_f(float x) {
  return __fdividef(x, 1.0f + __expf(-x));
}

struct TcSeg {
  int rows;   // weight rows of this segment
  int col0;   // first output column of this segment
  int tile0;  // first tile index
};

struct TcArgs {
  int M, K;
  int nseg, n_tiles, splits, kb_per_split, kb_total;
  int epi, dual, w_static, ld_out;
  int staged;  // splits == 1 and 8-element aligned outputs: epilogue through a bf16 tile in shared memory, 16-byte global accesses
  TcSeg seg[3];
  bf16* out;  // [M, ld_out]
};

template <int BN>
__global__ void __launch_bounds__(192) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_w0,
                                                            const __grid_constant__ CUtensorMap map_w1,
                                                            const __grid_constant__ CUtensorMap map_w2,
                                                            const __grid_constant__ CUtensorMap map_x, TcArgs g) {
  pdl_trigger();  // let the next kernel's CTAs get scheduled; they wait in their own pdl_wait()
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int A_BYTES = BM * BK * 2;  // 16 KiB
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  constexpr int TC_STAGES = TcCfg<BN>::STAGES;
  static_assert(TC_STAGES * STAGE_BYTES >= BN * BM * 4, "reduction tile must fit in the pipeline ring");
  // 1024-byte alignment for SWIZZLE_128B (same offset in every CTA of a cluster: the dynamic
  // shared window starts at the same address in all CTAs of a kernel)
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bars = smem0 + TC_STAGES * STAGE_BYTES;  // full[S], empty[S], tmem_full, slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (TC_STAGES + s); };
  const uint32_t tmem_full_bar = bars + 8u * (2 * TC_STAGES);
  const uint32_t slot = tmem_full_bar + 8u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x, split = blockIdx.y, ttile = blockIdx.z;
  int si = 0;
  while (si + 1 < g.nseg && tile >= g.seg[si + 1].tile0) ++si;
  // dual (gate/up): one tile = 64 gate rows + the same 64 up rows; segment 0 = gate map, 1 = up map
  const CUtensorMap* map_w = g.dual ? &map_w0 : (si == 0 ? &map_w0 : (si == 1 ? &map_w1 : &map_w2));
  const int row0 = g.dual ? tile * (BM / 2) : (tile - g.seg[si].tile0) * BM;
  const int seg_rows = g.seg[si].rows;
  const int col0 = g.dual ? 0 : g.seg[si].col0;
  const int kb0 = split * g.kb_per_split;
  const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
  const int nkb = max(0, kb1 - kb0);
  const int tok_base = ttile * BN;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_w) : "memory");
    if (g.dual) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));

  auto load_w = [&](int i, int s) {
    const uint32_t sa = smem0 + s * STAGE_BYTES;
    if (g.dual) {  // two 64-row boxes: gate rows -> lanes 0-63, up rows -> lanes 64-127
      tma_load_2d(sa, &map_w0, full_bar(s), (kb0 + i) * BK, row0);
      tma_load_2d(sa + A_BYTES / 2, &map_w1, full_bar(s), (kb0 + i) * BK, row0);
    } else {
      tma_load_2d(sa, map_w, full_bar(s), (kb0 + i) * BK, row0);
    }
  };

  if (warp == 0) {
    if (lane == 0) {
      // Weights never depend on the previous kernel: start streaming them BEFORE waiting for
      // it (programmatic dependent launch), so HBM stays busy across the kernel boundary.
      const int pre = g.w_static ? min(nkb, TC_STAGES) : 0;
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(full_bar(i), STAGE_BYTES);
        load_w(i, i);
      }
      pdl_wait();  // activations are produced by the previous kernel
      for (int i = 0; i < pre; ++i)
        tma_load_2d(smem0 + i * STAGE_BYTES + A_BYTES, &map_x, full_bar(i), (kb0 + i) * BK, tok_base);
      for (int i = pre; i < nkb; ++i) {
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        load_w(i, s);
        tma_load_2d(smem0 + s * STAGE_BYTES + A_BYTES, &map_x, full_bar(s), (kb0 + i) * BK, tok_base);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, BN);
      for (int i = 0; i < nkb; ++i) {
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = smem0 + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the 128-byte swizzle atom
          umma_bf16(tmem_base, make_desc(sa + k * 32), make_desc(sb + k * 32), idesc, (i | k) != 0);
        }
        umma_commit(empty_bar(s));  // frees the smem stage once the MMAs have read it
      }
      umma_commit(tmem_full_bar);  // accumulator complete
    }
  }

  // ================= epilogue =================
  if (g.staged) {
    // No split-K (prefill / large batches): lane = weight row, column = token, but the output is
    // [token][weight row] -- per-lane 2-byte global accesses would serialise (and the residual's
    // read-modify-write would expose one DRAM round trip per token).  The tile is transposed
    // through shared memory as bf16 [token][128 rows] (the pipeline ring is idle by now), then
    // written -- or read-modified-written -- with independent 16-byte accesses, a 256-byte
    // contiguous run per token.  Rounding points are the reference's (R3 / R8 / R9).
    constexpr uint32_t PITCH = BM * 2;
    const uint32_t stg = smem0;
    if (warp >= 2) {
      const int qd = warp & 3;
      pdl_wait();  // `out` may still be read (residual) by the previous kernel
      if (nkb > 0) {
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
      }
      const int lane_row = qd * 32 + lane;
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 16) {
        uint32_t r[16];
        if (nkb > 0) {
          tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + c0, r);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) r[j] = 0u;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const unsigned short hb = __bfloat16_as_ushort(f2bf(__uint_as_float(r[j])));
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(stg + (uint32_t)(c0 + j) * PITCH + (uint32_t)lane_row * 2u), "h"(hb) : "memory");
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 2, 128;" ::: "memory");  // the four epilogue warps
      const int t = threadIdx.x - 64;
      const int n_tok = min(BN, g.M - tok_base);
      auto lds128 = [](uint32_t addr) {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
        return v;
      };
      if (!g.dual) {
        const int total = n_tok * (BM / 8);
        for (int c = t; c < total; c += 4 * 128) {
          uint4 y[4], x[4];
          bf16* dst[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int cc = c + u * 128;
            const int tok = cc >> 4, part = cc & 15;
            const int w_row = row0 + part * 8;
            ok[u] = cc < total && w_row < seg_rows;
            dst[u] = g.out + (size_t)(tok_base + tok) * g.ld_out + col0 + w_row;
            if (ok[u]) {
              y[u] = lds128(stg + (uint32_t)tok * PITCH + (uint32_t)part * 16u);
              if (g.epi == EPI_RESIDUAL) x[u] = *reinterpret_cast<const uint4*>(dst[u]);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (!ok[u]) continue;
            if (g.epi == EPI_RESIDUAL) {
              const uint32_t xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w}, ys[4] = {y[u].x, y[u].y, y[u].z, y[u].w};
              uint32_t o4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                o4[e] = pack2(f2bf(__fadd_rn(lo2f(xs[e]), lo2f(ys[e]))), f2bf(__fadd_rn(hi2f(xs[e]), hi2f(ys[e]))));
              *reinterpret_cast<uint4*>(dst[u]) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            } else {
              *reinterpret_cast<uint4*>(dst[u]) = y[u];
            }
          }
        }
      } else {
        // gate rows sit in lanes 0-63, the matching up rows in lanes 64-127
        const int total = n_tok * (BM / 16);
        for (int c = t; c < total; c += 128) {
          const int tok = c >> 3, part = c & 7;
          const int w_row = row0 + part * 8;
          if (w_row >= seg_rows) continue;
          const uint4 gt = lds128(stg + (uint32_t)tok * PITCH + (uint32_t)part * 16u);
          const uint4 up = lds128(stg + (uint32_t)tok * PITCH + (uint32_t)(BM / 2) * 2u + (uint32_t)part * 16u);
          const uint32_t gs[4] = {gt.x, gt.y, gt.z, gt.w}, us[4] = {up.x, up.y, up.z, up.w};
          uint32_t o4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a0 = __fmul_rn(lo2f(us[e]), bf2f(f2bf(silu_ref_f(lo2f(gs[e])))));
            const float a1 = __fmul_rn(hi2f(us[e]), bf2f(f2bf(silu_ref_f(hi2f(gs[e])))));
            o4[e] = pack2(f2bf(a0), f2bf(a1));
          }
          *reinterpret_cast<uint4*>(g.out + (size_t)(tok_base + tok) * g.ld_out + w_row) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
    }
    __syncthreads();
    if (warp == 1) {
      tc_fence_after();
      tmem_dealloc(tmem_base, TMEM_COLS);
    }
    return;
  }
  // red[token][lane] fp32 (conflict-free: a warp writes 32 consecutive floats per token);
  // it reuses the pipeline ring, which is idle once tmem_full has fired.
  const uint32_t red = smem0;
  const bool epi_warp = warp >= 2;
  const int qd = warp & 3;  // TMEM lane quarter of this epilogue warp
  const bool via_smem = g.splits > 1 || g.dual;
  if (epi_warp) {
    pdl_wait();  // `out` may still be read (residual) by the previous kernel
    if (nkb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
    const int lane_row = qd * 32 + lane;  // TMEM lane = row inside the 128-row tile
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      if (nkb > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + c0, r);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
      if (via_smem) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(red + (uint32_t)((c0 + j) * BM + lane_row) * 4u), "r"(r[j]) : "memory");
      } else {
        // no split, single accumulator: epilogue straight from registers
        const int w_row = row0 + lane_row;
        if (w_row < seg_rows) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int tok = tok_base + c0 + j;
            if (tok < g.M) {
              bf16* dst = g.out + (size_t)tok * g.ld_out + col0 + w_row;
              const float v = __uint_as_float(r[j]);
              if (g.epi == EPI_RESIDUAL)
                *dst = f2bf(__fadd_rn(bf2f(*dst), bf2f(f2bf(v))));
              else
                *dst = f2bf(v);
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  if (via_smem) {
    // all partial tiles of the cluster are parked in shared memory
    if (g.splits > 1) cluster_sync_all();
    else __syncthreads();
    if (epi_warp) {
      const int t = threadIdx.x - 64;  // 0..127
      const int S = g.splits;
      const int tile_rows = g.dual ? BM / 2 : BM;          // output rows produced by this tile
      const int rows_per = (tile_rows + S - 1) / S;        // rows reduced by this CTA
      const int rbeg = split * rows_per;
      const int rcnt = max(0, min(rows_per, tile_rows - rbeg));
      const int n_tok = min(BN, g.M - tok_base);
      // Two outputs per thread and step, every load of the step requested before the first use: the S
      // distributed-shared-memory reads of an output (and the residual's old value) are independent, but
      // interleaved with their additions they went out one DSMEM round trip at a time.
      constexpr int MAXS = 8;  // portable cluster size
      uint32_t rb[MAXS];
#pragma unroll
      for (int sp = 0; sp < MAXS; ++sp) {
        rb[sp] = red;
        if (S > 1 && sp < S) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb[sp]) : "r"(red), "r"(sp));
      }
      const int total = rcnt * n_tok;
      for (int o0 = t; o0 < total; o0 += 2 * 128) {
        float v[2][MAXS], v2[2][MAXS];
        bf16* dst[2];
        bf16 old[2];
        bool ok[2], in_rows[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int o = o0 + u * 128;
          ok[u] = o < total;
          const int oo = ok[u] ? o : 0;
          const int tok = oo / rcnt, rr = rbeg + oo % rcnt;
          const uint32_t off1 = (uint32_t)(tok * BM + rr) * 4u;
          const uint32_t off2 = off1 + (uint32_t)(BM / 2) * 4u;  // dual: the up half of the tile
          const int w_row = row0 + rr;
          in_rows[u] = ok[u] && w_row < seg_rows;
          dst[u] = g.out + (size_t)(tok_base + tok) * g.ld_out + col0 + w_row;
          if (in_rows[u] && g.epi == EPI_RESIDUAL) old[u] = *dst[u];
#pragma unroll
          for (int sp = 0; sp < MAXS; ++sp) {
            v[u][sp] = 0.f;
            v2[u][sp] = 0.f;
            if (ok[u] && sp < S) {
              if (S > 1) {
                asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v[u][sp]) : "r"(rb[sp] + off1) : "memory");
                if (g.dual) asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v2[u][sp]) : "r"(rb[sp] + off2) : "memory");
              } else {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v[u][sp]) : "r"(red + off1) : "memory");
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v2[u][sp]) : "r"(red + off2) : "memory");
              }
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (!in_rows[u]) continue;
          float acc = 0.f, acc2 = 0.f;
#pragma unroll
          for (int sp = 0; sp < MAXS; ++sp)
            if (sp < S) {  // rank order, as before
              acc += v[u][sp];
              acc2 += v2[u][sp];
            }
          if (g.epi == EPI_STORE) {
            *dst[u] = f2bf(acc);
          } else if (g.epi == EPI_RESIDUAL) {
            *dst[u] = f2bf(__fadd_rn(bf2f(old[u]), bf2f(f2bf(acc))));
          } else {
            const float gt = bf2f(f2bf(acc)), up = bf2f(f2bf(acc2));
            *dst[u] = f2bf(__fmul_rn(up, bf2f(f2bf(silu_ref_f(gt)))));
          }
        }
      }
    }
    // nobody may exit (and release its shared memory) while a peer still reads it
    if (g.splits > 1) cluster_sync_all();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ persistent variant (prefill)
// Same tiles, same math, for launches without split-K and >= 129 tokens (BN = 256): ONE CTA per SM walks the
// output tiles (weight tile fastest, so neighbouring CTAs share a token tile in L2), the accumulator is double
// buffered in TMEM (2 x 256 columns), and the epilogue of tile j (TMEM -> bf16 staging tile in shared memory ->
// 16-byte global accesses) runs on warps 2-5 while warps 0/1 already stream and multiply tile j+1.  ncu on the
// one-tile-per-CTA kernel showed the tensor pipe 32-47 % active at K = 1536: per tile ~2.5 us of prologue and
// ~8 us of epilogue sat next to a ~7-9 us mainloop (profiles/r01c_ncu_full_prefill_kernels.csv).
static constexpr int PBN = 256;
// r02: four operand stages instead of three (ncu: tensor pipe 37-61 % active with L2 at 15-32 % of its peak -- the mainloop
// waited for TMA, 3 x 512 MMA cycles of lookahead do not cover the load latency); the room comes from a half-size
// staging tile: the epilogue drains and stores the accumulator in two halves of 128 tokens.
// (measured dead end: a 128-token tile variant -- 6 stages of 32 KiB -- for the launches whose 256-token tiles quantise badly
// over the SMs, e.g. o_proj / down_proj of the 1.5B shape, 192 tiles = 65 % of two waves: 15.5 -> 16.3 ms, 7B 62.8 -> 70.9 ms;
// the half-size tiles lose more per flop than the fuller waves give back.  Only the 0.5B shape gained, 8.5 -> 8.1 ms.)
static constexpr int P_STAGES = 4;
static constexpr int P_STAGE_BYTES = BM * BK * 2 + PBN * BK * 2;  // 48 KiB
static constexpr int P_HALF = PBN / 2;
static constexpr int P_EPI = 256;                 // epilogue threads (8 warps)
static constexpr int P_THREADS = 64 + P_EPI;      // + TMA warp + MMA warp
static constexpr int P_STAGING = P_HALF * BM * 2;                   // bf16 [128 tokens][128 rows] = 32 KiB

__global__ void __launch_bounds__(P_THREADS) gemm_tcgen05_persist_kernel(const __grid_constant__ CUtensorMap map_w0,
                                                                   const __grid_constant__ CUtensorMap map_w1,
                                                                   const __grid_constant__ CUtensorMap map_w2,
                                                                   const __grid_constant__ CUtensorMap map_x, TcArgs g,
                                                                   int token_tiles) {
  pdl_trigger();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr int A_BYTES = BM * BK * 2;
  constexpr uint32_t TMEM_COLS = 512;  // two accumulators
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t stg = smem0 + P_STAGES * P_STAGE_BYTES;
  const uint32_t bars = stg + P_STAGING;  // full[S], empty[S], tmem_full[2], tmem_empty[2], slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (P_STAGES + s); };
  auto tfull_bar = [&](int b) { return bars + 8u * (2 * P_STAGES + b); };
  auto tempty_bar = [&](int b) { return bars + 8u * (2 * P_STAGES + 2 + b); };
  const uint32_t slot = bars + 8u * (2 * P_STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = g.n_tiles * token_tiles;
  const int nkb = g.kb_total;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w2) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    for (int s = 0; s < P_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));

  // tile -> (weight tile, token tile, segment) ; dual: one tile = 64 gate rows + the same 64 up rows
  auto tile_geom = [&](int t, int& tile, int& ttile, int& si, int& row0) {
    tile = t % g.n_tiles;
    ttile = t / g.n_tiles;
    si = 0;
    while (si + 1 < g.nseg && tile >= g.seg[si + 1].tile0) ++si;
    row0 = g.dual ? tile * (BM / 2) : (tile - g.seg[si].tile0) * BM;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      pdl_wait();  // activations come from the previous kernel (these launches run for 30-300 us: no early weight prefetch)
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int tile, ttile, si, row0;
        tile_geom(t, tile, ttile, si, row0);
        const CUtensorMap* map_w = g.dual ? &map_w0 : (si == 0 ? &map_w0 : (si == 1 ? &map_w1 : &map_w2));
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % P_STAGES;
          const uint32_t ph = (it / P_STAGES) & 1;
          mbar_wait(empty_bar(s), ph ^ 1);
          mbar_expect_tx(full_bar(s), P_STAGE_BYTES);
          const uint32_t sa = smem0 + s * P_STAGE_BYTES;
          if (g.dual) {
            tma_load_2d(sa, &map_w0, full_bar(s), kb * BK, row0);
            tma_load_2d(sa + A_BYTES / 2, &map_w1, full_bar(s), kb * BK, row0);
          } else {
            tma_load_2d(sa, map_w, full_bar(s), kb * BK, row0);
          }
          tma_load_2d(sa + A_BYTES, &map_x, full_bar(s), kb * BK, ttile * PBN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BM, PBN);
      uint32_t it = 0;
      int j = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
        const int buf = j & 1;
        mbar_wait(tempty_bar(buf), (((uint32_t)j >> 1) & 1) ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t acc = tmem_base + (uint32_t)buf * PBN;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % P_STAGES;
          const uint32_t ph = (it / P_STAGES) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem0 + s * P_STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) umma_bf16(acc, make_desc(sa + k * 32), make_desc(sb + k * 32), idesc, (kb | k) != 0);
          umma_commit(empty_bar(s));
        }
        umma_commit(tfull_bar(buf));
      }
    }
  } else {
    // ---- epilogue warps 2-9: two warps per TMEM lane quarter, each draining half of the columns (r02: with four warps the
    // epilogue of a K = 1536 tile -- drain, SiLU * up, 64 KiB of stores -- took longer than the tile's mainloop)
    constexpr uint32_t PITCH = BM * 2;
    const int qd = warp & 3;
    const int lane_row = qd * 32 + lane;
    const int tq = threadIdx.x - 64;
    auto lds128 = [](uint32_t addr) {
      uint4 v;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
      return v;
    };
    pdl_wait();  // `out` may still be read (residual) by the previous kernel
    int j = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++j) {
      int tile, ttile, si, row0;
      tile_geom(t, tile, ttile, si, row0);
      const int seg_rows = g.seg[si].rows;
      const int col0 = g.dual ? 0 : g.seg[si].col0;
      const int tok_base = ttile * PBN;
      const int buf = j & 1;
      mbar_wait(tfull_bar(buf), ((uint32_t)j >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + (uint32_t)buf * PBN + ((uint32_t)(qd * 32) << 16);
      const int n_tok_all = min(PBN, g.M - tok_base);
#pragma unroll 1
      for (int hf = 0; hf < 2; ++hf) {  // tokens [128 hf, 128 hf + 128) of the tile through the 32 KiB staging tile
      const int h0 = hf * P_HALF;
      const int cw = ((warp - 2) >> 2) * (P_HALF / 2);  // this warp's columns of the half: [cw, cw + 64)
#pragma unroll 4
      for (int c0 = cw; c0 < cw + P_HALF / 2; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(acc + h0 + c0, r);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const unsigned short hb = __bfloat16_as_ushort(f2bf(__uint_as_float(r[q])));
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(stg + (uint32_t)(c0 + q) * PITCH + (uint32_t)lane_row * 2u), "h"(hb) : "memory");
        }
      }
      tc_fence_before();
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (hf == 1 && tq == 0) mbar_arrive_local(tempty_bar(buf));  // the MMA warp may overwrite this accumulator (tile j+2)
      const int n_tok = max(0, min(P_HALF, n_tok_all - h0));
      if (!g.dual) {
        const int totalc = n_tok * (BM / 8);
        for (int c = tq; c < totalc; c += 8 * P_EPI) {
          uint4 y[8], x[8];
          bf16* dst[8];
          bool ok[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int cc = c + u * P_EPI;
            const int tok = cc >> 4, part = cc & 15;
            const int w_row = row0 + part * 8;
            ok[u] = cc < totalc && w_row < seg_rows;
            dst[u] = g.out + (size_t)(tok_base + h0 + tok) * g.ld_out + col0 + w_row;
            if (ok[u]) {
              if (g.epi == EPI_RESIDUAL) x[u] = *reinterpret_cast<const uint4*>(dst[u]);
              y[u] = lds128(stg + (uint32_t)tok * PITCH + (uint32_t)part * 16u);
            }
          }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (!ok[u]) continue;
            if (g.epi == EPI_RESIDUAL) {
              const uint32_t xs[4] = {x[u].x, x[u].y, x[u].z, x[u].w}, ys[4] = {y[u].x, y[u].y, y[u].z, y[u].w};
              uint32_t o4[4];
#pragma unroll
              for (int e = 0; e < 4; ++e)
                o4[e] = pack2(f2bf(__fadd_rn(lo2f(xs[e]), lo2f(ys[e]))), f2bf(__fadd_rn(hi2f(xs[e]), hi2f(ys[e]))));
              *reinterpret_cast<uint4*>(dst[u]) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
            } else {
              *reinterpret_cast<uint4*>(dst[u]) = y[u];
            }
          }
        }
      } else {
        const int totalc = n_tok * (BM / 16);
        for (int c = tq; c < totalc; c += P_EPI) {
          const int tok = c >> 3, part = c & 7;
          const int w_row = row0 + part * 8;
          if (w_row >= seg_rows) continue;
          const uint4 gt = lds128(stg + (uint32_t)tok * PITCH + (uint32_t)part * 16u);
          const uint4 up = lds128(stg + (uint32_t)tok * PITCH + (uint32_t)(BM / 2) * 2u + (uint32_t)part * 16u);
          const uint32_t gs[4] = {gt.x, gt.y, gt.z, gt.w}, us[4] = {up.x, up.y, up.z, up.w};
          uint32_t o4[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a0 = __fmul_rn(lo2f(us[e]), bf2f(f2bf(silu_fast_f(lo2f(gs[e])))));
            const float a1 = __fmul_rn(hi2f(us[e]), bf2f(f2bf(silu_fast_f(hi2f(gs[e])))));
            o4[e] = pack2(f2bf(a0), f2bf(a1));
          }
          *reinterpret_cast<uint4*>(g.out + (size_t)(tok_base + h0 + tok) * g.ld_out + w_row) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
      if (hf == 0) asm volatile("bar.sync 2, 256;" ::: "memory");  // the staging tile is rewritten by the second half
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");  // the staging tile is rewritten by the next tile
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 row-major [rows, K] tensor, box = [box_rows, 64], 128-byte swizzle, OOB -> 0
cudaError_t make_tensor_map_2d(TensorMap2D* out, const bf16* base, int rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// Weight matrix [rows, K] viewed as a 3-D tensor {64 k-elements, rows, K/64 k-blocks} so that
// ONE TMA request moves a box of 8 rows x kc k-elements (8 x kc x 2 bytes) into shared
// memory as [k-block][row][64] with the 128-byte swizzle (decode_mega.cu weight ring).
cudaError_t make_tensor_map_w3d(TensorMap2D* out, const bf16* base, int rows, int K, int kc, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  if ((K % 64) || (kc % 64) || kc / 64 > 256) return cudaErrorInvalidValue;
  if (ld <= 0) ld = K;  // a column slice of a wider matrix (tensor parallel o_proj / down_proj) has ld > K
  cuuint64_t dims[3] = {64, (cuuint64_t)rows, (cuuint64_t)(K / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 128};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, (cuuint32_t)(kc / 64)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// KV pool [page][layer][k|v][head][slot][64] bf16 viewed as rows of 128 bytes (one head of one cached position).
// One TMA request moves box_rows consecutive slots of a (page, layer, K|V, head) chunk into shared memory; the
// 128-byte swizzle puts 16-byte chunk c of row r at chunk c ^ (r & 7), so that threads reading one row each
// (attention scores) and a warp reading one row together (PV) are both free of bank conflicts.
cudaError_t make_tensor_map_kv(TensorMap2D* out, const bf16* pool, unsigned long long rows, int hd, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return cudaErrorNotSupported;
  if ((hd != 64 && hd != 128) || box_rows < 1 || box_rows > 256 || rows == 0 || rows >= (1ull << 31)) return cudaErrorInvalidValue;
  cuuint64_t dims[2] = {(cuuint64_t)hd, (cuuint64_t)rows};  // head_dim 128: two 64-column boxes per row (the swizzle spans 128 bytes)
  cuuint64_t strides[1] = {(cuuint64_t)hd * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)pool, dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int BN>
static cudaError_t launch_bn(const TcGemm& t, const TcArgs& g, int token_tiles, cudaStream_t st) {
  constexpr int STAGE_BYTES = BM * BK * 2 + BN * BK * 2;
  constexpr size_t smem = (size_t)TcCfg<BN>::STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static PerDeviceOnce set;
  if (set.need()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
    set.done();
  }
  const CUtensorMap* w0 = reinterpret_cast<const CUtensorMap*>(t.w[0]);
  const CUtensorMap* w1 = reinterpret_cast<const CUtensorMap*>(t.w[t.nseg > 1 ? 1 : 0]);
  const CUtensorMap* w2 = reinterpret_cast<const CUtensorMap*>(t.w[t.nseg > 2 ? 2 : 0]);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(g.n_tiles, g.splits, token_tiles);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (g.splits > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 1;
    at[na].val.clusterDim.y = g.splits;
    at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (g_use_pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN>, *w0, *w1, *w2, *reinterpret_cast<const CUtensorMap*>(t.x), g);
}

int tc_token_tile(int M) { return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256; }

cudaError_t launch_gemm_tcgen05(const TcGemm& t, int num_sms, cudaStream_t st, int* launches) {
  if (t.M < 1 || t.K < BK || (t.K % 8) || t.nseg < 1 || t.nseg > 3) return cudaErrorInvalidValue;
  TcArgs g{};
  g.M = t.M;
  g.K = t.K;
  g.nseg = t.nseg;
  g.epi = t.epi;
  g.dual = t.epi == EPI_SILU_MUL;
  g.w_static = t.w_static;
  g.ld_out = t.ld_out;
  g.out = t.out;
  int tiles = 0, cols = 0;
  if (g.dual) {
    // segments = (gate, up) with equal row counts; the weight maps must have 64-row boxes
    if (t.nseg != 2 || t.rows[0] != t.rows[1]) return cudaErrorInvalidValue;
    g.seg[0].rows = t.rows[0];
    g.seg[0].col0 = 0;
    g.seg[0].tile0 = 0;
    g.nseg = 1;
    tiles = (t.rows[0] + BM / 2 - 1) / (BM / 2);
    cols = t.rows[0];
  } else {
    for (int i = 0; i < t.nseg; ++i) {
      g.seg[i].rows = t.rows[i];
      g.seg[i].col0 = cols;
      g.seg[i].tile0 = tiles;
      tiles += (t.rows[i] + BM - 1) / BM;
      cols += t.rows[i];
    }
  }
  (void)cols;
  g.n_tiles = tiles;
  g.kb_total = (t.K + BK - 1) / BK;
  const int BN = tc_token_tile(t.M);
  const int token_tiles = (t.M + BN - 1) / BN;
  // split-K (= cluster size along grid.y) so that roughly one CTA lands on every SM; every
  // slice keeps >= 2 k-blocks; cluster size <= 8 (the portable maximum)
  int splits = 1;
  const int ctas = tiles * token_tiles;
  if (ctas * 2 <= num_sms) {
    splits = num_sms / ctas;
    splits = std::min(splits, std::max(1, g.kb_total / 2));
    splits = std::min(splits, std::min(8, std::max(1, t.max_splits)));  // portable cluster size
  }
  g.kb_per_split = (g.kb_total + splits - 1) / splits;
  splits = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;
  g.splits = splits;
  {
    bool vec = (t.ld_out % 8 == 0) && ((reinterpret_cast<uintptr_t>(t.out) & 15) == 0);
    for (int i = 0; i < t.nseg; ++i) vec = vec && (t.rows[i] % 8 == 0);
    g.staged = splits == 1 && vec;
  }
  cudaError_t e;
  static const bool persist_on = [] { const char* v = getenv("QIE_TC_PERSIST"); return !(v && v[0] == '0'); }();
  if (persist_on && g.staged && BN == PBN && tiles * token_tiles > num_sms / 2) {
    // prefill-sized launches: persistent tile loop, epilogue overlapped with the next tile's mainloop
    constexpr size_t smem = (size_t)P_STAGES * P_STAGE_BYTES + P_STAGING + 1024 /*align*/ + 256 /*barriers*/;
    static PerDeviceOnce set;
    if (set.need()) {
      e = cudaFuncSetAttribute(gemm_tcgen05_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      set.done();
    }
    const CUtensorMap* w0 = reinterpret_cast<const CUtensorMap*>(t.w[0]);
    const CUtensorMap* w1 = reinterpret_cast<const CUtensorMap*>(t.w[t.nseg > 1 ? 1 : 0]);
    const CUtensorMap* w2 = reinterpret_cast<const CUtensorMap*>(t.w[t.nseg > 2 ? 2 : 0]);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(std::min(num_sms, tiles * token_tiles));
    cfg.blockDim = dim3(P_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    if (launches) *launches = 1;
    return cudaLaunchKernelEx(&cfg, gemm_tcgen05_persist_kernel, *w0, *w1, *w2, *reinterpret_cast<const CUtensorMap*>(t.x), g, token_tiles);
  }
  switch (BN) {
    case 16: e = launch_bn<16>(t, g, token_tiles, st); break;
    case 32: e = launch_bn<32>(t, g, token_tiles, st); break;
    case 64: e = launch_bn<64>(t, g, token_tiles, st); break;
    case 128: e = launch_bn<128>(t, g, token_tiles, st); break;
    default: e = launch_bn<256>(t, g, token_tiles, st); break;
  }
  if (launches) *launches = 1;
  return e;
}

}  // namespace qie
