// gemm_ref_order.cu -- skinny GEMM  C[M,Nout] = A[M,K] * W[Nout,K]^T  for decode / short
// prefill, reproducing the accumulation order of the reference's wmma kernel
// (/root/reference/layers/src/matrix_mul.cu:165-288):
//   * activations are the A operand (row-major 16 x 16 tiles, rows >= M are zero),
//     weight rows are the B operand ("col" = k-contiguous), exactly as the reference
//     feeds nvcuda::wmma m16n16k16 (:199-201) which lowers to two HMMA.16816.F32.BF16;
//   * the inner dimension is walked in 16-wide chunks in ascending order and every
//     chunk is accumulated into the same fp32 fragment (:206-259) -- no split-K;
//   * the fp32 accumulator is rounded to bf16 once (:272).
// Unlike the reference (scalar 2-byte loads, one warp per tile, W re-read per 16 rows of
// A) the weights are streamed exactly once with 16-byte cp.async through a multi-stage
// shared-memory ring, XOR-swizzled so ldmatrix is bank-conflict free, and the epilogue
// fuses what the reference does in separate kernels + HBM round trips:
//   EPI_STORE     C = bf16(acc)
//   EPI_RESIDUAL  x = bf16(float(x) + float(bf16(acc)))          (residual_add.cu:7-18)
//   EPI_SILU_MUL  h = bf16(bf16(up) * bf16(silu(bf16(gate))))    (SiLU.cu:10-23, element_add.cu:4-13)
// The rounding points are the reference's (SURVEY.md 8a list R3, R8, R9).
#include "common.cuh"
#include "kernels.h"
#include "launch.h"
#include "ref_math.cuh"

namespace qie {

static constexpr int KS = 64;  // k elements per pipeline stage (128 B per row)

__device__ __forceinline__ void cp_async16_zfill(uint32_t dst, const void* src, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}


// cp.async.wait_group needs an immediate; D is a launch parameter, so dispatch.
__device__ __forceinline__ void cp_async_wait_dyn(int n) {
  switch (n) {
#define QIE_W(i) case i: asm volatile("cp.async.wait_group " #i ";\n" ::: "memory"); break;
    QIE_W(0) QIE_W(1) QIE_W(2) QIE_W(3) QIE_W(4) QIE_W(5) QIE_W(6) QIE_W(7) QIE_W(8) QIE_W(9) QIE_W(10)
    QIE_W(11) QIE_W(12) QIE_W(13) QIE_W(14) QIE_W(15) QIE_W(16) QIE_W(17) QIE_W(18) QIE_W(19) QIE_W(20)
    QIE_W(21) QIE_W(22) QIE_W(23) QIE_W(24) QIE_W(25) QIE_W(26) QIE_W(27) QIE_W(28) QIE_W(29) QIE_W(30)
#undef QIE_W
    default: asm volatile("cp.async.wait_group 0;\n" ::: "memory"); break;
  }
}

// One warp = one unit = 8 consecutive weight rows over the full K.
// Block = NW warps; all warps share the A tile of each stage.
// MT = number of 16-row token tiles (M <= 16*MT).  MT == 0 is the decode case M <= 8:
// the A fragments are read with predicated 32-bit LDS (token rows >= M are register
// zeros), so no zero rows have to exist in shared memory.
template <int MT>
__global__ void __launch_bounds__(256) gemm_ref_order_kernel(GemmArgs g) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(128) unsigned char smem[];
  const int nw = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = g.stages;
  constexpr int MTT = MT == 0 ? 1 : MT;       // accumulator tiles
  constexpr int A_ROWS = MT == 0 ? 8 : MT * 16;
  const int a_bytes = A_ROWS * 128;
  const int w_tiles = g.dual ? 2 : 1;
  const int stage_bytes = a_bytes + nw * w_tiles * 1024;
  const uint32_t smem_base = smem_u32(smem);

  // unit -> segment
  const int unit = blockIdx.x * nw + warp;
  int seg = 0;
  bool active = unit < g.total_units;
  if (active) {
    while (seg + 1 < g.nseg && unit >= g.unit_begin[seg + 1]) ++seg;
  }
  const GemmSeg& S = g.seg[seg];
  const int row0 = active ? (unit - g.unit_begin[seg]) * 8 : 0;
  const int K = g.K;
  const int nk = (K + KS - 1) / KS;

  // rows >= M of every A tile are read by ldmatrix (MT >= 1) and must be zero; they are
  // never written by cp.async, so zero them once.
  if (MT != 0) {
    const int zr = A_ROWS - g.M;  // rows to clear per stage
    for (int i = threadIdx.x; i < D * zr * 8; i += blockDim.x) {
      int st = i / (zr * 8), rem = i - st * zr * 8;
      int r = g.M + (rem >> 3), c = rem & 7;
      *reinterpret_cast<uint4*>(smem + st * stage_bytes + r * 128 + (c << 4)) = make_uint4(0, 0, 0, 0);
    }
    __syncthreads();
  }

  auto load_stage = [&](int ks, int buf) {
    const uint32_t sb = smem_base + buf * stage_bytes;
    const int k0 = ks * KS;
    // A tile: M rows x 8 chunks, spread over the block
    for (int c = threadIdx.x; c < g.M * 8; c += blockDim.x) {
      int r = c >> 3, col = c & 7;
      int k = k0 + col * 8;
      bool ok = k < K;
      const bf16* src = g.A + (size_t)r * g.lda + (ok ? k : 0);
      cp_async16_zfill(sb + r * 128 + ((col ^ (r & 7)) << 4), src, ok);
    }
    if (active) {
      const uint32_t wb = sb + a_bytes + warp * w_tiles * 1024;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        int c = lane + 32 * i;
        int r = c >> 3, col = c & 7;
        int k = k0 + col * 8;
        bool ok = k < K;
        int gr = row0 + r;
        if (gr >= S.rows) gr = S.rows - 1;  // clamp: duplicates are never stored
        const size_t ldw = S.ldw ? (size_t)S.ldw : (size_t)K;  // weight row stride (tensor-parallel column slices)
        cp_async16_zfill(wb + r * 128 + ((col ^ r) << 4), S.W + (size_t)gr * ldw + (ok ? k : 0), ok);
        if (g.dual)
          cp_async16_zfill(wb + 1024 + r * 128 + ((col ^ r) << 4), S.W2 + (size_t)gr * ldw + (ok ? k : 0), ok);
      }
    }
  };

  float acc[MTT][4], acc2[MTT][4];
#pragma unroll
  for (int m = 0; m < MTT; ++m)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[m][i] = acc2[m][i] = 0.f;

  // prologue
  for (int s = 0; s < D - 1; ++s) {
    if (s < nk) load_stage(s, s);
    cp_async_commit();
  }

  for (int ks = 0; ks < nk; ++ks) {
    // stage ks is complete once at most D-2 younger groups are pending
    cp_async_wait_dyn(D - 2);
    __syncthreads();
    {
      int nxt = ks + D - 1;
      if (nxt < nk) load_stage(nxt, nxt % D);
      cp_async_commit();
    }
    if (active) {
      const uint32_t sb = smem_base + (ks % D) * stage_bytes;
      const uint32_t wb = sb + a_bytes + warp * w_tiles * 1024;
#pragma unroll
      for (int half = 0; half < 2; ++half) {  // two 16-wide k chunks per ldmatrix.x4 of W
        uint32_t b[4], b2[4];
        {
          int r = lane & 7, ch = half * 4 + (lane >> 3);
          ldmatrix_x4(b[0], b[1], b[2], b[3], wb + r * 128 + ((ch ^ r) << 4));
          if (g.dual) ldmatrix_x4(b2[0], b2[1], b2[2], b2[3], wb + 1024 + r * 128 + ((ch ^ r) << 4));
        }
#pragma unroll
        for (int kc = 0; kc < 2; ++kc) {
          int chunk0 = half * 4 + kc * 2;
#pragma unroll
          for (int m = 0; m < MTT; ++m) {
            uint32_t a[4];
            if (MT == 0) {
              // a0: (token lane/4, k = (lane%4)*2..+1) ; a2: same token, k + 8 ; tokens 8..15 = 0
              int tr = lane >> 2;
              uint32_t off = sb + tr * 128 + (lane & 3) * 4;
              a[1] = 0u;
              a[3] = 0u;
              if (tr < g.M) {
                asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(a[0]) : "r"(off + (((chunk0) ^ tr) << 4)));
                asm volatile("ld.shared.b32 %0, [%1];\n" : "=r"(a[2]) : "r"(off + (((chunk0 + 1) ^ tr) << 4)));
              } else {
                a[0] = 0u;
                a[2] = 0u;
              }
            } else {
              int r = m * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
              int ch = chunk0 + (lane >> 4);
              ldmatrix_x4(a[0], a[1], a[2], a[3], sb + r * 128 + ((ch ^ (r & 7)) << 4));
            }
            mma_bf16_16816(acc[m], a, b[kc * 2], b[kc * 2 + 1]);
            if (g.dual) mma_bf16_16816(acc2[m], a, b2[kc * 2], b2[kc * 2 + 1]);
          }
        }
      }
    }
  }

  if (!active) return;
  // epilogue: c0,c1 -> (token = lane/4, n = (lane%4)*2 + {0,1}); c2,c3 -> token + 8
  const int n0 = row0 + (lane & 3) * 2;
#pragma unroll
  for (int m = 0; m < MTT; ++m)
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      int tok = m * 16 + (lane >> 2) + hrow * 8;
      if (tok >= g.M) continue;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        int n = n0 + j;
        if (n >= S.rows) continue;
        float v = acc[m][hrow * 2 + j];
        bf16* dst = S.out + (size_t)tok * S.ld_out + n;
        if (g.epi == EPI_STORE) {
          *dst = f2bf(v);
        } else if (g.epi == EPI_STORE_F32) {
          reinterpret_cast<float*>(S.out)[(size_t)tok * S.ld_out + n] = v;
        } else if (g.epi == EPI_RESIDUAL) {
          float y = bf2f(f2bf(v));
          *dst = f2bf(__fadd_rn(bf2f(*dst), y));
        } else {  // EPI_SILU_MUL : acc = gate, acc2 = up
          float gt = bf2f(f2bf(v));
          float up = bf2f(f2bf(acc2[m][hrow * 2 + j]));
          float gs = bf2f(f2bf(silu_ref(gt)));
          *dst = f2bf(__fmul_rn(up, gs));
        }
      }
    }
}


template <int MT>
static cudaError_t launch_mt(const GemmArgs& g, int nw, size_t smem, cudaStream_t st) {
  static PerDeviceOnce attr_set;
  if (attr_set.need()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ref_order_kernel<MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024);
    if (e != cudaSuccess) return e;
    attr_set.done();
  }
  int grid = (g.total_units + nw - 1) / nw;
  (void)launch_k(gemm_ref_order_kernel<MT>, dim3(grid), dim3(nw * 32), smem, st, g);
  return cudaGetLastError();
}

cudaError_t launch_gemm_ref_order(const GemmArgs& a, int num_sms, cudaStream_t st) {
  GemmArgs g = a;
  if (g.M < 1 || g.M > 64 || g.K < 8 || (g.K & 7) || g.nseg < 1 || g.nseg > 3) return cudaErrorInvalidValue;
  int u = 0;
  for (int i = 0; i < g.nseg; ++i) {
    g.unit_begin[i] = u;
    u += (g.seg[i].rows + 7) / 8;
  }
  g.unit_begin[g.nseg] = u;
  g.total_units = u;
  g.dual = g.epi == EPI_SILU_MUL;
  const int mt = g.M <= 8 ? 0 : (g.M <= 16 ? 1 : (g.M <= 32 ? 2 : 4));
  // warps (= 8-row units) per block: aim at ~2 blocks per SM so the hardware scheduler
  // balances; small outputs (896 rows = 112 units) get one warp per block = one per SM.
  int per = mt == 0 ? u / (2 * num_sms) : u / num_sms;
  int nw = per >= 8 ? 8 : per >= 4 ? 4 : per >= 2 ? 2 : 1;
  const int a_bytes = (mt == 0 ? 8 : mt * 16) * 128;
  const int stage_bytes = a_bytes + nw * (g.dual ? 2 : 1) * 1024;
  const int nk = (g.K + KS - 1) / KS;
  int budget = nw == 1 ? 72 * 1024 : 64 * 1024;
  if (mt >= 2) budget = 96 * 1024;
  int D = budget / stage_bytes;
  if (D > 32) D = 32;
  if (D > nk + 1) D = nk + 1;
  if (D < 2) D = 2;
  g.stages = D;
  size_t smem = (size_t)D * stage_bytes;
  switch (mt) {
    case 0: return launch_mt<0>(g, nw, smem, st);
    case 1: return launch_mt<1>(g, nw, smem, st);
    case 2: return launch_mt<2>(g, nw, smem, st);
    default: return launch_mt<4>(g, nw, smem, st);
  }
}

}  // namespace qie
