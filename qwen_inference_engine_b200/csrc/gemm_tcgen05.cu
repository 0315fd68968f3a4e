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
// token.  Decode is HBM-bound on the weights, so the grid is (weight tiles x split-K) to
// put >= one CTA on every SM; each CTA streams its [128 x Kslice] weight slab exactly once
// through a 6-stage TMA/mbarrier ring.
//   warp 0 : TMA producer (one elected lane)
//   warp 1 : TMEM alloc/dealloc + MMA issuer (one elected lane, tcgen05.mma / commit)
//   warps 2-5 : epilogue, TMEM -> registers (tcgen05.ld 32x32b) -> global
// Split-K partials go to an fp32 workspace [split][token][Nout]; gemm_finalize_kernel sums
// them in split order and applies the epilogue (store / residual / SiLU*up) with the
// reference's rounding points (R3, R8, R9 of SURVEY.md 8a).
#include <cuda.h>

#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace qie {

static constexpr int BM = 128;     // weight rows per CTA (UMMA M)
static constexpr int BK = 64;      // k elements per stage (128 B rows, SWIZZLE_128B)
template <int BN>
struct TcCfg {
  static constexpr int STAGES = BN <= 128 ? 6 : 4;  // 6 x (16+16) KiB = 192 KiB at BN=128; 4 x 48 KiB at BN=256
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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

struct TcSeg {
  int rows;      // weight rows of this segment
  int col0;      // first output column of this segment in the [*, ld] output / workspace
  int tile0;     // first tile index
};

struct TcArgs {
  int M, K, ld;        // tokens, inner dim, total output columns (sum of segment rows)
  int nseg, n_tiles, splits, kb_per_split, kb_total;
  TcSeg seg[3];
  float* ws;           // [splits][M][ld] fp32 partials (splits > 1 or epilogue needs finalize)
  bf16* direct_out;    // splits == 1 && plain store: bf16 [M][ld]
  // fused split-K reduction: the LAST CTA to finish a tile group sums the partials in split
  // order and applies the epilogue (no separate finalize launch)
  int* counters;       // [token_tiles * groups], zero-initialised once, self-resetting (atomicInc wraps)
  int fuse, groups, expected, epi, n_out, ld_out, w_static;
  bf16* out;
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
  // 1024-byte alignment for SWIZZLE_128B
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
  const CUtensorMap* map_w = si == 0 ? &map_w0 : (si == 1 ? &map_w1 : &map_w2);
  const int row0 = (tile - g.seg[si].tile0) * BM;  // first weight row of this tile inside its segment
  const int kb0 = split * g.kb_per_split;
  const int kb1 = min(g.kb_total, kb0 + g.kb_per_split);
  const int nkb = max(0, kb1 - kb0);
  const int tok_base = ttile * BN;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_w) : "memory");
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

  if (warp == 0) {
    if (lane == 0) {
      // Weights never depend on the previous kernel: start streaming them BEFORE waiting for
      // it (programmatic dependent launch), so HBM stays busy across the kernel boundary.
      const int pre = g.w_static ? min(nkb, TC_STAGES) : 0;
      for (int i = 0; i < pre; ++i) {
        mbar_expect_tx(full_bar(i), STAGE_BYTES);
        tma_load_2d(smem0 + i * STAGE_BYTES, map_w, full_bar(i), (kb0 + i) * BK, row0);
      }
      pdl_wait();  // activations are produced by the previous kernel
      for (int i = 0; i < pre; ++i)
        tma_load_2d(smem0 + i * STAGE_BYTES + A_BYTES, &map_x, full_bar(i), (kb0 + i) * BK, tok_base);
      for (int i = pre; i < nkb; ++i) {
        const int s = i % TC_STAGES;
        const uint32_t ph = (i / TC_STAGES) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        mbar_expect_tx(full_bar(s), STAGE_BYTES);
        const uint32_t sa = smem0 + s * STAGE_BYTES, sb = sa + A_BYTES;
        tma_load_2d(sa, map_w, full_bar(s), (kb0 + i) * BK, row0);
        tma_load_2d(sb, &map_x, full_bar(s), (kb0 + i) * BK, tok_base);
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
      umma_commit(tmem_full_bar);   // accumulator complete
    }
  } else {
    // epilogue warps 2..5 -> TMEM lane quarter (warp % 4)
    pdl_wait();  // ws / out may still be read by the previous kernel
    const int qd = warp & 3;
    const int w_row = row0 + qd * 32 + lane;  // weight row inside the segment
    const bool row_ok = w_row < g.seg[si].rows;
    const int col = g.seg[si].col0 + w_row;
    if (nkb > 0) {
      mbar_wait(tmem_full_bar, 0);
      tc_fence_after();
    }
#pragma unroll
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t r[16];
      if (nkb > 0) {
        tmem_ld16(tmem_base + ((uint32_t)(qd * 32) << 16) + c0, r);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) r[j] = 0u;
      }
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int tok = tok_base + c0 + j;
          if (tok < g.M) {
            if (g.direct_out)
              g.direct_out[(size_t)tok * g.ld + col] = f2bf(__uint_as_float(r[j]));
            else
              g.ws[((size_t)split * g.M + tok) * g.ld + col] = __uint_as_float(r[j]);
          }
        }
      }
    }
    if (g.fuse) {
      // ---- fused split-K reduction + epilogue by the last CTA of this tile group
      __shared__ int s_last;
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int dual = g.epi == EPI_SILU_MUL;
      const int grp = ttile * g.groups + (dual ? tile % g.groups : tile);
      if (threadIdx.x == 64) {
        const unsigned old = atomicInc(reinterpret_cast<unsigned*>(g.counters) + grp, (unsigned)g.expected - 1u);
        s_last = old == (unsigned)g.expected - 1u;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (s_last) {
        __threadfence();
        // output column handled by this thread (dual: gate column c, up column n_out + c)
        const int seg_row = row0 + qd * 32 + lane;
        const int c = dual ? seg_row : col;
        const bool ok = seg_row < g.seg[si].rows;
        const int t_end = min(g.M, tok_base + BN);
        if (ok) {
          for (int tok = tok_base; tok < t_end; ++tok) {
            float acc = 0.f, acc2 = 0.f;
            for (int sp = 0; sp < g.splits; ++sp) {
              const float* pw = g.ws + ((size_t)sp * g.M + tok) * g.ld;
              acc += __ldcg(pw + c);
              if (dual) acc2 += __ldcg(pw + g.n_out + c);
            }
            bf16* dst = g.out + (size_t)tok * g.ld_out + c;
            if (g.epi == EPI_STORE) {
              *dst = f2bf(acc);
            } else if (g.epi == EPI_RESIDUAL) {
              *dst = f2bf(__fadd_rn(bf2f(*dst), bf2f(f2bf(acc))));
            } else {
              const float gt = bf2f(f2bf(acc)), up = bf2f(f2bf(acc2));
              *dst = f2bf(__fmul_rn(up, bf2f(f2bf(silu_ref_f(gt)))));
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------ split-K sum + epilogue
// ws: [splits][M][ld]; EPI_SILU_MUL: columns [0,n) = gate, [n,2n) = up, out has n columns
__global__ void gemm_finalize_kernel(const float* __restrict__ ws, bf16* __restrict__ out, int M, int ld, int n_out,
                                     int ld_out, int splits, int epi) {
  pdl_wait();
  pdl_trigger();
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * n_out) return;
  const int tok = (int)(i / n_out), c = (int)(i % n_out);
  float acc = 0.f, acc2 = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float* p = ws + ((size_t)s * M + tok) * ld;
    acc += p[c];
    if (epi == EPI_SILU_MUL) acc2 += p[n_out + c];
  }
  bf16* dst = out + (size_t)tok * ld_out + c;
  if (epi == EPI_STORE) {
    *dst = f2bf(acc);
  } else if (epi == EPI_RESIDUAL) {
    *dst = f2bf(__fadd_rn(bf2f(*dst), bf2f(f2bf(acc))));
  } else {
    const float gt = bf2f(f2bf(acc)), up = bf2f(f2bf(acc2));
    *dst = f2bf(__fmul_rn(up, bf2f(f2bf(silu_ref_f(gt)))));
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

template <int BN>
static cudaError_t launch_bn(const TcGemm& t, const TcArgs& g, int token_tiles, cudaStream_t st) {
  constexpr int STAGE_BYTES = BM * BK * 2 + BN * BK * 2;
  constexpr size_t smem = (size_t)TcCfg<BN>::STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tcgen05_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    set = true;
  }
  dim3 grid(g.n_tiles, g.splits, token_tiles);
  const CUtensorMap* w0 = reinterpret_cast<const CUtensorMap*>(t.w[0]);
  const CUtensorMap* w1 = reinterpret_cast<const CUtensorMap*>(t.w[t.nseg > 1 ? 1 : 0]);
  const CUtensorMap* w2 = reinterpret_cast<const CUtensorMap*>(t.w[t.nseg > 2 ? 2 : 0]);
  (void)launch_k(gemm_tcgen05_kernel<BN>, dim3(grid), dim3(192), smem, st, *w0, *w1, *w2, *reinterpret_cast<const CUtensorMap*>(t.x), g);
  return cudaGetLastError();
}

int tc_token_tile(int M) { return M <= 16 ? 16 : M <= 32 ? 32 : M <= 64 ? 64 : M <= 128 ? 128 : 256; }

cudaError_t launch_gemm_tcgen05(const TcGemm& t, int num_sms, cudaStream_t st, int* launches) {
  if (t.M < 1 || t.K < BK || (t.K % 8) || t.nseg < 1 || t.nseg > 3) return cudaErrorInvalidValue;
  TcArgs g{};
  g.M = t.M;
  g.K = t.K;
  g.nseg = t.nseg;
  int tiles = 0, cols = 0;
  for (int i = 0; i < t.nseg; ++i) {
    g.seg[i].rows = t.rows[i];
    g.seg[i].col0 = cols;
    g.seg[i].tile0 = tiles;
    tiles += (t.rows[i] + BM - 1) / BM;
    cols += t.rows[i];
  }
  g.n_tiles = tiles;
  g.ld = cols;
  g.kb_total = (t.K + BK - 1) / BK;
  const int BN = tc_token_tile(t.M);
  const int token_tiles = (t.M + BN - 1) / BN;
  // split-K so that >= ~1 CTA lands on every SM; every split keeps >= 2 k-blocks
  int splits = 1;
  if (tiles * token_tiles < num_sms) {
    splits = (num_sms + tiles * token_tiles - 1) / (tiles * token_tiles);
    splits = std::min(splits, std::max(1, g.kb_total / 2));
    splits = std::min(splits, t.max_splits);
    splits = std::min(splits, 8);  // the last CTA of a tile sums the partials: keep that tail short
  }
  g.kb_per_split = (g.kb_total + splits - 1) / splits;
  splits = (g.kb_total + g.kb_per_split - 1) / g.kb_per_split;
  g.splits = splits;
  const bool direct = splits == 1 && t.epi == EPI_STORE;
  g.direct_out = direct ? t.out : nullptr;
  g.ws = t.ws;
  const bool dual = t.epi == EPI_SILU_MUL;
  const int n_out_cols = dual ? cols / 2 : cols;
  // fused finalize needs whole tiles per group (dual: gate tile i pairs with up tile i)
  const bool can_fuse = !direct && t.counters && (!dual || (t.nseg == 2 && t.rows[0] == t.rows[1] && t.rows[0] % BM == 0));
  g.fuse = can_fuse ? 1 : 0;
  g.groups = dual ? tiles / 2 : tiles;
  g.expected = dual ? 2 * splits : splits;
  g.counters = t.counters;
  g.epi = t.epi;
  g.n_out = n_out_cols;
  g.ld_out = t.ld_out;
  g.out = t.out;
  g.w_static = t.w_static;
  if (can_fuse && (size_t)g.groups * token_tiles > (size_t)t.n_counters) return cudaErrorInvalidValue;
  if (!direct && (size_t)splits * t.M * cols * sizeof(float) > t.ws_bytes) return cudaErrorMemoryAllocation;
  cudaError_t e;
  switch (BN) {
    case 16: e = launch_bn<16>(t, g, token_tiles, st); break;
    case 32: e = launch_bn<32>(t, g, token_tiles, st); break;
    case 64: e = launch_bn<64>(t, g, token_tiles, st); break;
    case 128: e = launch_bn<128>(t, g, token_tiles, st); break;
    default: e = launch_bn<256>(t, g, token_tiles, st); break;
  }
  if (e != cudaSuccess) return e;
  if (launches) *launches = 1;
  if (!direct && !can_fuse) {
    const int n_out = t.epi == EPI_SILU_MUL ? cols / 2 : cols;
    const size_t n = (size_t)t.M * n_out;
    (void)launch_k(gemm_finalize_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, t.ws, t.out, t.M, cols, n_out, t.ld_out, splits,
                                                                      t.epi);
    e = cudaGetLastError();
    if (launches) *launches = 2;
  }
  return e;
}

}  // namespace qie
