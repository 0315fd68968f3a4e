// attn_decode_fast.cu -- split-KV flash-decoding over the paged KV pool (fast numerics).
//
// Replaces selfattention (/root/reference/layers/src/self_attension.cu:10-149) for decode
// rows when the context is long or the batch is large.  The reference walks a linked list
// of pages for every kv position and runs the softmax in one thread; its arithmetic order
// (sequential sum over positions) cannot be parallelised, so this kernel trades bit
// equality for bandwidth: results agree with the reference within the north-star 1e-2
// bf16 tolerance (tests/test_gpu_fast_path.py), the reference-order kernel in
// ops_ref_order.cu stays the bit-exact path.
//
// Mapping (HBM-bound: K and V are each read exactly once per step):
//   grid  = (n_splits, n_kv_heads, batch)    one CTA per (sequence, kv head, kv range)
//   block = 4 warps; every warp owns 16 of the 64 positions of a tile
//   GQA   : the G = n_q/n_kv query heads that share a kv head are the M rows (padded to
//           16) of one m16n8k16 bf16 MMA, so K/V are loaded once for all of them
//   KV    : page chunks [slot][hd] are contiguous in pool[page][layer][k|v][head][slot][hd];
//           16-byte cp.async into a 3-stage XOR-swizzled ring, ldmatrix (K) /
//           ldmatrix.trans (V) conflict-free
//   softmax: online (running max / sum in fp32, exp2f with the scale folded in), P split into
//           bf16 hi + lo parts for the PV MMA (two MMAs, 16 mantissa bits)
//   splits: partial (max, sum, unnormalised O) per split -> attn_combine_kernel
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace qie {

static constexpr int TILE = 64;    // kv positions per pipeline stage
#ifndef QIE_FD_STAGES
#define QIE_FD_STAGES 3
#endif
static constexpr int STAGES = QIE_FD_STAGES;   // 3 x 16 KiB (hd 64): 4 CTAs / SM

template <int HD>
struct FastAttnSmem {
  static constexpr int ROW_BYTES = HD * 2;
  static constexpr int TILE_BYTES = TILE * ROW_BYTES;       // one of K or V
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES;
  static constexpr int RING = STAGES * STAGE_BYTES;
  static constexpr int MAX_PAGES = 512;                     // page ids of this CTA's kv range, cached
  static constexpr int TOTAL = RING + MAX_PAGES * 4;
};

// swizzled byte offset of 16-byte chunk `ch` of row `r` (rows are HD*2 bytes)
template <int HD>
__device__ __forceinline__ uint32_t swz(int r, int ch) {
  return (uint32_t)(r * (HD * 2) + ((ch ^ (r & 7)) << 4));
}

template <int HD>
__global__ void __launch_bounds__(128) attn_decode_fast_kernel(FastAttnArgs a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(128) unsigned char smem[];
  using SM = FastAttnSmem<HD>;
  constexpr int CH = HD / 8;   // 16-byte chunks per row
  constexpr int KC = HD / 16;  // k-steps of the QK^T MMA
  constexpr int NT = HD / 8;   // n-tiles of the PV MMA

  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  const int G = a.n_q / a.kv.n_kv;
  const int kv_len = a.pos[b] + 1;
  const int psz = a.kv.page_size;
  const int* bt = a.block_table + (size_t)a.slot[b] * a.max_pages;

  const int tiles_total = (kv_len + TILE - 1) / TILE;
  const int tps = (tiles_total + a.n_splits - 1) / a.n_splits;
  const int t_begin = split * tps;
  const int t_end = min(tiles_total, t_begin + tps);
  const int n_tiles = max(0, t_end - t_begin);

  const uint32_t sbase = smem_u32(smem);
  // page ids of [t_begin*TILE, t_end*TILE) cached in shared memory (one global read each)
  int* s_pages = reinterpret_cast<int*>(smem + SM::RING);
  const int page0 = (t_begin * TILE) / psz;
  const int n_pg = n_tiles > 0 ? min(SM::MAX_PAGES, (min(kv_len, t_end * TILE) - 1) / psz - page0 + 1) : 0;
  for (int i = threadIdx.x; i < n_pg; i += 128) s_pages[i] = bt[page0 + i];
  __syncthreads();

  // ---- Q fragments: rows = the G query heads of this kv head (rows >= G are zero)
  uint32_t qf[KC][4];
  {
    const bf16* qrow = a.q + (size_t)b * a.n_q * HD + (size_t)(kvh * G) * HD;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      const int d0 = kc * 16 + c * 2;
      qf[kc][0] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + (size_t)g * HD + d0) : 0u;
      qf[kc][2] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + (size_t)g * HD + d0 + 8) : 0u;
      qf[kc][1] = (g + 8) < G ? *reinterpret_cast<const uint32_t*>(qrow + (size_t)(g + 8) * HD + d0) : 0u;
      qf[kc][3] = (g + 8) < G ? *reinterpret_cast<const uint32_t*>(qrow + (size_t)(g + 8) * HD + d0 + 8) : 0u;
    }
  }

  // every thread copies the same 16-byte chunk column of rows lr0, lr0 + RPP, ... of each tile: one page lookup
  // and one multiply-add per request instead of the generic i / CH, i % CH, p / psz, p % psz arithmetic
  constexpr int RPP = 128 / CH;
  const int lch = threadIdx.x % CH, lr0 = threadIdx.x / CH;
  const int psz_shift = (psz & (psz - 1)) == 0 ? __ffs(psz) - 1 : -1;
  const size_t v_off = a.kv.kv_stride();
  const size_t page_stride = a.kv.page_stride();
  const bf16* kv_base = a.kv.chunk(0, a.layer, 0, kvh) + lch * 8;
  auto load_tile = [&](int tile, int buf) {
    const uint32_t kb = sbase + buf * SM::STAGE_BYTES, vb = kb + SM::TILE_BYTES;
    const int p0 = tile * TILE;
#pragma unroll
    for (int k = 0; k < TILE / RPP; ++k) {
      const int r = lr0 + k * RPP;
      const int p = min(p0 + r, kv_len - 1);  // clamp: masked below
      const int pg = psz_shift >= 0 ? (p >> psz_shift) : p / psz;
      const int po = p - pg * psz;
      const int pi = pg - page0;
      const int page = pi < SM::MAX_PAGES ? s_pages[pi] : bt[pg];
      const bf16* src = kv_base + (size_t)page * page_stride + (size_t)po * HD;
      const uint32_t so = swz<HD>(r, lch);
      cp_async16(kb + so, src);
      cp_async16(vb + so, src + v_off);
    }
  };

  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl2 = a.scale_log2;

  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < n_tiles) load_tile(t_begin + s, s);
    cp_async_commit();
  }

  for (int it = 0; it < n_tiles; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nx = it + STAGES - 1;
      if (nx < n_tiles) load_tile(t_begin + nx, nx % STAGES);
      cp_async_commit();
    }
    const uint32_t kb = sbase + (it % STAGES) * SM::STAGE_BYTES, vb = kb + SM::TILE_BYTES;
    const int r0 = warp * 16;                       // this warp's 16 positions inside the tile
    const int pos0 = (t_begin + it) * TILE + r0;    // absolute position of row r0
    if (pos0 < kv_len) {
      // ---- S = Q K^T for 16 positions (two n8 tiles)
      float sc[2][4];
#pragma unroll
      for (int j = 0; j < 2; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        uint32_t k0, k1, k2, k3;
        const int r = r0 + ((lane >> 4) << 3) + (lane & 7);
        const int ch = 2 * kc + ((lane >> 3) & 1);
        ldmatrix_x4(k0, k1, k2, k3, kb + swz<HD>(r, ch));
        mma_bf16_16816(sc[0], qf[kc], k0, k1);
        mma_bf16_16816(sc[1], qf[kc], k2, k3);
      }
      // ---- scale, mask, online softmax (rows g and g+8)
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int p = pos0 + 8 * j + 2 * c + (e & 1);
          float v = p < kv_len ? sc[j][e] * sl2 : -INFINITY;
          sc[j][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      float corr[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
        const float m_new = fmaxf(m_run[h], mx[h]);  // finite: position pos0 is valid for every row
        corr[h] = exp2f(m_run[h] - m_new);           // m_run = -inf -> 0
        m_run[h] = m_new;
      }
      float rs[2] = {0.f, 0.f};
      uint32_t pf[4], pl[4];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float p0 = exp2f(sc[j][0] - m_run[0]), p1 = exp2f(sc[j][1] - m_run[0]);
        const float p2 = exp2f(sc[j][2] - m_run[1]), p3 = exp2f(sc[j][3] - m_run[1]);
        rs[0] += p0 + p1;
        rs[1] += p2 + p3;
        // P = hi + lo in bf16 (16 mantissa bits): the PV product keeps fp32-like weights at the
        // price of a second MMA, which is free in a kernel that waits for HBM
        const bf16 h0 = f2bf(p0), h1 = f2bf(p1), h2 = f2bf(p2), h3 = f2bf(p3);
        pf[2 * j] = pack2(h0, h1);      // a0 (j=0) / a2 (j=1): row g
        pf[2 * j + 1] = pack2(h2, h3);  // a1 / a3: row g+8
        pl[2 * j] = pack2(f2bf(p0 - bf2f(h0)), f2bf(p1 - bf2f(h1)));
        pl[2 * j + 1] = pack2(f2bf(p2 - bf2f(h2)), f2bf(p3 - bf2f(h3)));
      }
      l_run[0] = l_run[0] * corr[0] + rs[0];
      l_run[1] = l_run[1] * corr[1] + rs[1];
      const uint32_t pa[4] = {pf[0], pf[1], pf[2], pf[3]};
      const uint32_t pb[4] = {pl[0], pl[1], pl[2], pl[3]};
      // ---- O = O*corr + P V
#pragma unroll
      for (int j = 0; j < NT; j += 2) {
        o[j][0] *= corr[0];
        o[j][1] *= corr[0];
        o[j][2] *= corr[1];
        o[j][3] *= corr[1];
        o[j + 1][0] *= corr[0];
        o[j + 1][1] *= corr[0];
        o[j + 1][2] *= corr[1];
        o[j + 1][3] *= corr[1];
        uint32_t v0, v1, v2, v3;
        const int r = r0 + (((lane >> 3) & 1) << 3) + (lane & 7);
        const int ch = j + (lane >> 4);
        ldmatrix_x4_trans(v0, v1, v2, v3, vb + swz<HD>(r, ch));
        mma_bf16_16816(o[j], pa, v0, v1);
        mma_bf16_16816(o[j + 1], pa, v2, v3);
        mma_bf16_16816(o[j], pb, v0, v1);
        mma_bf16_16816(o[j + 1], pb, v2, v3);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();

  // ---- merge the 4 warps (each covered different positions) through shared memory
  float* red_o = reinterpret_cast<float*>(smem);                   // [4][16][HD]
  float* red_ml = red_o + 4 * 16 * HD;                             // [4][16][2]
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int d = j * 8 + 2 * c;
    red_o[(warp * 16 + g) * HD + d] = o[j][0];
    red_o[(warp * 16 + g) * HD + d + 1] = o[j][1];
    red_o[(warp * 16 + g + 8) * HD + d] = o[j][2];
    red_o[(warp * 16 + g + 8) * HD + d + 1] = o[j][3];
  }
  if (c == 0) {
    red_ml[(warp * 16 + g) * 2] = m_run[0];
    red_ml[(warp * 16 + g) * 2 + 1] = l_run[0];
    red_ml[(warp * 16 + g + 8) * 2] = m_run[1];
    red_ml[(warp * 16 + g + 8) * 2 + 1] = l_run[1];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < G * HD; i += 128) {
    const int r = i / HD, d = i % HD;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) M = fmaxf(M, red_ml[(w * 16 + r) * 2]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float mw = red_ml[(w * 16 + r) * 2];
      const float f = mw == -INFINITY ? 0.f : exp2f(mw - M);
      L += red_ml[(w * 16 + r) * 2 + 1] * f;
      O += red_o[(w * 16 + r) * HD + d] * f;
    }
    const int h = kvh * G + r;
    if (a.n_splits == 1) {
      a.out[(size_t)b * a.n_q * HD + (size_t)h * HD + d] = f2bf(O / L);
    } else {
      const size_t row = ((size_t)split * a.n_tok + b) * a.n_q + h;
      a.ws_o[row * HD + d] = O;
      if (d == 0) {
        a.ws_ml[row * 2] = M;
        a.ws_ml[row * 2 + 1] = L;
      }
    }
  }
}

// out[b,h,:] = sum_s O_s 2^(m_s-M) / sum_s l_s 2^(m_s-M)
__global__ void attn_combine_kernel(FastAttnArgs a, int hd) {
  pdl_wait();
  pdl_trigger();
  const int b = blockIdx.y, h = blockIdx.x, d = threadIdx.x;
  float M = -INFINITY;
  for (int s = 0; s < a.n_splits; ++s) M = fmaxf(M, a.ws_ml[(((size_t)s * a.n_tok + b) * a.n_q + h) * 2]);
  float L = 0.f, O = 0.f;
  for (int s = 0; s < a.n_splits; ++s) {
    const size_t row = ((size_t)s * a.n_tok + b) * a.n_q + h;
    const float ms = a.ws_ml[row * 2];
    const float f = ms == -INFINITY ? 0.f : exp2f(ms - M);
    L += a.ws_ml[row * 2 + 1] * f;
    O += a.ws_o[row * hd + d] * f;
  }
  a.out[(size_t)b * a.n_q * hd + (size_t)h * hd + d] = f2bf(O / L);
}

template <int HD>
static cudaError_t launch_hd(const FastAttnArgs& a, cudaStream_t st) {
  static PerDeviceOnce set;
  constexpr int smem = FastAttnSmem<HD>::TOTAL;
  static_assert(FastAttnSmem<HD>::RING >= 4 * 16 * HD * 4 + 4 * 16 * 2 * 4, "merge scratch must fit in the KV ring");
  if (set.need()) {
    cudaError_t e = cudaFuncSetAttribute(attn_decode_fast_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    set.done();
  }
  dim3 grid(a.n_splits, a.kv.n_kv, a.n_tok);
  (void)launch_k(attn_decode_fast_kernel<HD>, dim3(grid), dim3(128), smem, st, a);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (a.n_splits > 1) {
    (void)launch_k(attn_combine_kernel, dim3(dim3(a.n_q, a.n_tok)), dim3(HD), 0, st, a, HD);
    e = cudaGetLastError();
  }
  return e;
}

cudaError_t launch_attention_decode_fast(const FastAttnArgs& a, cudaStream_t st) {
  if (a.n_tok == 0) return cudaSuccess;
  if (a.n_q % a.kv.n_kv || a.n_q / a.kv.n_kv > 16) return cudaErrorInvalidValue;
  switch (a.kv.hd) {
    case 64: return launch_hd<64>(a, st);
    case 128: return launch_hd<128>(a, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace qie
