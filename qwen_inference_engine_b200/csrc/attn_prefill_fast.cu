// attn_prefill_fast.cu -- causal tiled attention for prefill rows over the paged KV pool
// (fast numerics).
//
// Replaces selfattention (/root/reference/layers/src/self_attension.cu:10-149) for the
// prefill branch of llm() (qwen_main.cu:160-176): there one block per head loops over the
// query tokens, walks the page list for every (token, kv position) pair and runs the softmax
// in one thread -- O(T^2) serial work per head.  Here the T x t score matrix is tiled the
// FlashAttention-2 way; results agree with the reference within the north-star 1e-2 bf16
// tolerance (tests/test_gpu_fast_path.py); ops_ref_order.cu stays the bit-exact path.
//
// Mapping (tensor-pipe bound for long prompts: 4*hd*T^2/2 flop per head):
//   rows  : the chunk's query rows are consecutive positions pos[0]+t of ONE sequence; their
//           K/V are already in the pool (qkv_post_kernel runs first), so chunked prefill with
//           an existing cache prefix is the same code path
//   grid  = (ceil(T/64), n_q heads); the longest (last) query tiles are scheduled first
//   block = 4 warps x 16 query rows; Q tile -> registers once (ldmatrix)
//   KV    : 64-position tiles, page chunks [slot][hd] contiguous in the pool, 16-byte cp.async
//           into a 2-stage XOR-swizzled ring; ldmatrix (K) / ldmatrix.trans (V)
//   math  : S = Q K^T and O += P V on mma.sync m16n8k16 bf16 (fp32 accumulate), online
//           softmax in fp32 with exp2f and the 1/sqrt(hd) scale folded in, P split into bf16 hi + lo
//           parts for the PV product (the kernel is bound by ldmatrix traffic, not by the MMAs); the causal mask is applied only on tiles that cross the diagonal
#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace qie {

static constexpr int PF_QT = 64;     // query rows per CTA
static constexpr int PF_KT = 64;     // kv positions per pipeline stage
static constexpr int PF_STAGES = 2;
static constexpr int PF_MAX_PAGES = 1024;  // page ids cached in shared memory

template <int HD>
struct PrefillAttnSmem {
  static constexpr int ROW_BYTES = HD * 2;
  static constexpr int Q_BYTES = PF_QT * ROW_BYTES;
  static constexpr int TILE_BYTES = PF_KT * ROW_BYTES;
  static constexpr int STAGE_BYTES = 2 * TILE_BYTES;
  static constexpr int RING = PF_STAGES * STAGE_BYTES;
  // the Q tile is only needed until its fragments sit in registers: it borrows the K area of
  // stage 1 (same size), so a CTA needs 64 KiB + page ids at hd 128 and three CTAs fit on an SM
  static constexpr int TOTAL = RING + PF_MAX_PAGES * 4;
  static_assert(Q_BYTES == TILE_BYTES, "Q tile aliases one K tile");
};

template <int HD>
__device__ __forceinline__ uint32_t pf_swz(int r, int ch) {
  return (uint32_t)(r * (HD * 2) + ((ch ^ (r & 7)) << 4));
}

template <int HD>
__global__ void __launch_bounds__(128, 3) attn_prefill_fast_kernel(FastAttnArgs a) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(128) unsigned char smem[];
  using SM = PrefillAttnSmem<HD>;
  constexpr int CH = HD / 8;   // 16-byte chunks per row
  constexpr int KC = HD / 16;  // k-steps of Q K^T
  constexpr int NT = HD / 8;   // n-tiles of P V
  constexpr int SN = PF_KT / 8;  // n-tiles of S per warp (8)

  const int qt = gridDim.y - 1 - blockIdx.y;  // heavy tiles first, all heads of a tile before the next tile
  const int h = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, c = lane & 3;
  const int G = a.n_q / a.kv.n_kv;
  const int kvh = h / G;
  const int t0 = qt * PF_QT;
  const int n_rows = min(PF_QT, a.n_tok - t0);
  const int psz = a.kv.page_size;
  const int* bt = a.block_table + (size_t)a.slot[t0] * a.max_pages;
  const int pos_first = a.pos[t0];                 // position of row t0; row t0+r sits at pos_first + r
  const int kv_len = pos_first + n_rows;           // positions visible to the last row of the tile
  const int n_tiles = (kv_len + PF_KT - 1) / PF_KT;
  const int Dq = a.n_q * HD;

  const uint32_t sring = smem_u32(smem);
  const uint32_t sq = sring + SM::STAGE_BYTES;  // K area of stage 1
  int* s_pages = reinterpret_cast<int*>(smem + SM::RING);
  const int n_pg = min(PF_MAX_PAGES, (kv_len - 1) / psz + 1);
  for (int i = threadIdx.x; i < n_pg; i += 128) s_pages[i] = bt[i];

  // ---- Q tile -> shared memory (rows past the chunk are clamped; never stored)
  for (int i = threadIdx.x; i < PF_QT * CH; i += 128) {
    const int r = i / CH, ch = i % CH;
    const int t = t0 + min(r, n_rows - 1);
    cp_async16(sq + pf_swz<HD>(r, ch), a.q + (size_t)t * Dq + (size_t)h * HD + ch * 8);
  }
  cp_async_commit();
  __syncthreads();  // s_pages visible

  // Every thread copies the same 16-byte chunk column `lch` of rows lr0, lr0 + RPP, ... of every tile, so the
  // address arithmetic per request is one page lookup and one multiply-add (the generic i / CH, i % CH, p / psz,
  // p % psz form cost as many instructions per tile as the softmax).
  constexpr int RPP = 128 / CH;  // rows covered by one pass of the CTA
  const int lch = threadIdx.x % CH, lr0 = threadIdx.x / CH;
  const int psz_shift = (psz & (psz - 1)) == 0 ? __ffs(psz) - 1 : -1;
  const size_t v_off = a.kv.kv_stride();  // V chunk of a (page, layer, head) sits kv_stride elements behind its K chunk
  const bf16* kv_base = a.kv.chunk(0, a.layer, 0, kvh) + lch * 8;
  const size_t page_stride = a.kv.page_stride();
  auto load_tile = [&](int tile, int buf) {
    const uint32_t kb = sring + buf * SM::STAGE_BYTES, vb = kb + SM::TILE_BYTES;
    const int p0 = tile * PF_KT;
#pragma unroll
    for (int k = 0; k < PF_KT / RPP; ++k) {
      const int r = lr0 + k * RPP;
      const int p = min(p0 + r, kv_len - 1);  // clamp: masked below
      const int pi = psz_shift >= 0 ? (p >> psz_shift) : p / psz;
      const int po = p - pi * psz;
      const int page = pi < PF_MAX_PAGES ? s_pages[pi] : bt[pi];
      const bf16* src = kv_base + (size_t)page * page_stride + (size_t)po * HD;
      const uint32_t so = pf_swz<HD>(r, lch);
      cp_async16(kb + so, src);
      cp_async16(vb + so, src + v_off);
    }
  };
  load_tile(0, 0);
  cp_async_commit();

  // ---- Q fragments of this warp's 16 rows
  cp_async_wait<1>();
  __syncthreads();
  uint32_t qf[KC][4];
  {
    const int r = warp * 16 + (lane & 15);
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) ldmatrix_x4(qf[kc][0], qf[kc][1], qf[kc][2], qf[kc][3], sq + pf_swz<HD>(r, 2 * kc + (lane >> 4)));
  }
  __syncthreads();  // stage 1 is free for K/V tile 1 from here on

  float o[NT][4];
#pragma unroll
  for (int j = 0; j < NT; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const float sl2 = a.scale_log2;
  // P = hi + lo (bf16) for short contexts, where few positions average the rounding error of P;
  // beyond 1024 positions the single bf16 term is below the output's own bf16 rounding
  const bool two_term = kv_len <= 1024;
  // last visible position of rows g and g+8 of this warp
  const int lim0 = pos_first + warp * 16 + g, lim1 = lim0 + 8;
  const int warp_last = pos_first + min(warp * 16 + 15, n_rows - 1);  // beyond this nothing is visible to the warp

  for (int it = 0; it < n_tiles; ++it) {
    if (it + 1 < n_tiles) load_tile(it + 1, (it + 1) & 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint32_t kb = sring + (it & 1) * SM::STAGE_BYTES, vb = kb + SM::TILE_BYTES;
    const int p0 = it * PF_KT;
    if (p0 <= warp_last) {
      // ---- S = Q K^T : 16 rows x 64 positions
      float sc[SN][4];
#pragma unroll
      for (int j = 0; j < SN; ++j) sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
#pragma unroll
        for (int np = 0; np < SN / 2; ++np) {
          uint32_t k0, k1, k2, k3;
          const int r = np * 16 + ((lane >> 4) << 3) + (lane & 7);
          const int ch = 2 * kc + ((lane >> 3) & 1);
          ldmatrix_x4(k0, k1, k2, k3, kb + pf_swz<HD>(r, ch));
          mma_bf16_16816(sc[2 * np], qf[kc], k0, k1);
          mma_bf16_16816(sc[2 * np + 1], qf[kc], k2, k3);
        }
      }
      // ---- causal mask on tiles that cross the diagonal, online softmax on the raw scores: the
      // 1/sqrt(hd) * log2(e) factor rides in the FFMA in front of exp2 (p = 2^(s*c - m*c))
      const bool need_mask = p0 + PF_KT - 1 > pos_first + warp * 16;
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < SN; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = sc[j][e];
          if (need_mask) {
            const int p = p0 + 8 * j + 2 * c + (e & 1);
            if (p > ((e >> 1) ? lim1 : lim0)) v = -INFINITY;
            sc[j][e] = v;
          }
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      float corr[2], nms[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
        const float m_new = fmaxf(m_run[hh], mx[hh]);
        // every row sees position 0 in tile 0, so m_new is finite from the first tile on
        corr[hh] = exp2f((m_run[hh] - m_new) * sl2);  // m_run = -inf -> 0
        m_run[hh] = m_new;
        nms[hh] = -m_new * sl2;
      }
      float rs[2] = {0.f, 0.f};
      uint32_t pa[SN / 2][4], pb[SN / 2][4];  // P = hi + lo in bf16 (16 mantissa bits)
#pragma unroll
      for (int j = 0; j < SN; ++j) {
        const float p0v = exp2f(fmaf(sc[j][0], sl2, nms[0])), p1v = exp2f(fmaf(sc[j][1], sl2, nms[0]));
        const float p2v = exp2f(fmaf(sc[j][2], sl2, nms[1])), p3v = exp2f(fmaf(sc[j][3], sl2, nms[1]));
        rs[0] += p0v + p1v;
        rs[1] += p2v + p3v;
        const bf16 h0 = f2bf(p0v), h1 = f2bf(p1v), h2 = f2bf(p2v), h3 = f2bf(p3v);
        pa[j >> 1][(j & 1) * 2] = pack2(h0, h1);      // a0 / a2: row g
        pa[j >> 1][(j & 1) * 2 + 1] = pack2(h2, h3);  // a1 / a3: row g+8
        if (two_term) {
          pb[j >> 1][(j & 1) * 2] = pack2(f2bf(p0v - bf2f(h0)), f2bf(p1v - bf2f(h1)));
          pb[j >> 1][(j & 1) * 2 + 1] = pack2(f2bf(p2v - bf2f(h2)), f2bf(p3v - bf2f(h3)));
        }
      }
      l_run[0] = l_run[0] * corr[0] + rs[0];
      l_run[1] = l_run[1] * corr[1] + rs[1];
      // the accumulator is rescaled only when some row of the warp saw a new maximum (rare once the
      // context is long)
      if (__any_sync(0xffffffffu, corr[0] != 1.f || corr[1] != 1.f)) {
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          o[j][0] *= corr[0];
          o[j][1] *= corr[0];
          o[j][2] *= corr[1];
          o[j][3] *= corr[1];
        }
      }
      // ---- O += P V
#pragma unroll
      for (int kk = 0; kk < SN / 2; ++kk) {
#pragma unroll
        for (int j = 0; j < NT; j += 2) {
          uint32_t v0, v1, v2, v3;
          const int r = kk * 16 + (((lane >> 3) & 1) << 3) + (lane & 7);
          const int ch = j + (lane >> 4);
          ldmatrix_x4_trans(v0, v1, v2, v3, vb + pf_swz<HD>(r, ch));
          mma_bf16_16816(o[j], pa[kk], v0, v1);
          mma_bf16_16816(o[j + 1], pa[kk], v2, v3);
          if (two_term) {
            mma_bf16_16816(o[j], pb[kk], v0, v1);
            mma_bf16_16816(o[j + 1], pb[kk], v2, v3);
          }
        }
      }
    }
    __syncthreads();  // the stage is refilled by the next iteration's load
  }
  cp_async_wait<0>();

  // ---- normalise and store
#pragma unroll
  for (int hh = 0; hh < 2; ++hh) {
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 1);
    l_run[hh] += __shfl_xor_sync(0xffffffffu, l_run[hh], 2);
  }
  const float inv0 = 1.f / l_run[0], inv1 = 1.f / l_run[1];
  const int r0 = warp * 16 + g, r1 = r0 + 8;
#pragma unroll
  for (int j = 0; j < NT; ++j) {
    const int d = j * 8 + 2 * c;
    if (r0 < n_rows)
      *reinterpret_cast<uint32_t*>(a.out + (size_t)(t0 + r0) * Dq + (size_t)h * HD + d) =
          pack2(f2bf(o[j][0] * inv0), f2bf(o[j][1] * inv0));
    if (r1 < n_rows)
      *reinterpret_cast<uint32_t*>(a.out + (size_t)(t0 + r1) * Dq + (size_t)h * HD + d) =
          pack2(f2bf(o[j][2] * inv1), f2bf(o[j][3] * inv1));
  }
}

template <int HD>
static cudaError_t launch_pf_hd(const FastAttnArgs& a, cudaStream_t st) {
  static PerDeviceOnce set;
  constexpr int smem = PrefillAttnSmem<HD>::TOTAL;
  if (set.need()) {
    cudaError_t e = cudaFuncSetAttribute(attn_prefill_fast_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    set.done();
  }
  dim3 grid(a.n_q, (a.n_tok + PF_QT - 1) / PF_QT);
  (void)launch_k(attn_prefill_fast_kernel<HD>, grid, dim3(128), smem, st, a);
  return cudaGetLastError();
}

// rows = consecutive positions of one sequence (pos[t] = pos[0] + t, slot[t] = slot[0])
cudaError_t launch_attention_prefill_fast(const FastAttnArgs& a, cudaStream_t st) {
  if (a.n_tok == 0) return cudaSuccess;
  if (a.n_q % a.kv.n_kv) return cudaErrorInvalidValue;
  switch (a.kv.hd) {
    case 64: return launch_pf_hd<64>(a, st);
    case 128: return launch_pf_hd<128>(a, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace qie
