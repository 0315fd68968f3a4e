// attn_prefill_tc.cu -- causal prefill attention on the 5th-gen tensor cores (tcgen05 + TMEM), head_dim 128.
//
// Same contract as attn_prefill_fast.cu (rows = consecutive positions of one sequence, K/V already in the paged
// pool, fast numerics), different machine mapping: a CTA owns 128 query rows of one head and walks the causal
// range in tiles of 128 keys;
//   S = Q K^T   : tcgen05.mma, A = Q tile, B = K tile (both K-major, 128-byte swizzle), D = 128 x 128 fp32 in TMEM
//   softmax     : warps 0-3, thread t = query row t = TMEM lane t reads its S row with tcgen05.ld, keeps the running
//                 max / sum, writes P (bf16) as the K-major A operand of the second product into shared memory
//   O_t = P V   : tcgen05.mma, B = V tile as stored in the pool ([key][hd]: N-contiguous = MN-major), fresh
//                 accumulator per tile in a second TMEM region; the softmax threads fold it into their fp32 row in
//                 registers (o = o * corr + O_t), so TMEM is never read-modify-written
// One elected thread of warp 4 issues the MMAs; all 160 threads copy K/V tiles with 16-byte cp.async into the
// swizzled layout the descriptors expect (pages are 16 positions: no single TMA box covers a tile), the next tile's
// copies are in flight while the current tile is multiplied.  Replaces the mma.sync kernel (141 TFLOP/s, bound by
// ldmatrix + softmax issue between the MMAs) for head_dim 128 and >= 256 rows.
#include <cuda.h>

#include "common.cuh"
#include "kernels.h"
#include "launch.h"

namespace qie {
namespace {

constexpr int TQ = 128;   // query rows per CTA
constexpr int TK = 128;   // keys per tile
constexpr int HDX = 128;  // head dim
constexpr int BLK = 128 * 64 * 2;            // one [128 rows][64 cols] bf16 column block = 16 KiB (8-row swizzle atoms)
constexpr int TILE_B = 2 * BLK;              // a 128 x 128 bf16 operand = 32 KiB
constexpr int SM_Q = 0, SM_P = TILE_B, SM_K0 = 2 * TILE_B;  // Q | P | K[2] | V[2]
constexpr int SM_V0 = SM_K0 + 2 * TILE_B;
constexpr int SM_BAR = SM_V0 + 2 * TILE_B;   // mbarriers + tmem slot
constexpr int SM_X = SM_BAR + 64;          // float xmax[2][128], xsum[2][128]: exchange between the two threads of a row
constexpr int SM_PAGES = SM_X + 2048;
constexpr int MAXP = 2048;
constexpr int SM_TOTAL = SM_PAGES + MAXP * 4 + 1024;
constexpr int NSOFT = 256;                  // 8 softmax warps: two threads per query row (keys / head-dim halves)
constexpr int NTHR = NSOFT + 32;            // + the copy/MMA warp

__device__ __forceinline__ void mbar_init_(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// (pure polling with test_wait instead of try_wait was measured: no difference)
__device__ __forceinline__ void mbar_wait_(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return;
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}
__device__ __forceinline__ void umma_(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,"
      "%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major operand, SWIZZLE_128B, 8-row atoms 1024 bytes apart (same encoding as gemm_tcgen05.cu)
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major operand (V as stored: [key][hd], hd contiguous), SWIZZLE_128B: column blocks of 64 hd values are `lbo`
// bytes apart, groups of 8 keys `sbo` bytes apart
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D = F32, A = B = BF16, N >> 3 at [17,23), M >> 4 at [24,29), bit 16: B is MN-major
__host__ __device__ constexpr uint32_t idesc_(int m, int n, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (b_mn ? (1u << 16) : 0u) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ float ex2(float x) {  // one MUFU instruction; 2^(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// byte offset of 16-byte chunk `ch` (0..15) of row r in a [128][128] bf16 operand made of two 64-column blocks
__device__ __forceinline__ uint32_t tile_off(int r, int ch) {
  return (uint32_t)((ch >> 3) * BLK + r * 128 + (((ch & 7) ^ (r & 7)) << 4));
}

}  // namespace

__global__ void __launch_bounds__(NTHR, 1) attn_prefill_tc_kernel(FastAttnArgs a, int lbo_sbo_swap,
                                                                  const __grid_constant__ CUtensorMap kvm, int use_tma) {
  pdl_wait();
  pdl_trigger();
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const uint32_t s0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char* sg = smem_raw + (s0 - smem_u32(smem_raw));
  // mbarriers: S product done (two: TMEM buffer t & 1; the copy path uses the first), P V done, K tile landed (two
  // buffers), V tile landed (two)
  const uint32_t bar_s = s0 + SM_BAR, bar_o = bar_s + 16, slot = bar_s + 24, bar_k = bar_s + 32, bar_v = bar_s + 48;
  int* s_pages = reinterpret_cast<int*>(sg + SM_PAGES);
  float* xmax = reinterpret_cast<float*>(sg + SM_X);  // [2][128]
  float* xsum = xmax + 256;                            // [2][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int MMA_WARP = NSOFT / 32;
  // CTAs are handed out x-fastest: all heads of the longest query tile first, then the next tile (longest-processing-
  // time order over the SMs; tile-fastest order left a 32-tile CTA for the end: makespan 61 vs 48 tile units at T = 4096)
  const int qt = gridDim.y - 1 - blockIdx.y, h = blockIdx.x;
  const int G = a.n_q / a.kv.n_kv, kvh = h / G;
  const int t0 = qt * TQ, n_rows = min(TQ, a.n_tok - t0);
  const int psz = a.kv.page_size;
  const int* bt = a.block_table + (size_t)a.slot[t0] * a.max_pages;
  const int pos_first = a.pos[t0];
  const int kv_len = pos_first + n_rows;
  const int n_tiles = (kv_len + TK - 1) / TK;
  const int Dq = a.n_q * HDX;

  if (threadIdx.x == 0) {
    mbar_init_(bar_s, 1);
    mbar_init_(bar_s + 8, 1);
    mbar_init_(bar_o, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init_(bar_k + 8 * i, 1);
      mbar_init_(bar_v + 8 * i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(512u) : "memory");  // S[2] (2 x 128 columns) + O (128): the next power of two
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const int n_pg = min(MAXP, (kv_len - 1) / psz + 1);
  for (int i = threadIdx.x; i < n_pg; i += NTHR) s_pages[i] = bt[i];
  for (int i = threadIdx.x; i < TQ * 16; i += NTHR) {  // Q tile
    const int r = i >> 4, ch = i & 15;
    const int t = t0 + min(r, n_rows - 1);
    cp_async16(s0 + SM_Q + tile_off(r, ch), a.q + (size_t)t * Dq + (size_t)h * HDX + ch * 8);
  }
  cp_async_commit();
  cp_async_wait<0>();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem) : "r"(slot));

  const size_t v_off = a.kv.kv_stride(), page_stride = a.kv.page_stride();
  const bf16* kv_base = a.kv.chunk(0, a.layer, 0, kvh);
  // 288 threads = 18 rows x 16 chunk columns per pass: a thread always copies chunk column `lch` of rows lr0 + 18 k
  // (one warp doing all copies put ~3k instructions in front of every S product: measured 1.8x slower)
  const int lch = threadIdx.x & 15, lr0 = threadIdx.x >> 4;
  const int psz_shift = (psz & (psz - 1)) == 0 ? __ffs(psz) - 1 : -1;
  const bf16* kv_col = kv_base + lch * 8;
  auto load_tile = [&](int tile, int buf) {
    const uint32_t kb = s0 + SM_K0 + buf * TILE_B, vb = s0 + SM_V0 + buf * TILE_B;
    const int p0 = tile * TK;
#pragma unroll 4
    for (int r = lr0; r < TK; r += NTHR / 16) {
      const int p = min(p0 + r, kv_len - 1);
      const int pi = psz_shift >= 0 ? (p >> psz_shift) : p / psz;
      const int po = p - pi * psz;
      const int page = pi < MAXP ? s_pages[pi] : bt[pi];
      const bf16* src = kv_col + (size_t)page * page_stride + (size_t)po * HDX;
      const uint32_t so = tile_off(r, lch);
      cp_async16(kb + so, src);
      cp_async16(vb + so, src + v_off);
    }
    cp_async_commit();
  };
  // K/V tiles by TMA (r02; pages of 16 positions): the 32 lanes of the MMA warp request the 2 x 8 x 2 boxes of a tile --
  // (K | V) x 8 page chunks x 2 column halves of 64, 2 KiB each, written with the 128-byte swizzle to exactly the place
  // tile_off() puts them -- so the eight softmax warps issue no copy instructions at all.  Chunks behind the last
  // page of the sequence repeat that page (finite values; their keys are masked).
  const int n_kv_pages = (kv_len - 1) / psz + 1;
  auto tma_kv = [&](int tile, int buf, int is_v) {  // MMA warp: lanes 0-15 request the 8 page chunks x 2 column halves of K or V
    const int pg = (lane >> 1) & 7, half = lane & 1;
    const int pi = min(tile * (TK / 16) + pg, n_kv_pages - 1);
    const int page = pi < MAXP ? s_pages[pi] : bt[pi];
    const int row = (((page * a.kv.n_layers + a.layer) * 2 + is_v) * a.kv.n_kv + kvh) * psz;
    const uint32_t dst = s0 + (is_v ? SM_V0 : SM_K0) + buf * TILE_B + half * BLK + pg * 16 * 128;
    const uint32_t bar = (is_v ? bar_v : bar_k) + 8u * (uint32_t)buf;
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TILE_B) : "memory");
    __syncwarp();
    if (lane < 16)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(dst), "l"(&kvm), "r"(bar), "r"(half * 64), "r"(row)
                   : "memory");
  };
  if (use_tma) {
    if (warp == MMA_WARP) {
      tma_kv(0, 0, 0);
      tma_kv(0, 0, 1);
      if (n_tiles > 1) tma_kv(1, 1, 0);
    }
  } else {
    load_tile(0, 0);
  }

  // Where the time goes (T = 4096, 12 heads, 171 us; probe builds that skip one part each): both MMA chains 35 us,
  // the exp + bf16-pack pass 53 us (16 384 ex2 + 8 192 cvt per tile on the 16-lane XU pipe), everything else --
  // tile copies, TMEM loads, row max, O fold, P stores, barriers, CTA prologue/epilogue -- 94 us.  An MN-major V
  // descriptor costs the same as a K-major one.
  // r02: with the cp.async copies, S in two TMEM buffers / S(t+1) issued before the softmax of tile t was a dead end (17.5 ->
  // 18.5 ms: the softmax warps are the critical path and a second copy pass cost them more than the S latency it hid);
  // with K/V by TMA the same schedule is free for them (the TMA path below).
  const uint32_t lbo = (lbo_sbo_swap & 1) ? 1024u : (uint32_t)BLK, sbo = (lbo_sbo_swap & 1) ? (uint32_t)BLK : 1024u;
  constexpr uint32_t ID_S = idesc_(128, 128, false), ID_O = idesc_(128, 128, true);
  const float sl2 = a.scale_log2;
  // softmax thread: query row `row` (= TMEM lane), half `hf`: keys [64 hf, 64 hf + 64) of every tile for the softmax,
  // head-dim columns [64 hf, 64 hf + 64) of O.  Warps w and w + 4 share a TMEM lane quarter (warp % 4).
  const bool soft = warp < MMA_WARP;
  const int row = (warp & 3) * 32 + lane, hf = (warp >> 2) & 1;
  const int my_pos = pos_first + row;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);
  float o[64];
  float m_run = -INFINITY, l_run = 0.f, corr = 1.f;
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;
  auto fold_o = [&]() {  // o = o * corr + (P V of the last tile), this thread's 64 head-dim columns
    uint32_t r[32], r2[32];
    tmem_ld32(lane_addr + 256 + 64 * hf, r);
    tmem_ld32(lane_addr + 256 + 64 * hf + 32, r2);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = fmaf(o[j], corr, __uint_as_float(r[j]));
#pragma unroll
    for (int j = 0; j < 32; ++j) o[32 + j] = fmaf(o[32 + j], corr, __uint_as_float(r2[j]));
  };

  // softmax of tile `it` for this thread's query row / key half: S from TMEM buffer sbuf, P (bf16) -> shared memory
  auto softmax_tile = [&](int it, int sbuf) {
    const int p0 = it * TK + 64 * hf;  // first key of this thread's half
    // ---- this thread's 64 scores stay in registers between the max and the exp pass
    uint32_t r[32], r2[32];
    tmem_ld32(lane_addr + (uint32_t)sbuf * 128u + 64 * hf, r);
    tmem_ld32(lane_addr + (uint32_t)sbuf * 128u + 64 * hf + 32, r2);
    tmem_wait_ld();
    const bool need_mask = p0 + 63 > my_pos;
    float mx = -INFINITY;
    if (need_mask) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (p0 + j > my_pos) r[j] = 0xff800000u;       // -inf
        if (p0 + 32 + j > my_pos) r2[j] = 0xff800000u;
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(r[j]), __uint_as_float(r2[j])));
    xmax[hf * 128 + row] = mx;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");  // the two warps that share these 32 rows meet
    const float m_new = fmaxf(m_run, fmaxf(mx, xmax[(hf ^ 1) * 128 + row]));
    corr = ex2((m_run - m_new) * sl2);
    m_run = m_new;
    const float nms = -m_new * sl2;
    float rs = 0.f;
    uint32_t pk[32];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float e0 = ex2(fmaf(__uint_as_float(r[j]), sl2, nms)), e1 = ex2(fmaf(__uint_as_float(r[j + 1]), sl2, nms));
      const float e2 = ex2(fmaf(__uint_as_float(r2[j]), sl2, nms)), e3 = ex2(fmaf(__uint_as_float(r2[j + 1]), sl2, nms));
      rs += (e0 + e1) + (e2 + e3);
      pk[j >> 1] = pack2(f2bf(e0), f2bf(e1));
      pk[16 + (j >> 1)] = pack2(f2bf(e2), f2bf(e3));
    }
#pragma unroll
    for (int q8 = 0; q8 < 8; ++q8)  // 8 chunks of 8 keys = this half's column block of the P row
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(s0 + SM_P + tile_off(row, 8 * hf + q8)), "r"(pk[4 * q8]),
                   "r"(pk[4 * q8 + 1]), "r"(pk[4 * q8 + 2]), "r"(pk[4 * q8 + 3])
                   : "memory");
    l_run = l_run * corr + rs;  // partial sum of this half; the halves are added once at the end
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // P (generic stores) -> tensor core
  };
  auto issue_s = [&](int t, int sbuf) {  // MMA thread: S(t) = Q K(t)^T into TMEM buffer sbuf
    const uint32_t kb = s0 + SM_K0 + (t & 1) * TILE_B;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int kk = 0; kk < HDX / 16; ++kk) {
      const uint32_t ko = (uint32_t)((kk >> 2) * BLK + (kk & 3) * 32);
      umma_(tmem + (uint32_t)sbuf * 128u, desc_kmajor(s0 + SM_Q + ko), desc_kmajor(kb + ko), ID_S, kk != 0);
    }
    umma_commit_(bar_s + 8u * (uint32_t)sbuf);
  };
  auto issue_pv = [&](int t) {  // MMA thread: O_t = P V(t)
    const uint32_t vb = s0 + SM_V0 + (t & 1) * TILE_B;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int kk = 0; kk < TK / 16; ++kk) {
      const uint32_t po = (uint32_t)((kk >> 2) * BLK + (kk & 3) * 32);   // P: K-major, K = keys
      const uint32_t vo = (uint32_t)(kk * 16 * 128);                     // V: 16 keys further down = 2 atoms of 8 rows
      umma_(tmem + 256, desc_kmajor(s0 + SM_P + po), desc_mnmajor(vb + vo, lbo, sbo), ID_O, kk != 0);
    }
    umma_commit_(bar_o);
  };

  if (use_tma) {
    // ---- TMA path, software pipeline: S(t+1) goes to the tensor pipe in front of the softmax of tile t (two S buffers in
    // TMEM); K(t+2) is requested as soon as S(t) is done, V(t+1) as soon as P V(t-1) is done.  The MMA warp does all of
    // that; the softmax warps only wait for mbarriers.
    if (warp == MMA_WARP && lane == 0) {
      mbar_wait_(bar_k, 0);
      issue_s(0, 0);
    }
    for (int it = 0; it < n_tiles; ++it) {
      const int buf = it & 1;
      if (warp == MMA_WARP) {
        if (lane == 0 && it + 1 < n_tiles) {
          mbar_wait_(bar_k + 8u * (uint32_t)(buf ^ 1), (uint32_t)(((it + 1) >> 1) & 1));  // K(it+1) has landed
          issue_s(it + 1, buf ^ 1);  // (everybody read S(it-1) out of that buffer in front of the last __syncthreads)
        }
        __syncwarp();
        if (it + 1 < n_tiles) {  // V(it+1) takes the place of V(it-1)
          if (it > 0) mbar_wait_(bar_o, (uint32_t)((it - 1) & 1));
          tma_kv(it + 1, buf ^ 1, 1);
        }
        if (it + 2 < n_tiles) {  // K(it+2) takes the place of K(it)
          mbar_wait_(bar_s + 8u * (uint32_t)buf, (uint32_t)((it >> 1) & 1));
          tma_kv(it + 2, buf, 0);
        }
      } else {
        if (it > 0) {  // fold the previous tile's P V product
          mbar_wait_(bar_o, (uint32_t)((it - 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          fold_o();
        }
        mbar_wait_(bar_s + 8u * (uint32_t)buf, (uint32_t)((it >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        softmax_tile(it, buf);
      }
      __syncthreads();
      if (warp == MMA_WARP && lane == 0) {
        mbar_wait_(bar_v + 8u * (uint32_t)buf, (uint32_t)((it >> 1) & 1));  // V(it) has landed
        issue_pv(it);
      }
    }
  } else {
  for (int it = 0; it < n_tiles; ++it) {
    const int buf = it & 1;
    cp_async_wait<0>();  // tile it (requested one iteration ago)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // cp.async (generic proxy) data -> tensor core (async proxy)
    __syncthreads();
    // ---- S = Q K^T
    if (warp == MMA_WARP && lane == 0) issue_s(it, 0);
    // Behind the S product (queued right after the previous tile's P V in the tensor pipe): the buffer tile it+1 goes
    // into was read by the P V product of tile it-1, so wait for that MMA, then request the copies.
    if (it > 0) mbar_wait_(bar_o, (uint32_t)((it - 1) & 1));
    if (it + 1 < n_tiles) load_tile(it + 1, buf ^ 1);
    if (soft) {
      if (it > 0) {  // fold the previous tile's P V product while the tensor core works on this tile's S
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        fold_o();
      }
      mbar_wait_(bar_s, (uint32_t)(it & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      softmax_tile(it, 0);
    }
    __syncthreads();
    // ---- O_t = P V
    if (warp == MMA_WARP && lane == 0) issue_pv(it);
  }
  }
  if (soft) {
    mbar_wait_(bar_o, (uint32_t)((n_tiles - 1) & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    fold_o();
    xsum[hf * 128 + row] = l_run;
    asm volatile("bar.sync %0, 64;" ::"r"(1 + (warp & 3)) : "memory");
    if (row < n_rows) {
      const float inv = 1.f / (l_run + xsum[(hf ^ 1) * 128 + row]);
      bf16* dst = a.out + (size_t)(t0 + row) * Dq + (size_t)h * HDX + 64 * hf;
#pragma unroll
      for (int c = 0; c < 64; c += 8) {
        uint4 v;
        v.x = pack2(f2bf(o[c] * inv), f2bf(o[c + 1] * inv));
        v.y = pack2(f2bf(o[c + 2] * inv), f2bf(o[c + 3] * inv));
        v.z = pack2(f2bf(o[c + 4] * inv), f2bf(o[c + 5] * inv));
        v.w = pack2(f2bf(o[c + 6] * inv), f2bf(o[c + 7] * inv));
        *reinterpret_cast<uint4*>(dst + c) = v;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == MMA_WARP) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// rows = consecutive positions of one sequence; head_dim 128 only
cudaError_t launch_attention_prefill_tc(const FastAttnArgs& a, int lbo_sbo_swap, cudaStream_t st) {
  if (a.n_tok == 0) return cudaSuccess;
  if (a.kv.hd != HDX || a.n_q % a.kv.n_kv) return cudaErrorInvalidValue;
  static PerDeviceOnce set;
  if (set.need()) {
    cudaError_t e = cudaFuncSetAttribute(attn_prefill_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL);
    if (e != cudaSuccess) return e;
    set.done();
  }
  // K/V by TMA: the pool as rows of one head (256 bytes), boxes of 16 slots x 64 columns; pages of 16 positions only
  // (QIE_ATTN_TC_TMA=0 keeps the cp.async copies: A/B knob)
  static const bool tma_env = [] { const char* v = getenv("QIE_ATTN_TC_TMA"); return !(v && v[0] == '0'); }();
  TensorMap2D tm{};
  const unsigned long long rows = (unsigned long long)a.kv.n_pages * a.kv.n_layers * 2ull * a.kv.n_kv * a.kv.page_size;
  const int use_tma = tma_env && a.kv.page_size == 16 && make_tensor_map_kv(&tm, a.kv.pool, rows, HDX, 16) == cudaSuccess;
  dim3 grid(a.n_q, (a.n_tok + TQ - 1) / TQ);
  (void)launch_k(attn_prefill_tc_kernel, grid, dim3(NTHR), (size_t)SM_TOTAL, st, a, lbo_sbo_swap, *reinterpret_cast<const CUtensorMap*>(&tm), use_tma);
  return cudaGetLastError();
}

}  // namespace qie
