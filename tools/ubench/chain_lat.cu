// microbenchmark: one warp walking the decode_mega MMA chunk loop (LDSM B + predicated LDS A + HMMA chain)
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../qwen_inference_engine_b200/csrc/common.cuh"
using namespace qie;
__device__ __forceinline__ uint32_t lds32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }

template <int VARIANT>
__device__ __forceinline__ void chunk(float (&acc)[4], uint32_t slot_addr, uint32_t a_addr, int AS, int nk16, int B, int lane) {
  const int br = lane & 7, bh = (lane >> 3) & 1;
  const uint32_t brow = slot_addr + br * 128;
  uint32_t boff[4];
#pragma unroll
  for (int s = 0; s < 4; ++s) boff[s] = (uint32_t)((((s << 1) + bh) ^ br) << 4);
  const bool row_valid = (lane >> 2) < B;
  const uint32_t a0 = a_addr + (lane >> 2) * AS + (lane & 3) * 4;
  if (VARIANT == 0) {  // simple loop, compiler-scheduled
#pragma unroll 4
    for (int j = 0; j < nk16; ++j) {
      uint32_t b0, b1;
      ldmatrix_x2(b0, b1, brow + (j >> 2) * 1024 + boff[j & 3]);
      uint32_t af[4] = {0u, 0u, 0u, 0u};
      if (row_valid) { af[0] = lds32(a0 + j * 32); af[2] = lds32(a0 + j * 32 + 16); }
      mma_bf16_16816(acc, af, b0, b1);
    }
  } else if (VARIANT == 1) {  // explicit double buffer in groups of 4
    uint32_t fb[2][4][2], fa[2][4][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) fa[q][s][i] = 0u;
    auto load = [&](int q, int j0) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int j = j0 + s;
        ldmatrix_x2(fb[q][s][0], fb[q][s][1], brow + (uint32_t)(j >> 2) * 1024u + boff[s]);
        if (row_valid) { fa[q][s][0] = lds32(a0 + j * 32); fa[q][s][2] = lds32(a0 + j * 32 + 16); }
      }
    };
    auto compute = [&](int q) {
#pragma unroll
      for (int s = 0; s < 4; ++s) mma_bf16_16816(acc, fa[q][s], fb[q][s][0], fb[q][s][1]);
    };
    load(0, 0);
    for (int j0 = 0; j0 < nk16; j0 += 8) {
      const bool more1 = j0 + 4 < nk16;
      if (more1) load(1, j0 + 4);
      compute(0);
      if (more1) { if (j0 + 8 < nk16) load(0, j0 + 8); compute(1); }
    }
  } else if (VARIANT == 3) {  // double buffer, B fragments by plain LDS.32 (no ldmatrix / WARPSYNC)
    uint32_t fb[2][4][2], fa[2][4][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) fa[q][s][i] = 0u;
    const int g = lane >> 2, c = lane & 3;
    const uint32_t wrow = slot_addr + g * 128 + c * 4;  // row g, element 2c within a 16-byte chunk
    auto load = [&](int q, int j0) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int j = j0 + s;
        const uint32_t kb = (uint32_t)(j >> 2) * 1024u;
        fb[q][s][0] = lds32(wrow + kb + ((((s << 1)) ^ g) << 4));
        fb[q][s][1] = lds32(wrow + kb + ((((s << 1) + 1) ^ g) << 4));
        if (row_valid) { fa[q][s][0] = lds32(a0 + j * 32); fa[q][s][2] = lds32(a0 + j * 32 + 16); }
      }
    };
    auto compute = [&](int q) {
#pragma unroll
      for (int s = 0; s < 4; ++s) mma_bf16_16816(acc, fa[q][s], fb[q][s][0], fb[q][s][1]);
    };
    load(0, 0);
    for (int j0 = 0; j0 < nk16; j0 += 8) {
      const bool more1 = j0 + 4 < nk16;
      if (more1) load(1, j0 + 4);
      compute(0);
      if (more1) { if (j0 + 8 < nk16) load(0, j0 + 8); compute(1); }
    }
  } else if (VARIANT == 6 || VARIANT == 7) {  // first MMA of the group, THEN the next group's loads, then the other 3 MMAs
    uint32_t fb[2][4][2], fa[2][4][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) fa[q][s][i] = 0u;
    auto load = [&](int q, int j0) {
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        const int j = j0 + s;
        ldmatrix_x2(fb[q][s][0], fb[q][s][1], brow + (uint32_t)(j >> 2) * 1024u + boff[s]);
        if (row_valid) { fa[q][s][0] = lds32(a0 + j * 32); fa[q][s][2] = lds32(a0 + j * 32 + 16); }
      }
    };
    load(0, 0);
    for (int j0 = 0; j0 < nk16; j0 += 8) {
      const bool more1 = j0 + 4 < nk16;
      mma_bf16_16816(acc, fa[0][0], fb[0][0][0], fb[0][0][1]);
      if (VARIANT == 7) mma_bf16_16816(acc, fa[0][1], fb[0][1][0], fb[0][1][1]);
      if (more1) load(1, j0 + 4);
      if (VARIANT == 6) mma_bf16_16816(acc, fa[0][1], fb[0][1][0], fb[0][1][1]);
      mma_bf16_16816(acc, fa[0][2], fb[0][2][0], fb[0][2][1]);
      mma_bf16_16816(acc, fa[0][3], fb[0][3][0], fb[0][3][1]);
      if (more1) {
        mma_bf16_16816(acc, fa[1][0], fb[1][0][0], fb[1][0][1]);
        if (VARIANT == 7) mma_bf16_16816(acc, fa[1][1], fb[1][1][0], fb[1][1][1]);
        if (j0 + 8 < nk16) load(0, j0 + 8);
        if (VARIANT == 6) mma_bf16_16816(acc, fa[1][1], fb[1][1][0], fb[1][1][1]);
        mma_bf16_16816(acc, fa[1][2], fb[1][2][0], fb[1][2][1]);
        mma_bf16_16816(acc, fa[1][3], fb[1][3][0], fb[1][3][1]);
      }
    }
  } else if (VARIANT == 8) {  // fine-grained interleave: after each MMA, the loads of the same step of the next group
    uint32_t fb[2][4][2], fa[2][4][4];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int i = 0; i < 4; ++i) fa[q][s][i] = 0u;
    auto load1 = [&](int q, int s, int j) {
      ldmatrix_x2(fb[q][s][0], fb[q][s][1], brow + (uint32_t)(j >> 2) * 1024u + boff[s]);
      if (row_valid) { fa[q][s][0] = lds32(a0 + j * 32); fa[q][s][2] = lds32(a0 + j * 32 + 16); }
    };
#pragma unroll
    for (int s = 0; s < 4; ++s) load1(0, s, s);
    for (int j0 = 0; j0 < nk16; j0 += 8) {
      const bool more1 = j0 + 4 < nk16, more2 = j0 + 8 < nk16;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        mma_bf16_16816(acc, fa[0][s], fb[0][s][0], fb[0][s][1]);
        if (more1) load1(1, s, j0 + 4 + s);
      }
      if (more1) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          mma_bf16_16816(acc, fa[1][s], fb[1][s][0], fb[1][s][1]);
          if (more2) load1(0, s, j0 + 8 + s);
        }
      }
    }
  } else if (VARIANT == 4) {  // no loads at all in the loop: operands fixed in registers (pure chain)
    uint32_t b0, b1; ldmatrix_x2(b0, b1, brow);
    uint32_t af[4] = {0u, 0u, 0u, 0u};
    if (row_valid) { af[0] = lds32(a0); af[2] = lds32(a0 + 16); }
    for (int j = 0; j < nk16; ++j) mma_bf16_16816(acc, af, b0, b1);
  } else if (VARIANT == 5) {  // 8 preloaded operand sets rotated, no loads in the loop
    uint32_t fb[8][2], fa[8][4];
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      ldmatrix_x2(fb[s][0], fb[s][1], brow + (s >> 2) * 1024 + boff[s & 3]);
      fa[s][0] = fa[s][1] = fa[s][2] = fa[s][3] = 0u;
      if (row_valid) { fa[s][0] = lds32(a0 + s * 32); fa[s][2] = lds32(a0 + s * 32 + 16); }
    }
    for (int j0 = 0; j0 < nk16; j0 += 8) {
#pragma unroll
      for (int s = 0; s < 8; ++s) mma_bf16_16816(acc, fa[s], fb[s][0], fb[s][1]);
    }
  } else {  // swapped operands: weights as A (16 rows), tokens as B (n = 8): no zero rows at all
    // not bit-compatible by construction; latency comparison only
#pragma unroll 4
    for (int j = 0; j < nk16; ++j) {
      uint32_t a[4];
      ldmatrix_x4(a[0], a[1], a[2], a[3], brow + (j >> 2) * 1024 + boff[j & 3]);
      uint32_t b0 = 0, b1 = 0;
      if (row_valid) { b0 = lds32(a0 + j * 32); b1 = lds32(a0 + j * 32 + 16); }
      mma_bf16_16816(acc, a, b0, b1);
    }
  }
}

template <int VARIANT>
__global__ void k(long long* out, float* sink, int reps, int nk16, int B) {
  extern __shared__ __align__(1024) unsigned char smem[];
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  __syncthreads();
  const uint32_t base = smem_u32(smem);
  float acc[4] = {0, 0, 0, 0};
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) chunk<VARIANT>(acc, base, base + 16384, (nk16 * 16 + 8) * 2, nk16, B, threadIdx.x & 31);
  long long t1 = clock64();
  sink[threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int V> void run(const char* name, int nk16, int B) {
  long long* d; float* sink; cudaMalloc(&d, 8); cudaMalloc(&sink, 4096);
  cudaFuncSetAttribute(k<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  const int reps = 200;
  k<V><<<1, 32, 65536>>>(d, sink, reps, nk16, B);
  k<V><<<1, 32, 65536>>>(d, sink, reps, nk16, B);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  printf("%-28s nk16=%3d B=%d : %.1f cycles per k16 step (%s)\n", name, nk16, B, (double)h / reps / nk16, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  run<0>("simple loop", 56, 1); run<1>("double-buffered groups of 4", 56, 1); run<2>("swapped operands", 56, 1);
  run<3>("double-buffered, B by LDS.32", 56, 1); run<4>("no loads (fixed operands)", 56, 1); run<5>("no loads (8 operand sets)", 56, 1);
  run<6>("mma0, loads(next), mma1-3", 56, 1); run<7>("mma0-1, loads(next), mma2-3", 56, 1);
  run<8>("interleaved mma/load per step", 56, 1); run<8>("interleaved mma/load per step", 28, 1); run<8>("interleaved mma/load per step", 56, 8);
  run<0>("simple loop", 28, 1); run<1>("double-buffered groups of 4", 28, 1);
  run<0>("simple loop", 56, 8); run<1>("double-buffered groups of 4", 56, 8);
  return 0;
}
