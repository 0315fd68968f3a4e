// microbenchmark: latency / throughput of the legacy mma.sync.m16n8k16 bf16 path on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void k(long long* out, float* sink, int iters, uint32_t seed) {
  float acc[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) acc[c][i] = 0.f;
  uint32_t a0 = seed * (threadIdx.x + 1), a1 = a0 ^ 0x3f803f80, b0 = 0x3f803f80, b1 = 0x3c003c00;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma(acc[c], a0, a1, a0, a1, b0, b1);
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += acc[c][i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x % 32 == 0) out[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int CHAINS>
void run(int warps, int iters) {
  long long* d; float* sink;
  cudaMalloc(&d, 1024 * sizeof(long long)); cudaMalloc(&sink, 148 * 1024 * sizeof(float));
  k<CHAINS><<<148, warps * 32>>>(d, sink, iters, 12345u);
  k<CHAINS><<<148, warps * 32>>>(d, sink, iters, 12345u);
  cudaDeviceSynchronize();
  long long h[64];
  cudaMemcpy(h, d, warps * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < warps; ++i) mx = h[i] > mx ? h[i] : mx;
  double per_step = (double)mx / iters;          // cycles for CHAINS mma per warp
  double per_smsp = per_step / CHAINS / ((warps + 3) / 4 > 0 ? 1.0 : 1.0);
  printf("chains/warp %d warps/CTA %2d : %.1f cyc per step (%.1f cyc per HMMA per warp); HMMA per SM per cycle %.4f -> %.0f TFLOP/s chip @1.9GHz\n",
         CHAINS, warps, per_step, per_smsp, CHAINS * warps / per_step, CHAINS * warps / per_step * 4096 * 148 * 1.9e9 / 1e12);
  cudaFree(d); cudaFree(sink);
}

int main() {
  const int iters = 4096;
  run<1>(1, iters); run<2>(1, iters); run<4>(1, iters); run<8>(1, iters);
  run<1>(4, iters); run<4>(4, iters); run<8>(4, iters);
  run<1>(8, iters); run<4>(8, iters); run<8>(8, iters);
  run<1>(16, iters); run<4>(16, iters); run<8>(16, iters);
  run<4>(32, iters);
  return 0;
}
