// fhfma_tp.cu -- issue cost of the mixed-precision FMA (fma.rn.f32.bf16 = FHFMA.BF16) against FFMA with unpacked
// operands, per SM sub-partition: 8 warps per block (2 per scheduler), 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fhfma_tp fhfma_tp.cu && ./fhfma_tp
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, const uint32_t* in, int iters, long long* cyc) {
  uint32_t w[8], x[8];
  for (int i = 0; i < 8; ++i) {
    w[i] = in[(threadIdx.x * 8 + i) & 1023];
    x[i] = in[(threadIdx.x * 8 + i + 512) & 1023];
  }
  float c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) {  // FHFMA.BF16: two per 32-bit word pair
        asm volatile("{\n.reg .b16 wl, wh, xl, xh;\nmov.b32 {wl, wh}, %2;\nmov.b32 {xl, xh}, %3;\n"
            "fma.rn.f32.bf16 %0, wl, xl, %0;\nfma.rn.f32.bf16 %1, wh, xh, %1;\n}\n"
            : "+f"(c[i]), "+f"(c[(i + 4) & 7]) : "r"(w[i]), "r"(x[i]));
      } else if (MODE == 1) {  // unpack (shift / mask) + FFMA
        const float wl = __uint_as_float(w[i] << 16), wh = __uint_as_float(w[i] & 0xffff0000u);
        const float xl = __uint_as_float(x[i] << 16), xh = __uint_as_float(x[i] & 0xffff0000u);
        c[i] = fmaf(wl, xl, c[i]);
        c[(i + 4) & 7] = fmaf(wh, xh, c[(i + 4) & 7]);
        w[i] += 0x10001u * (it & 1);
      } else {  // plain FFMA, operands already fp32
        c[i] = fmaf(__uint_as_float(w[i]), __uint_as_float(x[i]), c[i]);
        c[(i + 4) & 7] = fmaf(__uint_as_float(x[i]), __uint_as_float(w[i]), c[(i + 4) & 7]);
      }
    }
  }
  const long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out;
  uint32_t* in;
  long long* cyc;
  cudaMalloc(&out, 1 << 20);
  cudaMalloc(&in, 4096);
  cudaMemset(in, 0x3f, 4096);
  cudaMallocManaged(&cyc, 8);
  const int iters = 4096;
  const char* names[3] = {"FHFMA.BF16 (fma.rn.f32.bf16)", "shift/mask unpack + FFMA", "FFMA"};
  for (int warps = 4; warps <= 16; warps *= 2)
    for (int m = 0; m < 3; ++m) {
      for (int rep = 0; rep < 2; ++rep) {
        if (m == 0) k<0><<<1, warps * 32>>>(out, in, iters, cyc);
        if (m == 1) k<1><<<1, warps * 32>>>(out, in, iters, cyc);
        if (m == 2) k<2><<<1, warps * 32>>>(out, in, iters, cyc);
        cudaDeviceSynchronize();
      }
      const double fma_per_warp = (double)iters * 16;
      printf("%2d warps/SM  %-32s %8lld cycles  %.2f cycles per FMA warp-instruction per scheduler\n", warps, names[m], *cyc,
             (double)*cyc / (fma_per_warp * warps / 4.0));
    }
  return 0;
}
