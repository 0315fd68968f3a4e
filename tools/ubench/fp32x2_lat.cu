// microbenchmark: latency / throughput of the packed fp32x2 instructions the attention phase is built from
// (FFMA2 chains of the PV product, FADD chain of the softmax sum, FMUL2/FFMA2/FADD2 trees of the scores).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_lat fp32x2_lat.cu && ./fp32x2_lat
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ float2 cvt2(uint32_t v) { return make_float2(__uint_as_float(__byte_perm(v, 0u, 0x1044)), __uint_as_float(v & 0xffff0000u)); }

template <int V>
__global__ void k(float* out, long long* cyc, int n, float pin, uint32_t vin) {
  float2 o0 = make_float2(0.f, 0.f), o1 = make_float2(1.f, 1.f), o2 = make_float2(2.f, 2.f), o3 = make_float2(3.f, 3.f);
  float p = pin + threadIdx.x * 1e-9f;
  uint32_t vw = vin + threadIdx.x;
  float2 v = cvt2(vw);
  float s = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  if (V == 0) {  // one dependent FFMA2 chain
#pragma unroll 16
    for (int i = 0; i < n; ++i) o0 = __ffma2_rn(make_float2(p, p), v, o0);
  } else if (V == 1) {  // two independent FFMA2 chains
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
      o0 = __ffma2_rn(make_float2(p, p), v, o0);
      o1 = __ffma2_rn(make_float2(p, p), v, o1);
    }
  } else if (V == 2) {  // four independent FFMA2 chains
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
      o0 = __ffma2_rn(make_float2(p, p), v, o0);
      o1 = __ffma2_rn(make_float2(p, p), v, o1);
      o2 = __ffma2_rn(make_float2(p, p), v, o2);
      o3 = __ffma2_rn(make_float2(p, p), v, o3);
    }
  } else if (V == 3) {  // scalar FFMA: one chain of pairs (x, y independent)
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
      o0.x = __fmaf_rn(p, v.x, o0.x);
      o0.y = __fmaf_rn(p, v.y, o0.y);
    }
  } else if (V == 4) {  // dependent FADD chain
#pragma unroll 16
    for (int i = 0; i < n; ++i) s = __fadd_rn(s, p);
  } else if (V == 5) {  // one FFMA2 chain + widening of a bf16 pair per step (PRMT + LOP3)
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
      const float2 vv = cvt2(vw + i);
      o0 = __ffma2_rn(make_float2(p, p), vv, o0);
    }
  } else if (V == 6) {  // two chains + one widening per step
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
      const float2 vv = cvt2(vw + i);
      o0 = __ffma2_rn(make_float2(p, p), vv, o0);
      o1 = __ffma2_rn(make_float2(p, p), vv, o1);
    }
  } else if (V == 7) {  // 8 independent FMUL2 + FFMA2 + FADD2 streams (throughput)
    float2 a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = make_float2(p + j, p - j);
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = __ffma2_rn(a[j], v, __fmul2_rn(a[j], v));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) o0 = __fadd2_rn(o0, a[j]);
  } else if (V == 8) {  // 8 independent scalar FFMA streams (throughput reference)
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) a[j] = p + j;
#pragma unroll 4
    for (int i = 0; i < n; ++i) {
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = __fmaf_rn(a[j], v.x, __fmul_rn(a[j], v.y));
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) s += a[j];
  } else if (V == 9) {  // two scalar-FFMA chain pairs (4 independent scalar chains)
#pragma unroll 16
    for (int i = 0; i < n; ++i) {
      o0.x = __fmaf_rn(p, v.x, o0.x);
      o0.y = __fmaf_rn(p, v.y, o0.y);
      o1.x = __fmaf_rn(p, v.x, o1.x);
      o1.y = __fmaf_rn(p, v.y, o1.y);
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = o0.x + o0.y + o1.x + o1.y + o2.x + o3.y + s;
}

template <int V>
void run(const char* name, int threads, double per) {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 8);
  const int n = 4096;
  k<V><<<1, threads>>>(out, cyc, n, 1.0001f, 0x3f803f80u);
  k<V><<<1, threads>>>(out, cyc, n, 1.0001f, 0x3f803f80u);
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-58s threads %4d: %7.2f cycles per step (%.2f per instruction)\n", name, threads, (double)h / n, (double)h / n / per);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int th : {32, 128, 256}) {
    run<0>("FFMA2 one dependent chain", th, 1);
    run<1>("FFMA2 two chains", th, 2);
    run<2>("FFMA2 four chains", th, 4);
    run<3>("FFMA scalar pair chain (x,y)", th, 2);
    run<9>("FFMA scalar two pairs", th, 4);
    run<4>("FADD dependent chain", th, 1);
    run<5>("FFMA2 chain + PRMT/LOP3 widening", th, 3);
    run<6>("2 FFMA2 chains + PRMT/LOP3 widening", th, 4);
    run<7>("8 streams FMUL2+FFMA2 (throughput)", th, 16);
    run<8>("16 streams FMUL+FFMA scalar (throughput)", th, 32);
  }
  return 0;
}
