// micro-benchmark (round 2): cost of k independent SHFL.IDX from one source register inside a dependent
// chain (does the shuffle unit pipeline?), against an STS -> __syncwarp -> broadcast LDS exchange.
#include <cstdio>
#include <cuda_runtime.h>
template <int K>
__device__ float shfl_step(float v) {
  float acc = 0.f;
#pragma unroll
  for (int m = 0; m < K; ++m) acc += __shfl_sync(0xffffffffu, v, m);
  return acc * 0.03f + 1.f;
}
__global__ void k(float* out, long long* clk, int steps) {
  __shared__ __align__(16) float buf[64];
  const int lane = threadIdx.x & 31;
  float v = lane;
  long long t0, t1;
#define RUN(IDX, ...) t0 = clock64(); for (int i = 0; i < steps; ++i) { __VA_ARGS__; } t1 = clock64(); if (lane == 0) clk[IDX] = (t1 - t0) / steps;
  RUN(0, v = shfl_step<1>(v));
  RUN(1, v = shfl_step<2>(v));
  RUN(2, v = shfl_step<4>(v));
  RUN(3, v = shfl_step<6>(v));
  RUN(4, v = shfl_step<12>(v));
  // smem exchange: 12 lanes publish, everybody reads 12 floats as 3 x LDS.128
  RUN(5, { if (lane < 12) buf[lane] = v; __syncwarp();
           const float4 a = *reinterpret_cast<const float4*>(buf), b = *reinterpret_cast<const float4*>(buf + 4), c = *reinterpret_cast<const float4*>(buf + 8);
           __syncwarp();
           v = (a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w + c.x + c.y + c.z + c.w) * 0.03f + 1.f; });
  // 12-deep FMA chain alone
  RUN(6, { float a = v; 
#pragma unroll
           for (int m = 0; m < 12; ++m) a = fmaf(a, 0.999f, 0.001f); v = a; });
  // xor-shuffle butterfly (2 steps)
  RUN(7, { v += __shfl_xor_sync(0xffffffffu, v, 1); v += __shfl_xor_sync(0xffffffffu, v, 2); v = v * 0.1f + 1.f; });
  out[lane] = v;
}
int main() {
  float* d; long long* c; cudaMalloc(&d, 128); cudaMalloc(&c, 64);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(d, c, 20000);
  long long h[8]; cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
  printf("cycles: shfl x1 %lld  x2 %lld  x4 %lld  x6 %lld  x12 %lld | smem exchange(12) %lld | fma x12 %lld | xor butterfly x2 %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
  return 0;
}
