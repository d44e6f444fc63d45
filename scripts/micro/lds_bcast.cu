// Micro-benchmark: cost of warp-uniform (broadcast) LDS.32/64/128 vs per-lane LDS on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bcast lds_bcast.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int W, bool BCAST>
__global__ void k(float* out, long long* cyc, int iters) {
  __shared__ __align__(16) float s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = (float)i;
  __syncthreads();
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const int lane_off = BCAST ? 0 : (threadIdx.x & 31) * W;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int base = ((it * 16 + u) * 32 * W + lane_off) & 4095 & ~(W - 1);
      if (W == 4) {
        const float4 v = *reinterpret_cast<const float4*>(s + base);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
      } else if (W == 2) {
        const float2 v = *reinterpret_cast<const float2*>(s + base);
        acc[0] += v.x; acc[1] += v.y;
      } else {
        acc[0] += s[base];
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + acc[2] + acc[3];
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int W, bool B>
void run(const char* name, int threads) {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  k<W, B><<<148, threads>>>(out, cyc, iters);
  k<W, B><<<148, threads>>>(out, cyc, iters);
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / (iters * 16.0) / (threads / 32);
  printf("%-28s threads %4d: %.2f cycles per warp-LDS per SM (%.1f B/clk delivered)\n", name, threads, per, 32.0 * W * 4 / per);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  for (int t : {128, 512, 1024}) {
    run<1, true>("LDS.32 broadcast", t); run<1, false>("LDS.32 per-lane", t);
    run<2, true>("LDS.64 broadcast", t); run<2, false>("LDS.64 per-lane", t);
    run<4, true>("LDS.128 broadcast", t); run<4, false>("LDS.128 per-lane", t);
  }
  return 0;
}
