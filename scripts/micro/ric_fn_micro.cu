// micro-benchmark (round 2): the production riccati_sweeps<N> / gram_sweeps<N> (cmpc_riccati.cuh) on one warp
// with synthetic stage matrices: cycles per call, per stage.
// nvcc -I../../mpc-for-dynamic-locomotion-in-the-mit-cheetah-3_b200/csrc -arch=sm_100a -O3 -std=c++17 -o ric_fn_micro ric_fn_micro.cu
#include <cstdio>
#include "cmpc_riccati.cuh"
template <int N>
__global__ void k(float* out, long long* clk, int reps) {
  extern __shared__ __align__(16) float sm[];
  float* s_bw = sm;
  float* s_fw = s_bw + 144 * N;
  float* s_g = s_fw + 144 * N;
  float* s_w0 = s_g + 18 * N;
  float* s_xi = s_w0 + 6 * N;
  float* s_q = s_xi + 12 * (N + 1);
  float* s_s = s_q + 6 * N;
  for (int i = threadIdx.x; i < 288 * N; i += blockDim.x) sm[i] = 1e-3f * ((i * 7) % 13 - 6);
  for (int i = threadIdx.x; i < 18 * N; i += blockDim.x) s_g[i] = 0.01f * (i % 7);
  for (int i = threadIdx.x; i < 6 * N; i += blockDim.x) s_s[i] = 0.01f * (i % 5);
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int lane = threadIdx.x;
  const float qp = lane < 6 ? 2.f : 0.f, qv = lane < 6 ? 3.f : 0.f;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) cmpc::riccati_sweeps<N>(s_bw, s_fw, s_g, s_w0, s_xi, s_q, 0.01f, N, qp, qv, nullptr);
  long long t1 = clock64();
  for (int r = 0; r < reps; ++r) cmpc::gram_sweeps<N>(s_s, s_xi, s_q, 0.01f, N, qp, qv);
  long long t2 = clock64();
  if (lane == 0) { clk[0] = (t1 - t0) / reps; clk[1] = (t2 - t1) / reps; }
  out[lane] = s_q[lane] + s_xi[lane];
}
int main() {
  constexpr int N = 30;
  float* d; long long* c; cudaMalloc(&d, 128); cudaMalloc(&c, 64);
  const size_t smem = (288 * N + 18 * N + 6 * N * 3 + 12 * (N + 1)) * 4;
  cudaFuncSetAttribute(k<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int rep = 0; rep < 2; ++rep) k<N><<<1, 64, smem>>>(d, c, 200);
  long long h[2]; cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
  printf("N=%d: riccati_sweeps %lld cycles per call (%.0f per stage), gram_sweeps %lld (%.0f per stage)  [%s]\n", N, h[0], h[0] / (double)N, h[1], h[1] / (double)N,
         cudaGetErrorString(cudaGetLastError()));
  return 0;
}
