// micro-benchmark (round 2): latency of the building blocks of the stage recursions of
// solve_riccati_kernel on one warp: dependent SHFL.IDX, dependent LDS, and the backward / forward /
// adjoint step bodies.  nvcc -arch=sm_100a -O3 -o ric_micro ric_micro.cu && ./ric_micro
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, long long* clk, int steps, float dt) {
  __shared__ __align__(16) float bw[64 * 144];
  __shared__ __align__(16) float s_s[64 * 6], s_pv[64 * 6], s_xi[65 * 12], s_q[64 * 6];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 144; i += blockDim.x) bw[i] = 1e-3f * (i % 17);
  for (int i = threadIdx.x; i < 64 * 6; i += blockDim.x) { s_s[i] = 0.01f * i; s_pv[i] = 0.f; }
  for (int i = threadIdx.x; i < 65 * 12; i += blockDim.x) s_xi[i] = 0.001f * i;
  __syncthreads();
  if (threadIdx.x >= 32) return;
  long long t0, t1;
  float v = lane;
  // (a) dependent shuffle chain
  t0 = clock64();
  for (int i = 0; i < steps; ++i) v = __shfl_sync(0xffffffffu, v, (lane + 1) & 31) + 1.f;
  t1 = clock64();
  if (lane == 0) clk[0] = (t1 - t0) / steps;
  // (b) dependent LDS chain
  int idx = lane;
  t0 = clock64();
  for (int i = 0; i < steps; ++i) idx = ((int)bw[idx] + idx + 1) & 1023;
  t1 = clock64();
  if (lane == 0) clk[1] = (t1 - t0) / steps;
  v += idx;
  // (c) backward step as in the kernel
  const int l12 = lane < 12 ? lane : 0;
  float pc = 0.f;
  t0 = clock64();
  for (int it = 0; it < steps / 64; ++it)
    for (int kk = 63; kk >= 0; --kk) {
      const float4* b = reinterpret_cast<const float4*>(bw + 144 * kk + 12 * l12);
      const float4 b0 = b[0], b1 = b[1], b2 = b[2];
      const float lt[6] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y};
      const float yr[6] = {b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
      const float2* sp = reinterpret_cast<const float2*>(s_s + 6 * kk);
      const float2 s01 = sp[0], s23 = sp[1], s45 = sp[2];
      float g = yr[0] * s01.x;
      g = fmaf(yr[1], s01.y, g); g = fmaf(yr[2], s23.x, g); g = fmaf(yr[3], s23.y, g);
      g = fmaf(yr[4], s45.x, g); g = fmaf(yr[5], s45.y, g);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int m = 0; m < 6; m += 2) {
        a0 = fmaf(lt[m], __shfl_sync(0xffffffffu, pc, 6 + m), a0);
        a1 = fmaf(lt[m + 1], __shfl_sync(0xffffffffu, pc, 7 + m), a1);
      }
      const float pp = __shfl_sync(0xffffffffu, pc, lane < 6 ? lane : lane - 6);
      if (lane >= 6 && lane < 12) s_pv[6 * kk + lane - 6] = pc;
      pc = (lane < 6 ? pc : fmaf(dt, pp, pc)) - dt * (a0 + a1) + g;
    }
  t1 = clock64();
  if (lane == 0) clk[2] = (t1 - t0) / steps;
  v += pc;
  // (d) forward step
  float xc = 0.f;
  t0 = clock64();
  for (int it = 0; it < steps / 64; ++it)
    for (int kk = 0; kk < 64; ++kk) {
      const float4* fw = reinterpret_cast<const float4*>(bw + 144 * kk + 12 * l12);
      const float4 f0 = fw[0], f1 = fw[1], f2 = fw[2];
      const float fr[12] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w, f2.x, f2.y, f2.z, f2.w};
      const float2* sp = reinterpret_cast<const float2*>(s_s + 6 * kk);
      const float2* pp = reinterpret_cast<const float2*>(s_pv + 6 * kk);
      const float2 s01 = sp[0], s23 = sp[1], s45 = sp[2], p01 = pp[0], p23 = pp[1], p45 = pp[2];
      const float sv[6] = {s01.x, s01.y, s23.x, s23.y, s45.x, s45.y};
      const float pv[6] = {p01.x, p01.y, p23.x, p23.y, p45.x, p45.y};
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int m = 0; m < 6; ++m) {
        const float xa = __shfl_sync(0xffffffffu, xc, m), xb = __shfl_sync(0xffffffffu, xc, 6 + m);
        a0 = fmaf(fr[m], lane < 6 ? sv[m] : xa, a0);
        a1 = fmaf(fr[6 + m], lane < 6 ? -dt * pv[m] : xb, a1);
      }
      const float acc = a0 + a1;
      const float w0 = __shfl_sync(0xffffffffu, acc, lane >= 6 ? lane - 6 : lane);
      const float vel = __shfl_sync(0xffffffffu, xc, lane < 6 ? lane + 6 : lane);
      xc = lane < 6 ? fmaf(dt, vel, xc) : xc + dt * (w0 - acc);
      if (lane < 12) s_xi[12 * (kk + 1) + lane] = xc;
    }
  t1 = clock64();
  if (lane == 0) clk[3] = (t1 - t0) / steps;
  v += xc;
  // (e) adjoint step
  float mp = 0.f, mv = 0.f;
  t0 = clock64();
  for (int it = 0; it < steps / 64; ++it)
    if (lane < 6)
      for (int kk = 64; kk >= 1; --kk) {
        const float pos = s_xi[12 * kk + lane], vel = s_xi[12 * kk + 6 + lane];
        const float mvn = 2.f * vel + fmaf(dt, mp, mv);
        mp = 3.f * pos + mp;
        mv = mvn;
        s_q[6 * (kk - 1) + lane] = dt * mv;
      }
  t1 = clock64();
  if (lane == 0) clk[4] = (t1 - t0) / steps;
  out[lane] = v + mp + mv;
}
int main() {
  float* d; long long* c; cudaMalloc(&d, 128); cudaMalloc(&c, 64);
  const int steps = 64 * 200;
  for (int rep = 0; rep < 2; ++rep) k<<<1, 128>>>(d, c, steps, 0.01f);
  long long h[5]; cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
  printf("cycles per step: shfl-chain %lld  lds-chain %lld  backward %lld  forward %lld  adjoint %lld\n", h[0], h[1], h[2], h[3], h[4]);
  return 0;
}
