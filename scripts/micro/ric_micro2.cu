// micro-benchmark (round 2): the software-pipelined stage recursions of solve_riccati_kernel on one warp,
// shuffle exchange against shared-memory exchange of the 12-vector.
#include <cstdio>
#include <cuda_runtime.h>
#define NS 64
__global__ void k(float* out, long long* clk, int reps, float dt) {
  __shared__ __align__(16) float mat[NS * 144];
  __shared__ __align__(16) float s_g[NS * 12], s_w0[NS * 6], s_xi[(NS + 1) * 12], s_pv[NS * 6], xb[2][16];
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NS * 144; i += blockDim.x) mat[i] = 1e-3f * (i % 17);
  for (int i = threadIdx.x; i < NS * 12; i += blockDim.x) s_g[i] = 0.01f * (i % 7);
  for (int i = threadIdx.x; i < NS * 6; i += blockDim.x) s_w0[i] = 0.01f * (i % 5);
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int l12 = lane < 12 ? lane : 0;
  const int partner = lane < 6 ? lane : (lane < 12 ? lane - 6 : 0);
  long long t0, t1;
  float acc_out = 0.f;
  // (0) backward, shuffles (as in the kernel)
  {
    const float cpp = (lane >= 6 && lane < 12) ? dt : 0.f;
    float pc = 0.f;
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const float* row = mat + 144 * (NS - 1) + 12 * l12;
      float4 c0 = *reinterpret_cast<const float4*>(row);
      float2 c1 = *reinterpret_cast<const float2*>(row + 4);
      float g = s_g[12 * (NS - 1) + l12];
      for (int k = NS - 1; k >= 0; --k) {
        const int kn = k > 0 ? k - 1 : 0;
        const float* nrow = mat + 144 * kn + 12 * l12;
        const float4 n0 = *reinterpret_cast<const float4*>(nrow);
        const float2 n1 = *reinterpret_cast<const float2*>(nrow + 4);
        const float gn = s_g[12 * kn + l12];
        const float y0 = __shfl_sync(0xffffffffu, pc, 6), y1 = __shfl_sync(0xffffffffu, pc, 7),
                    y2 = __shfl_sync(0xffffffffu, pc, 8), y3 = __shfl_sync(0xffffffffu, pc, 9),
                    y4 = __shfl_sync(0xffffffffu, pc, 10), y5 = __shfl_sync(0xffffffffu, pc, 11);
        const float pp = __shfl_sync(0xffffffffu, pc, partner);
        if (lane >= 6 && lane < 12) s_pv[6 * k + lane - 6] = pc;
        const float a0 = fmaf(c0.z, y2, fmaf(c0.y, y1, c0.x * y0));
        const float a1 = fmaf(c1.y, y5, fmaf(c1.x, y4, c0.w * y3));
        pc = fmaf(-dt, a0 + a1, fmaf(cpp, pp, pc) + g);
        c0 = n0; c1 = n1; g = gn;
      }
    }
    t1 = clock64();
    if (lane == 0) clk[0] = (t1 - t0) / (reps * NS);
    acc_out += pc;
  }
  // (1) forward, shuffles
  {
    const bool vlane = lane >= 6 && lane < 12;
    float xc = 0.f;
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const float* row = mat + 12 * l12;
      float4 c0 = *reinterpret_cast<const float4*>(row), c1 = *reinterpret_cast<const float4*>(row + 4), c2 = *reinterpret_cast<const float4*>(row + 8);
      float w0 = vlane ? s_w0[lane - 6] : 0.f;
      for (int k = 0; k < NS; ++k) {
        const int kn = k + 1 < NS ? k + 1 : k;
        const float* nrow = mat + 144 * kn + 12 * l12;
        const float4 n0 = *reinterpret_cast<const float4*>(nrow), n1 = *reinterpret_cast<const float4*>(nrow + 4), n2 = *reinterpret_cast<const float4*>(nrow + 8);
        const float wn = vlane ? s_w0[6 * kn + lane - 6] : 0.f;
        float xv[12];
#pragma unroll
        for (int m = 0; m < 12; ++m) xv[m] = __shfl_sync(0xffffffffu, xc, m);
        const float vel = __shfl_sync(0xffffffffu, xc, lane < 6 ? lane + 6 : lane);
        const float a0 = fmaf(c0.w, xv[3], fmaf(c0.z, xv[2], fmaf(c0.y, xv[1], c0.x * xv[0])));
        const float a1 = fmaf(c1.w, xv[7], fmaf(c1.z, xv[6], fmaf(c1.y, xv[5], c1.x * xv[4])));
        const float a2 = fmaf(c2.w, xv[11], fmaf(c2.z, xv[10], fmaf(c2.y, xv[9], c2.x * xv[8])));
        const float inc = vlane ? w0 - ((a0 + a1) + a2) : vel;
        xc = fmaf(dt, inc, xc);
        if (lane < 12) s_xi[12 * (k + 1) + lane] = xc;
        c0 = n0; c1 = n1; c2 = n2; w0 = wn;
      }
    }
    t1 = clock64();
    if (lane == 0) clk[1] = (t1 - t0) / (reps * NS);
    acc_out += xc;
  }
  // (2) forward, shared-memory exchange: the state is written to s_xi anyway; read it back as 3 x LDS.128
  {
    const bool vlane = lane >= 6 && lane < 12;
    float xc = 0.f;
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const float* row = mat + 12 * l12;
      float4 c0 = *reinterpret_cast<const float4*>(row), c1 = *reinterpret_cast<const float4*>(row + 4), c2 = *reinterpret_cast<const float4*>(row + 8);
      float w0 = vlane ? s_w0[lane - 6] : 0.f;
      if (lane < 12) s_xi[lane] = 0.f;
      __syncwarp();
      for (int k = 0; k < NS; ++k) {
        const int kn = k + 1 < NS ? k + 1 : k;
        const float* nrow = mat + 144 * kn + 12 * l12;
        const float4 n0 = *reinterpret_cast<const float4*>(nrow), n1 = *reinterpret_cast<const float4*>(nrow + 4), n2 = *reinterpret_cast<const float4*>(nrow + 8);
        const float wn = vlane ? s_w0[6 * kn + lane - 6] : 0.f;
        const float4 x0 = *reinterpret_cast<const float4*>(s_xi + 12 * k), x1 = *reinterpret_cast<const float4*>(s_xi + 12 * k + 4),
                     x2 = *reinterpret_cast<const float4*>(s_xi + 12 * k + 8);
        const float vel = s_xi[12 * k + (lane < 6 ? lane + 6 : l12)];
        const float a0 = fmaf(c0.w, x0.w, fmaf(c0.z, x0.z, fmaf(c0.y, x0.y, c0.x * x0.x)));
        const float a1 = fmaf(c1.w, x1.w, fmaf(c1.z, x1.z, fmaf(c1.y, x1.y, c1.x * x1.x)));
        const float a2 = fmaf(c2.w, x2.w, fmaf(c2.z, x2.z, fmaf(c2.y, x2.y, c2.x * x2.x)));
        const float inc = vlane ? w0 - ((a0 + a1) + a2) : vel;
        xc = fmaf(dt, inc, xc);
        if (lane < 12) s_xi[12 * (k + 1) + lane] = xc;
        __syncwarp();
        c0 = n0; c1 = n1; c2 = n2; w0 = wn;
      }
    }
    t1 = clock64();
    if (lane == 0) clk[2] = (t1 - t0) / (reps * NS);
    acc_out += xc;
  }
  // (3) backward, shared-memory exchange through a 2-slot ring
  {
    const float cpp = (lane >= 6 && lane < 12) ? dt : 0.f;
    float pc = 0.f;
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const float* row = mat + 144 * (NS - 1) + 12 * l12;
      float4 c0 = *reinterpret_cast<const float4*>(row);
      float2 c1 = *reinterpret_cast<const float2*>(row + 4);
      float g = s_g[12 * (NS - 1) + l12];
      if (lane < 12) xb[1][lane] = pc;
      __syncwarp();
      for (int k = NS - 1; k >= 0; --k) {
        const int kn = k > 0 ? k - 1 : 0;
        const float* nrow = mat + 144 * kn + 12 * l12;
        const float4 n0 = *reinterpret_cast<const float4*>(nrow);
        const float2 n1 = *reinterpret_cast<const float2*>(nrow + 4);
        const float gn = s_g[12 * kn + l12];
        const float* xx = xb[(k + 1) & 1];
        const float2 ya = *reinterpret_cast<const float2*>(xx + 6), yb = *reinterpret_cast<const float2*>(xx + 8), yc = *reinterpret_cast<const float2*>(xx + 10);
        const float pp = xx[partner];
        const float a0 = fmaf(c0.z, yb.x, fmaf(c0.y, ya.y, c0.x * ya.x));
        const float a1 = fmaf(c1.y, yc.y, fmaf(c1.x, yc.x, c0.w * yb.y));
        pc = fmaf(-dt, a0 + a1, fmaf(cpp, pp, pc) + g);
        if (lane < 12) xb[k & 1][lane] = pc;
        __syncwarp();
        c0 = n0; c1 = n1; g = gn;
      }
    }
    t1 = clock64();
    if (lane == 0) clk[3] = (t1 - t0) / (reps * NS);
    acc_out += pc;
  }
  out[lane] = acc_out;
}
int main() {
  float* d; long long* c; cudaMalloc(&d, 128); cudaMalloc(&c, 64);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 128>>>(d, c, 200, 0.01f);
  long long h[4]; cudaMemcpy(h, c, sizeof h, cudaMemcpyDeviceToHost);
  printf("cycles per stage: backward(shfl) %lld  forward(shfl) %lld  forward(smem) %lld  backward(smem) %lld\n", h[0], h[1], h[2], h[3]);
  return 0;
}
