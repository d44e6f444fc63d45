// micro-benchmark / prototype (round 2): blocked symmetric sweep of a 64x64 SPD matrix on the legacy tensor path
// (mma.sync.m16n8k8 TF32, 3xTF32 split), against the question "would tensor cores pay for the factorisation".
//   block step k (8 pivots): L = A[:,k] - E_k, A_kk = Lc Lc', Y = L Lc^-T, A -= Y Y'  (rank-8 update, 64 HMMA-tiles)
//   after 8 steps A = -A^-1 (+2 on the diagonal, removed at the end) - the same sweep the SIMT kernel runs pivot by pivot.
// One problem per 64-thread CTA (2 warps x 32 rows x 64 columns of accumulator fragments), like solve_kernel<10,...>.
// nvcc -arch=sm_100a -O3 -std=c++17 -lineinfo -o mma_sweep mma_sweep.cu
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ float tf32r(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void mma8(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                 "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])));
}
__device__ __forceinline__ void mma8u(float (&c)[4], const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// four 8x4 tiles of 32-bit words (= 8x8 b16 matrices): register i of lane l = word (row l/4, col l%4) of tile i;
// lanes 8i..8i+7 supply the row addresses of tile i
__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], const float* p) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}

// ---- 1. raw HMMA issue rate: NACC independent accumulators per warp ------------------------------------
template <int NACC>
__global__ void mma_rate(float* out, int reps) {
  float c[NACC][4];
  float a[4], b[2];
  for (int i = 0; i < 4; ++i) a[i] = tf32r(1e-3f * (threadIdx.x + i));
  for (int i = 0; i < 2; ++i) b[i] = tf32r(1e-3f * (threadIdx.x + 7 * i));
  for (int q = 0; q < NACC; ++q)
    for (int i = 0; i < 4; ++i) c[q][i] = 0.f;
  for (int r = 0; r < reps; ++r)
#pragma unroll
    for (int q = 0; q < NACC; ++q) mma8(c[q], a, b);
  float s = 0.f;
  for (int q = 0; q < NACC; ++q)
    for (int i = 0; i < 4; ++i) s += c[q][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// ---- 2. the blocked sweep ----------------------------------------------------------------------------
constexpr int NP = 64;
constexpr int YS = 12;   // row stride of Y in words: ldmatrix rows (16 B at 48 B stride) hit 8 distinct bank groups
template <int MINB, bool SPLIT3, bool ROLL = false>
__global__ __launch_bounds__(64, MINB) void sweep_mma(const float* __restrict__ Ain, float* __restrict__ out, int nmat, int nout,
                                                      long long* clk) {
  __shared__ __align__(16) float sL[NP][8];
  __shared__ __align__(16) float sYh[NP][YS];
  __shared__ __align__(16) float sYl[NP][YS];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;
  const float* A = Ain + (size_t)(blockIdx.x % nmat) * NP * NP;
  float acc[2][8][4];        // the NEGATED matrix: the update is acc += Y Y'
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int r0 = 32 * w + 16 * mt + g, c0 = 8 * nt + 2 * t;
      const float2 u = *reinterpret_cast<const float2*>(A + r0 * NP + c0);
      const float2 v = *reinterpret_cast<const float2*>(A + (r0 + 8) * NP + c0);
      acc[mt][nt][0] = -u.x; acc[mt][nt][1] = -u.y; acc[mt][nt][2] = -v.x; acc[mt][nt][3] = -v.y;
    }
  // ldmatrix row addresses of this lane: A operand (16 rows x 8 words -> tiles (rows 0-7 | 8-15) x (words 0-3 | 4-7)),
  // B operand (two 8-row groups x (words 0-3 | 4-7))
  const int a_row = 32 * w + (lane & 7) + 8 * ((lane >> 3) & 1), a_col = 4 * (lane >> 4);
  const int b_row = 8 * (lane >> 4) + (lane & 7), b_col = 4 * ((lane >> 3) & 1);
  const long long t0 = clock64();
#pragma unroll(ROLL ? 1 : 8)
  for (int k = 0; k < 8; ++k) {
    // a. publish column block k (rolled loop: predicated stores instead of a dynamically indexed register array)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
      if (nt == k) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int r0 = 32 * w + 16 * mt + g;
          *reinterpret_cast<float2*>(&sL[r0][2 * t]) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
          *reinterpret_cast<float2*>(&sL[r0 + 8][2 * t]) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
        }
      }
    __syncthreads();
    // b. Cholesky of A_kk (every thread, redundantly: no exchange), own row of Y = L Lc^-T
    float lc[8][8], inv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 lo = *reinterpret_cast<const float4*>(&sL[8 * k + i][0]);
      const float4 hi = *reinterpret_cast<const float4*>(&sL[8 * k + i][4]);
      lc[i][0] = -lo.x; lc[i][1] = -lo.y; lc[i][2] = -lo.z; lc[i][3] = -lo.w;
      lc[i][4] = -hi.x; lc[i][5] = -hi.y; lc[i][6] = -hi.z; lc[i][7] = -hi.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float s = lc[j][j];
#pragma unroll
      for (int m = 0; m < j; ++m) s = fmaf(-lc[j][m], lc[j][m], s);
      inv[j] = rsqrtf(s);
#pragma unroll
      for (int i = j + 1; i < 8; ++i) {
        float v = lc[i][j];
#pragma unroll
        for (int m = 0; m < j; ++m) v = fmaf(-lc[i][m], lc[j][m], v);
        lc[i][j] = v * inv[j];
      }
    }
    float y[8];
    {
      const float4 lo = *reinterpret_cast<const float4*>(&sL[tid][0]);
      const float4 hi = *reinterpret_cast<const float4*>(&sL[tid][4]);
      y[0] = -lo.x; y[1] = -lo.y; y[2] = -lo.z; y[3] = -lo.w; y[4] = -hi.x; y[5] = -hi.y; y[6] = -hi.z; y[7] = -hi.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = y[j] - ((tid - 8 * k) == j ? 1.f : 0.f);      // L = A[:,k] - E_k
#pragma unroll
      for (int m = 0; m < j; ++m) v = fmaf(-y[m], lc[j][m], v);
      y[j] = v * inv[j];
    }
    float yh[8], yl[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      yh[j] = tf32r(y[j]);
      yl[j] = tf32r(y[j] - yh[j]);
    }
    *reinterpret_cast<float4*>(&sYh[tid][0]) = make_float4(yh[0], yh[1], yh[2], yh[3]);
    *reinterpret_cast<float4*>(&sYh[tid][4]) = make_float4(yh[4], yh[5], yh[6], yh[7]);
    if (SPLIT3) {
      *reinterpret_cast<float4*>(&sYl[tid][0]) = make_float4(yl[0], yl[1], yl[2], yl[3]);
      *reinterpret_cast<float4*>(&sYl[tid][4]) = make_float4(yl[4], yl[5], yl[6], yl[7]);
    }
    __syncthreads();
    // c. rank-8 update (-A) += Y Y'
    uint32_t ah[2][4], al[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      ldsm4(ah[mt], &sYh[a_row + 16 * mt][a_col]);
      if (SPLIT3) ldsm4(al[mt], &sYl[a_row + 16 * mt][a_col]);
    }
#pragma unroll
    for (int np = 0; np < 4; ++np) {       // two column tiles per ldmatrix.x4
      uint32_t bh[4], bl[4];
      ldsm4(bh, &sYh[16 * np + b_row][b_col]);
      if (SPLIT3) ldsm4(bl, &sYl[16 * np + b_row][b_col]);
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          if (SPLIT3) {
            mma8u(acc[mt][2 * np + h], al[mt], bh + 2 * h);
            mma8u(acc[mt][2 * np + h], ah[mt], bl + 2 * h);
          }
          mma8u(acc[mt][2 * np + h], ah[mt], bh + 2 * h);
        }
    }
  }
  const long long t1 = clock64();
  if (clk && blockIdx.x == 0 && tid == 0) clk[0] = t1 - t0;
  if ((int)blockIdx.x < nout) {          // acc = A^-1 - 2 I; written out as -A^-1 (what the SIMT sweep leaves in its rows)
    float* O = out + (size_t)blockIdx.x * NP * NP;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const int r0 = 32 * w + 16 * mt + g, c0 = 8 * nt + 2 * t;
        float v0 = acc[mt][nt][0], v1 = acc[mt][nt][1], v2 = acc[mt][nt][2], v3 = acc[mt][nt][3];
        if (r0 == c0) v0 += 2.f;
        if (r0 == c0 + 1) v1 += 2.f;
        if (r0 + 8 == c0) v2 += 2.f;
        if (r0 + 8 == c0 + 1) v3 += 2.f;
        *reinterpret_cast<float2*>(O + r0 * NP + c0) = make_float2(-v0, -v1);
        *reinterpret_cast<float2*>(O + (r0 + 8) * NP + c0) = make_float2(-v2, -v3);
      }
  }
}

// fp64 / fp32 reference inverses on the host (Gauss-Jordan, SPD: no pivoting)
template <typename T>
static void host_inverse(const float* A, std::vector<T>& X) {
  std::vector<T> a(NP * NP);
  for (int i = 0; i < NP * NP; ++i) a[i] = (T)A[i];
  X.assign(NP * NP, (T)0);
  for (int i = 0; i < NP; ++i) X[i * NP + i] = (T)1;
  for (int k = 0; k < NP; ++k) {
    const T d = (T)1 / a[k * NP + k];
    for (int j = 0; j < NP; ++j) { a[k * NP + j] *= d; X[k * NP + j] *= d; }
    for (int i = 0; i < NP; ++i)
      if (i != k) {
        const T f = a[i * NP + k];
        for (int j = 0; j < NP; ++j) { a[i * NP + j] -= f * a[k * NP + j]; X[i * NP + j] -= f * X[k * NP + j]; }
      }
  }
}

int main() {
  // 1. HMMA rate
  {
    float* d; cudaMalloc(&d, 148 * 8 * 64 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 20000;
    for (int warps_per_sm : {4, 8, 16}) {
      const int blocks = 148 * warps_per_sm / 2;
      mma_rate<8><<<blocks, 64>>>(d, 10);
      cudaEventRecord(e0);
      mma_rate<8><<<blocks, 64>>>(d, reps);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double n = (double)blocks * 2 * reps * 8;
      printf("mma.sync m16n8k8 tf32: %2d warps/SM, 8 accumulators each: %.3f HMMA/ns per GPU = %.2f per SM-cycle at 1.965 GHz (%.1f TFLOP/s)\n",
             warps_per_sm, n / (ms * 1e6), n / (ms * 1e6) / 148 / 1.965, n * 2048 / (ms * 1e9));
    }
  }
  // 2. blocked sweep: accuracy and time
  const int nmat = 64;
  std::vector<float> A((size_t)nmat * NP * NP);
  srand(1);
  for (int b = 0; b < nmat; ++b) {   // SPD with unit diagonal, condition number ~1e3..1e4 (like the Jacobi-scaled P)
    std::vector<double> R(NP * 12), S(NP * NP, 0.0);
    for (auto& v : R) v = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < NP; ++i)
      for (int j = 0; j < NP; ++j) {
        double s = (i == j) ? 2e-3 : 0.0;
        for (int m = 0; m < 12; ++m) s += R[i * 12 + m] * R[j * 12 + m];
        S[i * NP + j] = s;
      }
    for (int i = 0; i < NP; ++i)
      for (int j = 0; j < NP; ++j) A[(size_t)b * NP * NP + i * NP + j] = (float)(S[i * NP + j] / std::sqrt(S[i * NP + i] * S[j * NP + j]));
    for (int i = 60; i < NP; ++i)    // padding rows like the 60 -> 64 pad of the real kernel
      for (int j = 0; j < NP; ++j) A[(size_t)b * NP * NP + i * NP + j] = A[(size_t)b * NP * NP + j * NP + i] = (i == j) ? 1.f : 0.f;
  }
  float *dA, *dO; long long* dclk;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dO, A.size() * 4); cudaMalloc(&dclk, 64);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  std::vector<float> O(A.size());
  for (int split = 1; split >= 0; --split) {
    if (split) sweep_mma<8, true><<<nmat, 64>>>(dA, dO, nmat, nmat, dclk);
    else sweep_mma<8, false><<<nmat, 64>>>(dA, dO, nmat, nmat, dclk);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, worst32 = 0, worst_id = 0;
    for (int b = 0; b < 8; ++b) {
      std::vector<double> X; std::vector<float> X32;
      host_inverse<double>(&A[(size_t)b * NP * NP], X);
      host_inverse<float>(&A[(size_t)b * NP * NP], X32);
      double mx = 0, e = 0, e32 = 0;
      for (int i = 0; i < NP * NP; ++i) {
        mx = std::max(mx, std::fabs(X[i]));
        e = std::max(e, std::fabs(-(double)O[(size_t)b * NP * NP + i] - X[i]));
        e32 = std::max(e32, std::fabs((double)X32[i] - X[i]));
      }
      // residual |A X - I| with the device result
      double rid = 0;
      for (int i = 0; i < NP; ++i)
        for (int j = 0; j < NP; ++j) {
          double s = 0;
          for (int m = 0; m < NP; ++m) s += (double)A[(size_t)b * NP * NP + i * NP + m] * -(double)O[(size_t)b * NP * NP + m * NP + j];
          rid = std::max(rid, std::fabs(s - (i == j)));
        }
      worst = std::max(worst, e / mx); worst32 = std::max(worst32, e32 / mx); worst_id = std::max(worst_id, rid);
    }
    printf("%s: max |X - X64| / max|X64| = %.3e (host fp32 Gauss-Jordan: %.3e), max |A X - I| = %.3e  [%s]\n", split ? "3xTF32" : "1xTF32", worst, worst32,
           worst_id, cudaGetErrorString(cudaGetLastError()));
  }
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int roll = 0; roll < 2; ++roll)
  for (int B : {1, 148, 592, 1184, 2368, 4736}) {
    auto kern = roll ? sweep_mma<8, true, true> : sweep_mma<8, true, false>;
    kern<<<B, 64>>>(dA, dO, nmat, 0, dclk);
    cudaEventRecord(e0);
    for (int r = 0; r < 10; ++r) kern<<<B, 64>>>(dA, dO, nmat, 0, dclk);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, dclk, 8, cudaMemcpyDeviceToHost);
    printf("3xTF32 sweep (%s): B = %4d (%4.1f CTAs/SM): %.2f us per launch, CTA 0 sweep = %lld cycles (%lld per 8-pivot block step)\n",
           roll ? "rolled" : "unrolled", B, B / 148.0, ms * 100, c, c / 8);
  }
  {  // the rolled kernel computes the same thing
    sweep_mma<8, true, true><<<nmat, 64>>>(dA, dO, nmat, nmat, dclk);
    std::vector<float> O2(A.size());
    cudaMemcpy(O2.data(), dO, O2.size() * 4, cudaMemcpyDeviceToHost);
    double d = 0;
    for (size_t i = 0; i < O.size(); ++i) d = std::max(d, (double)std::fabs(O2[i]));
    sweep_mma<8, true, false><<<nmat, 64>>>(dA, dO, nmat, nmat, dclk);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    double e = 0;
    for (size_t i = 0; i < O.size(); ++i) e = std::max(e, (double)std::fabs(O2[i] - O[i]));
    printf("rolled vs unrolled: max |diff| = %.3e (max |X| = %.3e)\n", e, d);
  }
  return 0;
}
