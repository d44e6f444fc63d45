// cmpc_cluster.cuh - long-horizon variant of the solve kernel: one THREAD-BLOCK CLUSTER per
// problem (sm_90+/sm_100a clusters + distributed shared memory).
//
// The 6N x 6N wrench matrix P lives in registers; at N = 60 it is
// 518 KB, more than one SM's register file (256 KB) or shared memory (227 KB).  CL CTAs of a
// cluster each own NL = N/CL consecutive stages: their 4 NL legs and the 6 NL rows of P
// (all 6N columns, SPLIT threads per row).  What crosses CTAs goes through distributed
// shared memory, pushed by the producer:
//   * the pivot row of every sweep step (the owning CTA stores it into all CTAs' buffers),
//   * the wrench-space right-hand side s = G D^-1 b of every ADMM iteration (each CTA
//     contributes its stages), together with the residual statistics of that iteration,
//   * the Jacobi scaling vector and the wrench sequence for refresh / X output,
// followed by one cluster barrier.  All all-gather buffers are double-buffered on a running
// counter: a CTA is never more than one cluster barrier ahead of the slowest one.
// Same algorithm, same iterates as solve_kernel (the per-thread code is the same), so the
// same oracle tests apply.
#pragma once
#include <cooperative_groups.h>

#include "cmpc_kernels.cuh"

namespace cmpc {
namespace cg = cooperative_groups;

template <int NL, int CL, int SPLIT>
struct CGeo {
  static constexpr int N = NL * CL;
  static constexpr int NWL = 6 * NL;                                 // local rows
  static constexpr int NWT = 6 * N;                                  // columns
  static constexpr int NWP = ((NWT + 4 * SPLIT - 1) / (4 * SPLIT)) * (4 * SPLIT);
  static constexpr int COLS = NWP / SPLIT;
  static constexpr int NLEG = 4 * NL;                                // local legs
  static constexpr int ROWT = NWL * SPLIT;
  static constexpr int TMAX = ROWT > NLEG ? ROWT : NLEG;
  static constexpr int THREADS = ((TMAX + 31) / 32) * 32;
  static constexpr int LWARPS = (NLEG + 31) / 32;
  static constexpr int NX = 13 * (N + 1);
};

template <int NL, int CL, int SPLIT, int MINB>
__global__ void __launch_bounds__(CGeo<NL, CL, SPLIT>::THREADS, MINB)
solve_cluster_kernel(const SolveParams p) {
  using G_ = CGeo<NL, CL, SPLIT>;
  constexpr int N = G_::N, NWL = G_::NWL, NWT = G_::NWT, NWP = G_::NWP, COLS = G_::COLS;
  constexpr int NLEG = G_::NLEG, THREADS = G_::THREADS, LWARPS = G_::LWARPS, NX = G_::NX;

  __shared__ __align__(16) float s_x0[16];
  __shared__ __align__(16) float s_xd[NX + 3];
  __shared__ __align__(16) float s_G[NLEG][12];
  constexpr int NB = NL;                                // pivots per cluster barrier (divides 6 NL)
  __shared__ __align__(16) float s_row[2][NB][NWP + 4]; // pivot rows of one block, pushed by the owner CTA
  __shared__ __align__(16) float s_s[2][NWP];           // all-gathered wrench vectors
  __shared__ __align__(16) float s_q[SPLIT][NWL];       // q = P^-1 s (local rows), partial sums
  __shared__ __align__(16) float s_v[NWL];
  __shared__ __align__(16) float s_h[NWL];
  __shared__ __align__(16) float s_S[NWP];              // Jacobi scaling, all-gathered
  __shared__ float s_red[2][CL][LWARPS][8];             // all-gathered residual statistics
  __shared__ int s_mask[NL];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int prob = (int)blockIdx.x / CL;
  // the whole cluster takes the same exit (p.B is uniform)
  if (prob >= p.B) return;
  const int b = p.order ? p.order[prob] : prob;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const bool leg_warp = warp < LWARPS;
  const int slot = p.slot0 + b;
  const int j0 = rank * NL;                   // first global stage of this CTA
  const int g0 = rank * NWL;                  // first global row of this CTA

  // remote views of the all-gather buffers
  float* r_row[CL];
  float* r_s[CL];
  float* r_S[CL];
  float* r_red[CL];
#pragma unroll
  for (int c = 0; c < CL; ++c) {
    r_row[c] = cluster.map_shared_rank(&s_row[0][0][0], c);
    r_s[c] = cluster.map_shared_rank(&s_s[0][0], c);
    r_S[c] = cluster.map_shared_rank(&s_S[0], c);
    r_red[c] = cluster.map_shared_rank(&s_red[0][0][0][0], c);
  }

  // ---- phase 0: stage the problem record -------------------------------------------------
  for (int i = tid; i < 13; i += THREADS) s_x0[i] = __ldg(p.x0 + (size_t)b * 13 + i);
  for (int i = tid; i < NX; i += THREADS) s_xd[i] = __ldg(p.x_des + (size_t)b * NX + i);
  for (int i = tid; i < NL; i += THREADS) s_mask[i] = (int)__ldg(p.mask + (size_t)b * N + j0 + i);
  for (int i = tid; i < 2 * NWP; i += THREADS) (&s_s[0][0])[i] = 0.f;
  for (int i = tid; i < NWL; i += THREADS) { s_v[i] = 0.f; s_h[i] = 0.f; }
  const float mu = __ldg(p.mu + b);
  const bool is_leg = tid < NLEG;
  const int gleg = rank * NLEG + tid;         // global leg-stage index of a leg thread
  float r_in[3] = {0.f, 0.f, 0.f};
  float wx_in[3] = {0.f, 0.f, 0.f};
  float wy_in[3] = {0.f, 0.f, 0.f};
  const bool warm = p.warm_mode != 0 && p.warm_valid[slot] != 0;
  if (is_leg) {
    const float* rp = p.r + ((size_t)b * 4 * N + gleg) * 3;
    r_in[0] = __ldg(rp); r_in[1] = __ldg(rp + 1); r_in[2] = __ldg(rp + 2);
    if (warm) {
      const float* wx = p.warm_x + ((size_t)slot * 4 * N + gleg) * 3;
      wx_in[0] = wx[0]; wx_in[1] = wx[1]; wx_in[2] = wx[2];
      if (p.warm_mode == 2) {
        const float* wy = p.warm_y + ((size_t)slot * 4 * N + gleg) * 3;
        wy_in[0] = wy[0]; wy_in[1] = wy[1]; wy_in[2] = wy[2];
      }
    }
  }
  // no CTA may push into a peer's shared memory before that peer has started
  cluster.sync();

  float sn, cs;
  sincosf(s_x0[2], &sn, &cs);
  const float im = p.inv_mass;
  const float alpha = p.alpha;
  float rho = p.rho;
  float rho_inv = 1.f / rho;

  // ---- phase 1: leg geometry, linear term ------------------------------------------------
  const int lj = tid >> 2, ll = tid & 3;                 // local stage, leg
  bool stance = false;
  float Gh[3][3];
  float dinv = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int k = 0; k < 3; ++k) Gh[a][k] = 0.f;
  if (is_leg) {
    stance = (s_mask[lj] >> ll) & 1;
    leg_map(cs, sn, p.ib, r_in[0], r_in[1], r_in[2], Gh);
    if (!stance) {
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int k = 0; k < 3; ++k) Gh[a][k] = 0.f;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int k = 0; k < 3; ++k) s_G[tid][3 * a + k] = Gh[a][k];
    s_G[tid][11] = stance ? 1.f : 0.f;
  }
  const bool is_row = tid < G_::ROWT;
  const int rs = SPLIT == 1 ? 0 : tid / NWL;             // slice
  const int ri = tid - rs * NWL;                         // local row
  const int gi = g0 + ri;                                // global row
  const int rjl = ri / 6, ra = ri % 6;                   // local stage, axis
  const int rj = j0 + rjl;                               // global stage
  if (is_row && rs == 0)
    s_h[ri] = wrench_linear_term<N>(rj, ra, s_x0, s_xd, cs, sn, p.w, p.dt, p.n_eff > 0 ? p.n_eff : N);
  __syncthreads();

  int ag = 0;        // all-gather counter: buffer parity of s_s / s_red

  // ---- factorisation (re-runnable for adaptive rho) ----------------------------------------
  float row[COLS];
  auto factorize = [&]() {
    dinv = stance ? 1.f / (p.sigma + 2.f * p.r_weight + rho) : 0.f;
    if (is_leg) s_G[tid][9] = dinv;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < COLS; ++c) row[c] = 0.f;
    float E[6];
#pragma unroll
    for (int a2 = 0; a2 < 6; ++a2) E[a2] = 0.f;
    const float* mi = p.Minv + ((size_t)(is_row ? ra : 0) * N + (is_row ? rj : 0)) * N;
    float sc_i = 1.f;
    if (is_row) {
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        const float* g = s_G[4 * rjl + l];
        const float d = g[9];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float mine = (ra < 3 ? g[3 * ra + c] : (ra - 3 == c ? im : 0.f)) * d;
#pragma unroll
          for (int a2 = 0; a2 < 3; ++a2) E[a2] += mine * g[3 * a2 + c];
          E[3 + c] += mine * im;
        }
      }
      sc_i = rsqrtf(__ldg(mi + rj) + E[ra]);
      if (rs == 0) {
#pragma unroll
        for (int c = 0; c < CL; ++c) r_S[c][gi] = sc_i;      // all-gather of the scaling
      }
    }
    cluster.sync();
    if (is_row) {
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const int col = rs * COLS + c;
        const int j2 = col / 6, a2 = col % 6;
        float v = 0.f;
        if (col < NWT) {
          if (a2 == ra) v = __ldg(mi + j2);
          if (j2 == rj) {
            float e = 0.f;
#pragma unroll
            for (int q = 0; q < 6; ++q) e = (a2 == q) ? E[q] : e;
            v += e;
          }
          v *= sc_i * s_S[col];
        }
        row[c] = v;
      }
    }
    // symmetric Gauss-Jordan sweep over all 6N pivots (see solve_kernel); the owner CTA pushes
    // the pivot row into every CTA's buffer, one cluster barrier per pivot
    float diag = 1.f, rdiag = 1.f;
    if (is_row) {
      diag = (__ldg(mi + rj) + E[ra]) * sc_i * sc_i;
      rdiag = __fdividef(1.f, diag);
    }
    // Blocked over NB pivots: the CTA that owns a block runs its NB steps with local barriers
    // only (pushing each pivot row to every peer as it goes), one cluster barrier publishes the
    // block, then the peers apply the NB pivots back to back from their local copies (each
    // thread only touches its own row slice, so they need no barrier in between).  Buffers
    // alternate with the block parity: nobody can be more than one block ahead.
    auto apply_pivot = [&](const float* buf, int k) {
      const float d = buf[NWP];
      const float m = buf[gi];
      const bool own = gi == k;
      const float nf = own ? d - 1.f : -m * d;
      diag = own ? -d : fmaf(nf, m, diag);
      rdiag = __fdividef(1.f, diag);
      const float* pr = buf + rs * COLS;
#pragma unroll
      for (int c = 0; c < COLS; c += 4) {
        const float4 pv = *reinterpret_cast<const float4*>(pr + c);
        row[c] = fmaf(nf, pv.x, row[c]);
        row[c + 1] = fmaf(nf, pv.y, row[c + 1]);
        row[c + 2] = fmaf(nf, pv.z, row[c + 2]);
        row[c + 3] = fmaf(nf, pv.w, row[c + 3]);
      }
    };
    for (int blk = 0; blk < NWT / NB; ++blk) {
      const int k0 = blk * NB;
      const int boff = (blk & 1) * NB * (NWP + 4);
      if (rank == k0 / NWL) {
        for (int j = 0; j < NB; ++j) {
          const int k = k0 + j;
          float* lbuf = &s_row[0][0][0] + boff + j * (NWP + 4);
          if (is_row && gi == k) {
#pragma unroll
            for (int c = 0; c < COLS; c += 4)
              *reinterpret_cast<float4*>(lbuf + rs * COLS + c) =
                  make_float4(row[c], row[c + 1], row[c + 2], row[c + 3]);
            if (k / COLS == rs) {
              lbuf[k] = diag - 1.f;
              lbuf[NWP] = rdiag;
            }
          }
          __syncthreads();
          {   // the whole CTA pushes the finished pivot row to the peers (fire and forget; the
              // block's cluster barrier publishes it), instead of 6 threads x 15 x CL stores
            constexpr int NF4 = (NWP + 4) / 4;
            const float4* src = reinterpret_cast<const float4*>(lbuf);
            for (int e = tid; e < NF4 * (CL - 1); e += THREADS) {
              const int cc = e / NF4, off = e - cc * NF4;
              int dst = rank + 1 + cc;
              if (dst >= CL) dst -= CL;
              reinterpret_cast<float4*>(r_row[dst] + boff + j * (NWP + 4))[off] = src[off];
            }
          }
          if (is_row) apply_pivot(&s_row[0][0][0] + boff + j * (NWP + 4), k);
        }
        cluster.sync();
      } else {
        cluster.sync();
        if (is_row) {
          for (int j = 0; j < NB; ++j) apply_pivot(&s_row[0][0][0] + boff + j * (NWP + 4), k0 + j);
        }
      }
    }
    if (is_row) {
#pragma unroll
      for (int c = 0; c < COLS; ++c) {
        const int col = rs * COLS + c;
        const float v = col == gi ? diag : row[c];
        row[c] = v * sc_i * (col < NWT ? s_S[col] : 0.f);
      }
    }
    // the last pivot buffer may still be read by a slow CTA while a fast one starts the next
    // all-gather into s_s: different buffers, nothing to wait for here
    __syncthreads();
  };
  factorize();

  // ---- initial iterate ------------------------------------------------------------------------
  const float fmin = p.f_min, fmax = p.f_max;
  const float inv1 = 1.f / (1.f + mu * mu), inv2 = 1.f / (1.f + 2.f * mu * mu);
  auto project = [&](float wx, float wy, float wz, float& zx, float& zy, float& zz) {
    const float ax = fabsf(wx), ay = fabsf(wy);
    const float big = fmaxf(ax, ay), small = fminf(ax, ay);
    const float f2 = (wz + mu * big) * inv1;
    const float f1 = (wz + mu * (ax + ay)) * inv2;
    float fz = (mu * wz >= big) ? wz : ((mu * f2 >= small) ? f2 : f1);
    fz = fminf(fmaxf(fz, fmin), fmax);
    const float lim = mu * fz;
    zx = fminf(fmaxf(wx, -lim), lim);
    zy = fminf(fmaxf(wy, -lim), lim);
    zz = fz;
  };
  float x[3] = {0.f, 0.f, 0.f}, y[3] = {0.f, 0.f, 0.f}, z[3] = {0.f, 0.f, 0.f};
  float gl[3] = {0.f, 0.f, 0.f};
  float hj[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float vh[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (is_leg && stance) {
#pragma unroll
    for (int a = 0; a < 6; ++a) hj[a] = s_h[6 * lj + a];
#pragma unroll
    for (int k = 0; k < 3; ++k)
      gl[k] = Gh[0][k] * hj[0] + Gh[1][k] * hj[1] + Gh[2][k] * hj[2] + im * hj[3 + k];
    if (warm) {
      x[0] = wx_in[0]; x[1] = wx_in[1]; x[2] = wx_in[2];
      y[0] = wy_in[0]; y[1] = wy_in[1]; y[2] = wy_in[2];
    }
    project(x[0], x[1], x[2], z[0], z[1], z[2]);
  }
#pragma unroll
  for (int a = 0; a < 6; ++a) vh[a] = hj[a];

  // push this CTA's part of a wrench-space vector G * v (v per leg) into every CTA
  auto gather_wrench = [&](const float v3[3]) {
    if (leg_warp) {
      float wv[6];
#pragma unroll
      for (int a = 0; a < 3; ++a)
        wv[a] = quad_sum(Gh[a][0] * v3[0] + Gh[a][1] * v3[1] + Gh[a][2] * v3[2]);
#pragma unroll
      for (int k = 0; k < 3; ++k) wv[3 + k] = quad_sum(v3[k] * im);
      if (is_leg && ll < 3) {
        const int off = (ag & 1) * NWP + 6 * (j0 + lj) + 2 * ll;
        const float2 val = pick_pair(wv, ll);
#pragma unroll
        for (int c = 0; c < CL; ++c) *reinterpret_cast<float2*>(r_s[c] + off) = val;
      }
    }
  };

  auto refresh_gradient = [&]() {
    gather_wrench(x);
    cluster.sync();
    const float* sw = &s_s[ag & 1][0];
    ++ag;
    if (is_row && rs == 0) {
      const float* mg = p.Mg + ((size_t)ra * N + rj) * N;
      float acc = 0.f;
      for (int j2 = 0; j2 < N; ++j2) acc = fmaf(__ldg(mg + j2), sw[6 * j2 + ra], acc);
      s_v[ri] = acc;
    }
    __syncthreads();
    if (is_leg && stance) {
#pragma unroll
      for (int a = 0; a < 6; ++a) vh[a] = s_v[6 * lj + a] + hj[a];
    }
    __syncthreads();
  };
  if (warm) refresh_gradient();

  // ||G' h||_inf over the whole cluster
  float ng = 0.f;
  {
    float m = fmaxf(fabsf(gl[0]), fmaxf(fabsf(gl[1]), fabsf(gl[2])));
    if (leg_warp) {
      m = warp_max_nonneg(m);
      if (lane == 0) {
        const int off = (((ag & 1) * CL + rank) * LWARPS + warp) * 8;
#pragma unroll
        for (int c = 0; c < CL; ++c) r_red[c][off] = m;
      }
    }
    cluster.sync();
#pragma unroll
    for (int c = 0; c < CL; ++c)
#pragma unroll
      for (int wv = 0; wv < LWARPS; ++wv) ng = fmaxf(ng, s_red[ag & 1][c][wv][0]);
    ++ag;
  }

  // ---- ADMM -------------------------------------------------------------------------------------
  int it = 0, status = 0;
  float pri = 0.f, dua = 0.f;
  const float two_rw = 2.f * p.r_weight;
  int next_chk = 0;
  int next_ref = p.refresh_every > 0 ? p.refresh_every : -1;
  int next_adp = p.adaptive_rho_interval > 0 ? p.adaptive_rho_interval : -1;
  for (;;) {
    if (it == next_ref) {
      next_ref += p.refresh_every;
      refresh_gradient();
    }
    const bool adapt_now = it == next_adp;
    if (adapt_now) next_adp += p.adaptive_rho_interval;
    const bool chk = it == next_chk || it >= p.max_iter || adapt_now;
    if (it == next_chk) next_chk += p.check_every;
    float t[3] = {0.f, 0.f, 0.f};
    if (leg_warp) {
      float gr[3], rp[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        gr[k] = Gh[0][k] * vh[0] + Gh[1][k] * vh[1] + Gh[2][k] * vh[2] + im * vh[3 + k] + two_rw * x[k];
        rp[k] = x[k] - z[k];
        t[k] = -dinv * (gr[k] + y[k] + rho * rp[k]);
      }
      gather_wrench(t);
      if (chk) {
        float st_[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        float sum = 0.f;
        if (stance) {
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const float rd = gr[k] + y[k];
            st_[0] = fmaxf(st_[0], fabsf(rp[k]));
            st_[1] = fmaxf(st_[1], fabsf(rd));
            st_[2] = fmaxf(st_[2], fmaxf(fabsf(x[k]), fabsf(z[k])));
            st_[3] = fmaxf(st_[3], fmaxf(fabsf(gr[k] - gl[k]), fabsf(y[k])));
            sum += x[k] + rd;
          }
        }
        st_[4] = fabsf(sum * 0.f);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const float m = warp_max_nonneg(st_[k]);
          if (lane == 0) {
            const int off = (((ag & 1) * CL + rank) * LWARPS + warp) * 8 + k;
#pragma unroll
            for (int c = 0; c < CL; ++c) r_red[c][off] = m;
          }
        }
      }
    }
    cluster.sync();
    const int par = ag & 1;
    ++ag;
    if (chk) {
      float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
      for (int c = 0; c < CL; ++c)
#pragma unroll
        for (int wv = 0; wv < LWARPS; ++wv) {
          m0 = fmaxf(m0, s_red[par][c][wv][0]);
          m1 = fmaxf(m1, s_red[par][c][wv][1]);
          m2 = fmaxf(m2, s_red[par][c][wv][2]);
          m3 = fmaxf(m3, s_red[par][c][wv][3]);
          m4 += s_red[par][c][wv][4];
        }
      pri = m0;
      dua = m1;
      const float eps_p = p.eps_abs + p.eps_rel * m2;
      const float eps_d = p.eps_abs + p.eps_rel * fmaxf(m3, ng);
      if (!(m4 == 0.f) || !(m0 == m0) || !(m1 == m1)) { status = -1; break; }
      if (pri <= eps_p && dua <= eps_d) { status = 1; break; }
      if (it >= p.max_iter) { status = 0; break; }
      if (adapt_now) {
        const float pr_n = m0 / (m2 + 1e-10f);
        const float du_n = m1 / (fmaxf(m3, ng) + 1e-10f);
        float rn = rho * sqrtf(pr_n / (du_n + 1e-10f));
        rn = fminf(fmaxf(rn, p.rho_min), p.rho_max);
        if (fmaxf(pr_n, du_n) > p.rho_adapt_floor &&
            (rn > rho * p.adaptive_rho_tolerance || rn * p.adaptive_rho_tolerance < rho)) {
          rho = rn;
          rho_inv = 1.f / rho;
          factorize();
          continue;
        }
      }
    }
    if (is_row) {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      const float* sp = &s_s[par][0] + rs * COLS;
#pragma unroll
      for (int c = 0; c < COLS; c += 4) {
        const float4 sv4 = *reinterpret_cast<const float4*>(sp + c);
        a0 = fmaf(row[c], sv4.x, a0);
        a1 = fmaf(row[c + 1], sv4.y, a1);
        a2 = fmaf(row[c + 2], sv4.z, a2);
        a3 = fmaf(row[c + 3], sv4.w, a3);
      }
      s_q[rs][ri] = -((a0 + a1) + (a2 + a3));
    }
    __syncthreads();
    if (is_leg && stance) {
      float qv[6];
#pragma unroll
      for (int a = 0; a < 6; ++a) {
        float acc = s_q[0][6 * lj + a];
#pragma unroll
        for (int sl = 1; sl < SPLIT; ++sl) acc += s_q[sl][6 * lj + a];
        qv[a] = acc;
      }
#pragma unroll
      for (int a = 0; a < 6; ++a) vh[a] = fmaf(alpha, qv[a], vh[a]);
      float w3[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float dl = t[k] - dinv * (Gh[0][k] * qv[0] + Gh[1][k] * qv[1] + Gh[2][k] * qv[2] + im * qv[3 + k]);
        const float xt = x[k] + dl;
        x[k] = fmaf(alpha, dl, x[k]);
        const float zh = alpha * xt + (1.f - alpha) * z[k];
        w3[k] = fmaf(y[k], rho_inv, zh);
      }
      project(w3[0], w3[1], w3[2], z[0], z[1], z[2]);
#pragma unroll
      for (int k = 0; k < 3; ++k) y[k] = rho * (w3[k] - z[k]);
    }
    // s_q is rewritten only after the next cluster barrier: no extra local barrier needed
    ++it;
  }

  // ---- outputs ----------------------------------------------------------------------------------
  if (is_leg) {
    float* up = p.U + ((size_t)b * 4 * N + gleg) * 3;
    up[0] = x[0]; up[1] = x[1]; up[2] = x[2];
    float* wx = p.warm_x + ((size_t)slot * 4 * N + gleg) * 3;
    wx[0] = x[0]; wx[1] = x[1]; wx[2] = x[2];
    float* wy = p.warm_y + ((size_t)slot * 4 * N + gleg) * 3;
    wy[0] = y[0]; wy[1] = y[1]; wy[2] = y[2];
  }
  if (tid == 0 && rank == 0) {
    if (p.iters) p.iters[b] = it;
    if (p.pri_res) p.pri_res[b] = pri;
    if (p.dua_res) p.dua_res[b] = dua;
    if (p.status) p.status[b] = status;
    p.warm_valid[slot] = status >= 0 ? 1 : 0;
  }
  // all-gather the final wrench sequence; this barrier also keeps every CTA alive until its
  // peers have stopped pushing into its shared memory
  gather_wrench(x);
  cluster.sync();
  const float* sw = &s_s[ag & 1][0];
  if (p.X) {
    const float dt = p.dt, g = s_x0[12];
    for (int o = rank * THREADS + tid; o < NX; o += CL * THREADS) {
      const int k = o / 13, cidx = o % 13;
      const float kf = (float)k;
      float val;
      if (cidx == 12) {
        val = g;
      } else if (cidx < 3) {
        const float rw0 = cidx == 0 ? (cs * s_x0[6] - sn * s_x0[7])
                                    : (cidx == 1 ? (sn * s_x0[6] + cs * s_x0[7]) : s_x0[8]);
        float acc = 0.f;
        for (int j = 0; j < k; ++j) acc += (float)(k - 1 - j) * sw[6 * j + cidx];
        val = s_x0[cidx] + kf * dt * rw0 + dt * dt * acc;
      } else if (cidx < 6) {
        const int aa = cidx - 3;
        float acc = 0.f;
        for (int j = 0; j < k; ++j) acc += (float)(k - 1 - j) * sw[6 * j + 3 + aa];
        val = s_x0[cidx] + kf * dt * s_x0[9 + aa] + dt * dt * acc;
        if (aa == 2) val += 0.5f * kf * (kf - 1.f) * dt * dt * g;
      } else if (cidx < 9) {
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int j = 0; j < k; ++j) { sx += sw[6 * j]; sy += sw[6 * j + 1]; sz += sw[6 * j + 2]; }
        const int aa = cidx - 6;
        const float rot = aa == 0 ? (cs * sx + sn * sy) : (aa == 1 ? (-sn * sx + cs * sy) : sz);
        val = s_x0[cidx] + dt * rot;
      } else {
        const int aa = cidx - 9;
        float acc = 0.f;
        for (int j = 0; j < k; ++j) acc += sw[6 * j + 3 + aa];
        val = s_x0[cidx] + dt * acc;
        if (aa == 2) val += kf * dt * g;
      }
      p.X[(size_t)b * NX + o] = val;
    }
  }
}

}  // namespace cmpc
