/* ORACLE (test infrastructure and CPU baseline, not product code).
 *
 * Plain-C fp64 restatement of the reference's per-tick QP solve: the QP that
 * `MPC.__init__` declares through CasADi Opti('conic') (reference src/mpc.py:49-173)
 * and that `self.opt.solve()` (src/mpc.py:258) hands to OSQP.  CasADi and OSQP are
 * un-vendored, un-pinned third-party dependencies of the reference (README.md:138)
 * and are absent from this image, so the algorithm restated here is OSQP 0.6-series'
 * published ADMM (Stellato et al. 2020) with its default settings, Ruiz equilibration,
 * rho rules and a sparse quasi-definite LDL' of the KKT matrix (the classic up-looking
 * LDL' of T. Davis that OSQP's QDLDL derives from), in the CasADi call sequence of
 * SURVEY.md Appendix A.  Pinned by tests/golden/simulation_log_golden.npz: it reproduces
 * all 1000 x 12 forces the reference logged (tests/test_oracle_c.py).
 *
 * Same numerical recipe as oracle/osqp_ref.py; this file exists so that the CPU baseline
 * of bench.py runs at C speed on all host cores.
 *
 * Variable order  zeta = [vec(U) (12N) ; vec(X) (13(N+1))], constraint order as CasADi
 * emits it (identity block of variable bounds first, then the 13+66N rows of g).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_TOL 1e-4
#define RHO_EQ_OVER_RHO_INEQ 1e3
#define OSQP_INFTY 1e30
#define BIG 1e300 /* stands for +-inf in l, u */

static const double W_STATE[13] = {1e4, 2.7e4, 1e4, 2.7e5, 2.7e5, 2.7e5, 1e4,
                                   1e4, 1e4,   1.6e4, 1.6e4, 1.6e4, 0.0}; /* src/mpc.py:121-134 */
static const double MASS = 8.885;                                         /* src/mpc.py:71 */
static const double IBODY_INV[3] = {1.0 / 0.24, 1.0, 1.0};                /* src/mpc.py:73-76 */
static const double F_MIN = 3.0, F_MAX = 100.0;                           /* src/mpc.py:45-46 */

typedef struct {
  int N, n, mg, m, nk;          /* variables, g rows, OSQP rows (n+mg), KKT dimension   */
  /* constraint matrix A (m x n) in CSC with fixed structural pattern                   */
  int nnzA, *Ap, *Ai;           /* column pointers / row indices                        */
  int *trip2csc;                /* assembly order -> CSC slot                           */
  int ntrip;
  /* permuted upper-triangular KKT pattern + symbolic factorisation (shared, read-only) */
  int *perm, *pinv, *Kp, *Ki, nnzK;
  int *posP, *posA, *posR;      /* where diag(P+sigma), A entries, -1/rho land in Kx    */
  int *Lp, *Parent, *Lnz0, nnzL;
} Sym;

typedef struct {
  const Sym *s;
  double *Ax, *As;              /* unscaled / scaled A values (CSC order)               */
  double *Pd, *q, *l, *u;       /* scaled problem data                                  */
  double *D, *E, *Dt, *Et;      /* Ruiz scalings                                        */
  double c;
  double *Kx, *Lx, *Dg, *Y;     /* KKT values, factor                                   */
  int *Li, *Lnz, *Pattern, *Flag;
  double *rho_vec, *x, *z, *y, *xt, *zt, *xp, *zp, *rhs, *tmpn, *tmpm, *tmpm2;
  double rho;                   /* persists across solves (OSQP workspace semantics)    */
  double *tv;                   /* triplet values scratch                               */
  /* settings */
  double sigma, alpha, eps_abs, eps_rel;
  int max_iter, scaling, check_termination, adaptive_rho_interval;
  double adaptive_rho_tolerance;
  int last_iters, last_status, rho_updates;
} Work;

/* ---------------------------------------------------------------- assembly ----------- */
static void rotz(double yaw, double R[3][3]) { /* src/mpc.py:64-69 */
  double c = cos(yaw), s = sin(yaw);
  R[0][0] = c; R[0][1] = -s; R[0][2] = 0;
  R[1][0] = s; R[1][1] = c;  R[1][2] = 0;
  R[2][0] = 0; R[2][1] = 0;  R[2][2] = 1;
}

/* Emits the entries of the g-part of A in a fixed order.  If rows != NULL records the
 * pattern, if vals != NULL the numeric values.  Also fills bounds lb/ub (mg) when given.
 * x0 (13), r (N*4*3), swing (4*N, [l*N+i]), mu. */
static int assemble_g(int N, const double *x0, const double *r, const double *swing, double mu,
                      double delta, double g, int *rows, int *cols, double *vals, double *lb,
                      double *ub) {
  const int nU = 12 * N;
  int t = 0;
#define CU(k, i) (12 * (i) + (k))
#define CX(k, i) (nU + 13 * (i) + (k))
#define EMIT(rw, cl, v)                 \
  do {                                  \
    if (rows) { rows[t] = (rw); cols[t] = (cl); } \
    if (vals) vals[t] = (v);            \
    ++t;                                \
  } while (0)
  double Rz[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, Ihat[3][3] = {{0}};
  if (x0) {
    rotz(x0[2], Rz);
    for (int a = 0; a < 3; ++a) /* I_hat_inv = Rz diag Rz' (src/mpc.py:78) */
      for (int b = 0; b < 3; ++b) {
        double acc = 0;
        for (int k = 0; k < 3; ++k) acc += Rz[a][k] * IBODY_INV[k] * Rz[b][k];
        Ihat[a][b] = acc;
      }
  }
  /* X[:,0] == x0                                      src/mpc.py:113 */
  for (int k = 0; k < 13; ++k) {
    EMIT(k, CX(k, 0), 1.0);
    if (lb) lb[k] = ub[k] = x0[k];
  }
  /* X_{i+1} - X_i - delta (A X_i + B_i U_i) == 0      src/mpc.py:116-117 */
  for (int i = 0; i < N; ++i) {
    const int base = 13 + 13 * i;
    for (int k = 0; k < 13; ++k) {
      EMIT(base + k, CX(k, i + 1), 1.0);
      EMIT(base + k, CX(k, i), -1.0);
      if (lb) lb[base + k] = ub[base + k] = 0.0;
    }
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) EMIT(base + a, CX(6 + b, i), -delta * Rz[a][b]); /* Theta' = Rz w */
      EMIT(base + 3 + a, CX(9 + a, i), -delta);                                     /* p' = v */
    }
    EMIT(base + 11, CX(12, i), -delta);                                             /* vz' += g-state */
    for (int l = 0; l < 4; ++l) {
      const double *rl = r ? r + (i * 4 + l) * 3 : NULL;
      double S[3][3] = {{0}};
      if (rl) { /* src/utils.py:43-56 */
        S[0][1] = -rl[2]; S[0][2] = rl[1];
        S[1][0] = rl[2];  S[1][2] = -rl[0];
        S[2][0] = -rl[1]; S[2][1] = rl[0];
      }
      for (int a = 0; a < 3; ++a) {
        for (int b = 0; b < 3; ++b) {
          double acc = 0;
          for (int k = 0; k < 3; ++k) acc += Ihat[a][k] * S[k][b];
          EMIT(base + 6 + a, CU(3 * l + b, i), -delta * acc);
        }
        EMIT(base + 9 + a, CU(3 * l + a, i), -delta / MASS);
      }
    }
  }
  /* swing[l,i] * U[3l:3l+3, i] == 0                   src/mpc.py:139-144 */
  for (int i = 0; i < N; ++i)
    for (int l = 0; l < 4; ++l)
      for (int k = 0; k < 3; ++k) {
        const int rw = 13 + 13 * N + 12 * i + 3 * l + k;
        EMIT(rw, CU(3 * l + k, i), swing ? swing[l * N + i] : 0.0);
        if (lb) lb[rw] = ub[rw] = 0.0;
      }
  /* per-stage block of 41 rows                        src/mpc.py:148-173 */
  for (int i = 0; i < N; ++i) {
    const int base = 13 + 25 * N + 41 * i;
    EMIT(base, CX(12, i), 1.0);
    if (lb) lb[base] = ub[base] = g;
    for (int l = 0; l < 4; ++l) {
      const double cond = swing ? 1.0 - swing[l * N + i] : 0.0;
      EMIT(base + 1 + 2 * l, CU(3 * l + 2, i), cond);
      EMIT(base + 2 + 2 * l, CU(3 * l + 2, i), cond);
      if (lb) {
        lb[base + 1 + 2 * l] = cond * F_MIN; ub[base + 1 + 2 * l] = BIG;
        lb[base + 2 + 2 * l] = -BIG;         ub[base + 2 + 2 * l] = cond * F_MAX;
      }
    }
    for (int blk = 0; blk < 2; ++blk) { /* fy rows (offset 9) then fx rows (offset 25) */
      const int off = blk == 0 ? 9 : 25, comp = blk == 0 ? 1 : 0;
      static const double sg[4] = {-1.0, 1.0, 1.0, -1.0};
      for (int l = 0; l < 4; ++l)
        for (int j = 0; j < 4; ++j) {
          const int rw = base + off + 4 * l + j;
          EMIT(rw, CU(3 * l + comp, i), sg[j]);
          EMIT(rw, CU(3 * l + 2, i), -mu);
          if (lb) { lb[rw] = -BIG; ub[rw] = 0.0; }
        }
    }
  }
  return t;
#undef CU
#undef CX
#undef EMIT
}

/* ---------------------------------------------------------------- ordering ----------- */
typedef struct { int *v; int len, cap; } IVec;
static void iv_push(IVec *a, int x) {
  if (a->len == a->cap) { a->cap = a->cap ? 2 * a->cap : 8; a->v = (int *)realloc(a->v, sizeof(int) * a->cap); }
  a->v[a->len++] = x;
}
static int iv_has(const IVec *a, int x) {
  for (int i = 0; i < a->len; ++i) if (a->v[i] == x) return 1;
  return 0;
}
static void iv_del(IVec *a, int x) {
  for (int i = 0; i < a->len; ++i) if (a->v[i] == x) { a->v[i] = a->v[--a->len]; return; }
}

/* exact minimum-degree ordering on the elimination graph (done once per horizon) */
static void min_degree(int n, const IVec *adj0, int *perm) {
  IVec *adj = (IVec *)calloc(n, sizeof(IVec));
  char *dead = (char *)calloc(n, 1);
  for (int i = 0; i < n; ++i) for (int k = 0; k < adj0[i].len; ++k) iv_push(&adj[i], adj0[i].v[k]);
  for (int step = 0; step < n; ++step) {
    int best = -1, bd = 1 << 30;
    for (int i = 0; i < n; ++i) if (!dead[i] && adj[i].len < bd) { bd = adj[i].len; best = i; }
    perm[step] = best;
    dead[best] = 1;
    IVec *nb = &adj[best];
    for (int a = 0; a < nb->len; ++a) iv_del(&adj[nb->v[a]], best);
    for (int a = 0; a < nb->len; ++a)
      for (int b = a + 1; b < nb->len; ++b) {
        int u = nb->v[a], v = nb->v[b];
        if (!iv_has(&adj[u], v)) { iv_push(&adj[u], v); iv_push(&adj[v], u); }
      }
    free(nb->v); nb->v = NULL; nb->len = nb->cap = 0;
  }
  for (int i = 0; i < n; ++i) free(adj[i].v);
  free(adj); free(dead);
}

/* ---------------------------------------------------------------- symbolic ----------- */
static int cmp_int2(const void *a, const void *b) {
  const int *x = (const int *)a, *y = (const int *)b;
  if (x[0] != y[0]) return x[0] - y[0];
  return x[1] - y[1];
}

Sym *osqpref_sym_create(int N) {
  Sym *s = (Sym *)calloc(1, sizeof(Sym));
  s->N = N;
  s->n = 12 * N + 13 * (N + 1);
  s->mg = 13 + 66 * N;
  s->m = s->n + s->mg;
  s->nk = s->n + s->m;
  const int n = s->n, m = s->m;
  /* pattern of A = [I ; A_g] */
  int ntg = assemble_g(N, NULL, NULL, NULL, 0, 0, 0, NULL, NULL, NULL, NULL, NULL);
  int *rows = (int *)malloc(sizeof(int) * ntg), *cols = (int *)malloc(sizeof(int) * ntg);
  assemble_g(N, NULL, NULL, NULL, 0, 0, 0, rows, cols, NULL, NULL, NULL);
  s->ntrip = ntg;
  /* CSC of A: identity entries first in each column, then g entries sorted by row */
  int (*key)[3] = (int (*)[3])malloc(sizeof(int[3]) * (ntg + n));
  for (int j = 0; j < n; ++j) { key[j][0] = j; key[j][1] = j; key[j][2] = -1 - j; }
  for (int t = 0; t < ntg; ++t) { key[n + t][0] = cols[t]; key[n + t][1] = n + rows[t]; key[n + t][2] = t; }
  qsort(key, ntg + n, sizeof(int[3]), cmp_int2);
  s->nnzA = ntg + n;
  s->Ap = (int *)calloc(n + 1, sizeof(int));
  s->Ai = (int *)malloc(sizeof(int) * s->nnzA);
  s->trip2csc = (int *)malloc(sizeof(int) * ntg);
  for (int e = 0; e < s->nnzA; ++e) {
    s->Ap[key[e][0] + 1]++;
    s->Ai[e] = key[e][1];
    if (key[e][2] >= 0) s->trip2csc[key[e][2]] = e;
  }
  for (int j = 0; j < n; ++j) s->Ap[j + 1] += s->Ap[j];
  free(key); free(rows); free(cols);
  /* KKT graph: variable j <-> constraint node n+i for every A entry */
  const int nk = s->nk;
  IVec *adj = (IVec *)calloc(nk, sizeof(IVec));
  for (int j = 0; j < n; ++j)
    for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) {
      int i = n + s->Ai[p];
      iv_push(&adj[j], i); iv_push(&adj[i], j);
    }
  s->perm = (int *)malloc(sizeof(int) * nk);
  s->pinv = (int *)malloc(sizeof(int) * nk);
  min_degree(nk, adj, s->perm);
  for (int i = 0; i < nk; ++i) { s->pinv[s->perm[i]] = i; free(adj[i].v); }
  free(adj);
  /* permuted upper-triangular KKT: entries = diag(n) + A entries + diag(m) */
  const int ne = n + s->nnzA + m;
  int (*ek)[3] = (int (*)[3])malloc(sizeof(int[3]) * ne);
  int e = 0;
  for (int j = 0; j < n; ++j, ++e) { ek[e][0] = s->pinv[j]; ek[e][1] = s->pinv[j]; ek[e][2] = e; }
  for (int j = 0; j < n; ++j)
    for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p, ++e) {
      int a = s->pinv[j], b = s->pinv[n + s->Ai[p]];
      ek[e][0] = a > b ? a : b;  /* column */
      ek[e][1] = a > b ? b : a;  /* row <= column */
      ek[e][2] = e;
    }
  for (int i = 0; i < m; ++i, ++e) { ek[e][0] = s->pinv[n + i]; ek[e][1] = s->pinv[n + i]; ek[e][2] = e; }
  qsort(ek, ne, sizeof(int[3]), cmp_int2);
  s->nnzK = ne;
  s->Kp = (int *)calloc(nk + 1, sizeof(int));
  s->Ki = (int *)malloc(sizeof(int) * ne);
  int *pos = (int *)malloc(sizeof(int) * ne);
  for (int k = 0; k < ne; ++k) { s->Kp[ek[k][0] + 1]++; s->Ki[k] = ek[k][1]; pos[ek[k][2]] = k; }
  for (int j = 0; j < nk; ++j) s->Kp[j + 1] += s->Kp[j];
  s->posP = (int *)malloc(sizeof(int) * n);
  s->posA = (int *)malloc(sizeof(int) * s->nnzA);
  s->posR = (int *)malloc(sizeof(int) * m);
  memcpy(s->posP, pos, sizeof(int) * n);
  memcpy(s->posA, pos + n, sizeof(int) * s->nnzA);
  memcpy(s->posR, pos + n + s->nnzA, sizeof(int) * m);
  free(pos); free(ek);
  /* elimination tree and column counts (up-looking LDL' symbolic phase) */
  s->Lp = (int *)calloc(nk + 1, sizeof(int));
  s->Parent = (int *)malloc(sizeof(int) * nk);
  s->Lnz0 = (int *)calloc(nk, sizeof(int));
  int *Flag = (int *)malloc(sizeof(int) * nk);
  for (int k = 0; k < nk; ++k) {
    s->Parent[k] = -1; Flag[k] = k;
    for (int p = s->Kp[k]; p < s->Kp[k + 1]; ++p) {
      int i = s->Ki[p];
      if (i < k)
        for (; Flag[i] != k; i = s->Parent[i]) {
          if (s->Parent[i] == -1) s->Parent[i] = k;
          s->Lnz0[i]++; Flag[i] = k;
        }
    }
  }
  for (int k = 0; k < nk; ++k) s->Lp[k + 1] = s->Lp[k] + s->Lnz0[k];
  s->nnzL = s->Lp[nk];
  free(Flag);
  return s;
}

void osqpref_sym_free(Sym *s) {
  if (!s) return;
  free(s->Ap); free(s->Ai); free(s->trip2csc); free(s->perm); free(s->pinv); free(s->Kp); free(s->Ki);
  free(s->posP); free(s->posA); free(s->posR); free(s->Lp); free(s->Parent); free(s->Lnz0); free(s);
}

int osqpref_sym_nnzL(const Sym *s) { return s->nnzL; }
int osqpref_sym_nvars(const Sym *s) { return s->n; }

/* ---------------------------------------------------------------- workspace ---------- */
#define DALLOC(k) ((double *)calloc((size_t)(k) + 1, sizeof(double)))
Work *osqpref_work_create(const Sym *s) {
  Work *w = (Work *)calloc(1, sizeof(Work));
  const int n = s->n, m = s->m, nk = s->nk;
  w->s = s;
  w->Ax = DALLOC(s->nnzA); w->As = DALLOC(s->nnzA);
  w->Pd = DALLOC(n); w->q = DALLOC(n); w->l = DALLOC(m); w->u = DALLOC(m);
  w->D = DALLOC(n); w->E = DALLOC(m); w->Dt = DALLOC(n); w->Et = DALLOC(m);
  w->Kx = DALLOC(s->nnzK); w->Lx = DALLOC(s->nnzL); w->Dg = DALLOC(nk); w->Y = DALLOC(nk);
  w->Li = (int *)calloc(s->nnzL + 1, sizeof(int)); w->Lnz = (int *)calloc(nk, sizeof(int));
  w->Pattern = (int *)calloc(nk, sizeof(int)); w->Flag = (int *)calloc(nk, sizeof(int));
  w->rho_vec = DALLOC(m); w->x = DALLOC(n); w->z = DALLOC(m); w->y = DALLOC(m);
  w->xt = DALLOC(n); w->zt = DALLOC(m); w->xp = DALLOC(n); w->zp = DALLOC(m);
  w->rhs = DALLOC(nk); w->tmpn = DALLOC(n); w->tmpm = DALLOC(m); w->tmpm2 = DALLOC(m);
  w->tv = DALLOC(s->ntrip);
  w->rho = 0.1; w->sigma = 1e-6; w->alpha = 1.6; w->eps_abs = 1e-3; w->eps_rel = 1e-3;
  w->max_iter = 1000; w->scaling = 10; w->check_termination = 25; w->adaptive_rho_interval = 100;
  w->adaptive_rho_tolerance = 5.0;
  return w;
}
void osqpref_work_free(Work *w) {
  if (!w) return;
  free(w->Ax); free(w->As); free(w->Pd); free(w->q); free(w->l); free(w->u); free(w->D); free(w->E);
  free(w->Dt); free(w->Et); free(w->Kx); free(w->Lx); free(w->Dg); free(w->Y); free(w->Li); free(w->Lnz);
  free(w->Pattern); free(w->Flag); free(w->rho_vec); free(w->x); free(w->z); free(w->y); free(w->xt);
  free(w->zt); free(w->xp); free(w->zp); free(w->rhs); free(w->tmpn); free(w->tmpm); free(w->tmpm2);
  free(w->tv); free(w);
}
void osqpref_set_rho(Work *w, double rho) { w->rho = rho; }
double osqpref_get_rho(const Work *w) { return w->rho; }
void osqpref_set_tolerances(Work *w, double eps_abs, double eps_rel, int max_iter) {
  w->eps_abs = eps_abs; w->eps_rel = eps_rel; w->max_iter = max_iter;
}
int osqpref_last_iters(const Work *w) { return w->last_iters; }
int osqpref_rho_updates(const Work *w) { return w->rho_updates; }

static double limit_scaling(double v) {
  if (v < MIN_SCALING) v = 1.0;
  if (v > MAX_SCALING) v = MAX_SCALING;
  return v;
}

/* Ruiz equilibration, SURVEY.md Appendix A-3 */
static void scale_data(Work *w) {
  const Sym *s = w->s;
  const int n = s->n, m = s->m;
  for (int j = 0; j < n; ++j) w->D[j] = 1.0;
  for (int i = 0; i < m; ++i) w->E[i] = 1.0;
  w->c = 1.0;
  memcpy(w->As, w->Ax, sizeof(double) * s->nnzA);
  for (int pass = 0; pass < w->scaling; ++pass) {
    for (int i = 0; i < m; ++i) w->Et[i] = 0.0;
    for (int j = 0; j < n; ++j) {
      double cn = fabs(w->Pd[j]);
      for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) {
        const double a = fabs(w->As[p]);
        if (a > cn) cn = a;
        if (a > w->Et[s->Ai[p]]) w->Et[s->Ai[p]] = a;
      }
      w->Dt[j] = 1.0 / sqrt(limit_scaling(cn));
    }
    for (int i = 0; i < m; ++i) w->Et[i] = 1.0 / sqrt(limit_scaling(w->Et[i]));
    double pmean = 0.0, qn = 0.0;
    for (int j = 0; j < n; ++j) {
      w->Pd[j] = w->Dt[j] * w->Pd[j] * w->Dt[j];
      for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) w->As[p] *= w->Et[s->Ai[p]] * w->Dt[j];
      w->q[j] *= w->Dt[j];
      w->D[j] *= w->Dt[j];
      pmean += fabs(w->Pd[j]);
      if (fabs(w->q[j]) > qn) qn = fabs(w->q[j]);
    }
    for (int i = 0; i < m; ++i) w->E[i] *= w->Et[i];
    double ct = pmean / n;
    qn = limit_scaling(qn);
    if (qn > ct) ct = qn;
    ct = 1.0 / limit_scaling(ct);
    for (int j = 0; j < n; ++j) { w->Pd[j] *= ct; w->q[j] *= ct; }
    w->c *= ct;
  }
  for (int i = 0; i < m; ++i) {
    if (w->l[i] > -BIG) w->l[i] *= w->E[i];
    if (w->u[i] < BIG) w->u[i] *= w->E[i];
  }
}

static void compute_rho_vec(Work *w) {
  const int m = w->s->m;
  for (int i = 0; i < m; ++i) {
    if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) w->rho_vec[i] = RHO_MIN;
    else if (w->u[i] - w->l[i] < RHO_TOL) w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->rho;
    else w->rho_vec[i] = w->rho;
  }
}

/* numeric up-looking LDL' of the permuted KKT matrix */
static int factor(Work *w) {
  const Sym *s = w->s;
  const int n = s->n, m = s->m, nk = s->nk;
  for (int j = 0; j < n; ++j) w->Kx[s->posP[j]] = w->Pd[j] + w->sigma;
  for (int p = 0; p < s->nnzA; ++p) w->Kx[s->posA[p]] = w->As[p];
  for (int i = 0; i < m; ++i) w->Kx[s->posR[i]] = -1.0 / w->rho_vec[i];
  double *Y = w->Y, *Lx = w->Lx, *D = w->Dg;
  int *Li = w->Li, *Lnz = w->Lnz, *Pattern = w->Pattern, *Flag = w->Flag;
  const int *Lp = s->Lp, *Parent = s->Parent;
  for (int k = 0; k < nk; ++k) {
    Y[k] = 0.0;
    int top = nk;
    Flag[k] = k;
    Lnz[k] = 0;
    for (int p = s->Kp[k]; p < s->Kp[k + 1]; ++p) {
      int i = s->Ki[p];
      Y[i] += w->Kx[p];
      int len = 0;
      for (; Flag[i] != k; i = Parent[i]) { Pattern[len++] = i; Flag[i] = k; }
      while (len > 0) Pattern[--top] = Pattern[--len];
    }
    D[k] = Y[k];
    Y[k] = 0.0;
    for (; top < nk; ++top) {
      const int i = Pattern[top];
      const double yi = Y[i];
      Y[i] = 0.0;
      const int p2 = Lp[i] + Lnz[i];
      for (int p = Lp[i]; p < p2; ++p) Y[Li[p]] -= Lx[p] * yi;
      const double lki = yi / D[i];
      D[k] -= lki * yi;
      Li[p2] = k;
      Lx[p2] = lki;
      Lnz[i]++;
    }
    if (D[k] == 0.0) return -1;
  }
  return 0;
}

/* solves K sol = rhs in place; rhs is in original ordering */
static void kkt_solve(Work *w, double *b) {
  const Sym *s = w->s;
  const int nk = s->nk;
  double *X = w->Y;
  for (int i = 0; i < nk; ++i) X[i] = b[s->perm[i]];
  for (int j = 0; j < nk; ++j) {
    const double xj = X[j];
    for (int p = s->Lp[j]; p < s->Lp[j] + w->Lnz[j]; ++p) X[w->Li[p]] -= w->Lx[p] * xj;
  }
  for (int j = 0; j < nk; ++j) X[j] /= w->Dg[j];
  for (int j = nk - 1; j >= 0; --j) {
    double xj = X[j];
    for (int p = s->Lp[j]; p < s->Lp[j] + w->Lnz[j]; ++p) xj -= w->Lx[p] * X[w->Li[p]];
    X[j] = xj;
  }
  for (int i = 0; i < nk; ++i) b[s->perm[i]] = X[i];
  for (int i = 0; i < nk; ++i) X[i] = 0.0;
}

static void A_mul(const Work *w, const double *x, double *out) { /* out = As x */
  const Sym *s = w->s;
  memset(out, 0, sizeof(double) * s->m);
  for (int j = 0; j < s->n; ++j) {
    const double xj = x[j];
    if (xj != 0.0)
      for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) out[s->Ai[p]] += w->As[p] * xj;
  }
}
static void At_mul(const Work *w, const double *y, double *out) { /* out = As' y */
  const Sym *s = w->s;
  for (int j = 0; j < s->n; ++j) {
    double acc = 0.0;
    for (int p = s->Ap[j]; p < s->Ap[j + 1]; ++p) acc += w->As[p] * y[s->Ai[p]];
    out[j] = acc;
  }
}
static double ninf(const double *v, int k) {
  double m = 0.0;
  for (int i = 0; i < k; ++i) if (fabs(v[i]) > m) m = fabs(v[i]);
  return m;
}
static double ninf_scaled(const double *v, const double *sc, int k, int inv) {
  double m = 0.0;
  for (int i = 0; i < k; ++i) {
    const double a = fabs(inv ? v[i] / sc[i] : v[i] * sc[i]);
    if (a > m) m = a;
  }
  return m;
}

/* One CasADi `solve()`.  x_warm: previous UNSCALED primal solution or NULL (cold).
 * sol (n) receives [vec(U); vec(X)].  Returns 1 if "solved", 0 otherwise. */
int osqpref_solve(Work *w, const double *x0, const double *r, const double *swing,
                  const double *x_des /* 13 x (N+1), row-major [k*(N+1)+i] */, double mu,
                  double delta, double g, const double *x_warm, double *sol) {
  const Sym *s = w->s;
  const int N = s->N, n = s->n, m = s->m, mg = s->mg;
  /* numeric data */
  assemble_g(N, x0, r, swing, mu, delta, g, NULL, NULL, w->tv, w->l + n, w->u + n);
  for (int j = 0; j < n; ++j) { w->l[j] = -BIG; w->u[j] = BIG; }
  (void)mg;
  for (int j = 0; j < n; ++j) w->Ax[s->Ap[j]] = 1.0; /* identity entry is first in each column */
  for (int t = 0; t < s->ntrip; ++t) w->Ax[s->trip2csc[t]] = w->tv[t];
  for (int j = 0; j < 12 * N; ++j) { w->Pd[j] = 0.0; w->q[j] = 0.0; }
  for (int i = 0; i <= N; ++i)
    for (int k = 0; k < 13; ++k) { /* src/mpc.py:120-136 */
      const int j = 12 * N + 13 * i + k;
      w->Pd[j] = 2.0 * W_STATE[k];
      w->q[j] = -2.0 * W_STATE[k] * x_des[k * (N + 1) + i];
    }
  scale_data(w);
  compute_rho_vec(w);
  if (factor(w)) return -1;
  /* warm start: x <- D^-1 x_prev, z <- A x, y <- 0 */
  for (int j = 0; j < n; ++j) w->x[j] = x_warm ? x_warm[j] / w->D[j] : 0.0;
  A_mul(w, w->x, w->z);
  memset(w->y, 0, sizeof(double) * m);
  int status = 0, it;
  w->rho_updates = 0;
  const double cinv = 1.0 / w->c;
  for (it = 1; it <= w->max_iter; ++it) {
    memcpy(w->xp, w->x, sizeof(double) * n);
    memcpy(w->zp, w->z, sizeof(double) * m);
    for (int j = 0; j < n; ++j) w->rhs[j] = w->sigma * w->xp[j] - w->q[j];
    for (int i = 0; i < m; ++i) w->rhs[n + i] = w->zp[i] - w->y[i] / w->rho_vec[i];
    kkt_solve(w, w->rhs);
    for (int j = 0; j < n; ++j) w->x[j] = w->alpha * w->rhs[j] + (1.0 - w->alpha) * w->xp[j];
    for (int i = 0; i < m; ++i) {
      const double zt = w->zp[i] + (w->rhs[n + i] - w->y[i]) / w->rho_vec[i];
      const double zh = w->alpha * zt + (1.0 - w->alpha) * w->zp[i];
      double zn = zh + w->y[i] / w->rho_vec[i];
      if (zn < w->l[i]) zn = w->l[i];
      if (zn > w->u[i]) zn = w->u[i];
      w->y[i] += w->rho_vec[i] * (zh - zn);
      w->z[i] = zn;
    }
    const int check = w->check_termination && it % w->check_termination == 0;
    const int adapt = w->adaptive_rho_interval && it % w->adaptive_rho_interval == 0;
    if (check || adapt) {
      A_mul(w, w->x, w->tmpm);            /* Ax  */
      At_mul(w, w->y, w->tmpn);           /* A'y */
    }
    if (check) {
      double pri = 0, nAx = 0, nz = 0, dua = 0, nPx = 0, nAty = 0, nq = 0;
      for (int i = 0; i < m; ++i) {
        const double ei = 1.0 / w->E[i];
        const double a = fabs(ei * (w->tmpm[i] - w->z[i]));
        if (a > pri) pri = a;
        if (fabs(ei * w->tmpm[i]) > nAx) nAx = fabs(ei * w->tmpm[i]);
        if (fabs(ei * w->z[i]) > nz) nz = fabs(ei * w->z[i]);
      }
      for (int j = 0; j < n; ++j) {
        const double di = 1.0 / w->D[j], px = w->Pd[j] * w->x[j];
        const double a = fabs(di * (px + w->q[j] + w->tmpn[j]));
        if (a > dua) dua = a;
        if (fabs(di * px) > nPx) nPx = fabs(di * px);
        if (fabs(di * w->tmpn[j]) > nAty) nAty = fabs(di * w->tmpn[j]);
        if (fabs(di * w->q[j]) > nq) nq = fabs(di * w->q[j]);
      }
      dua *= cinv;
      double mx = nPx > nAty ? nPx : nAty;
      if (nq > mx) mx = nq;
      const double eps_p = w->eps_abs + w->eps_rel * (nAx > nz ? nAx : nz);
      const double eps_d = w->eps_abs + w->eps_rel * cinv * mx;
      if (pri < eps_p && dua < eps_d) { status = 1; break; }
    }
    if (adapt) {
      double pr = 0, dr = 0, nPx = 0;
      for (int i = 0; i < m; ++i) { const double a = fabs(w->tmpm[i] - w->z[i]); if (a > pr) pr = a; }
      for (int j = 0; j < n; ++j) {
        const double px = w->Pd[j] * w->x[j];
        const double a = fabs(px + w->q[j] + w->tmpn[j]);
        if (a > dr) dr = a;
        if (fabs(px) > nPx) nPx = fabs(px);
      }
      const double nAx = ninf(w->tmpm, m), nz = ninf(w->z, m);
      const double nAty = ninf(w->tmpn, n), nq = ninf(w->q, n);
      pr /= (nAx > nz ? nAx : nz) + 1e-10;
      double mx = nPx > nAty ? nPx : nAty;
      if (nq > mx) mx = nq;
      dr /= mx + 1e-10;
      double rho_new = w->rho * sqrt(pr / (dr + 1e-10));
      if (rho_new < RHO_MIN) rho_new = RHO_MIN;
      if (rho_new > RHO_MAX) rho_new = RHO_MAX;
      if (rho_new > w->rho * w->adaptive_rho_tolerance || rho_new < w->rho / w->adaptive_rho_tolerance) {
        w->rho = rho_new;
        compute_rho_vec(w);
        if (factor(w)) return -1;
        w->rho_updates++;
      }
    }
  }
  if (it > w->max_iter) it = w->max_iter;
  w->last_iters = it;
  w->last_status = status;
  for (int j = 0; j < n; ++j) sol[j] = w->D[j] * w->x[j];
  (void)ninf_scaled;
  return status;
}

/* Batch of independent cold-start problems over all host threads (CPU baseline of bench.py).
 * x0 [B,13], r [B,N,4,3], stance [B,N,4] (1 = stance), x_des [B,N+1,13] (ABI layout), mu [B].
 * U_out [B,N,12], iters [B], status [B].  Returns the number of threads used. */
typedef struct {
  const Sym *s;
  int N, B;
  const double *x0, *r, *stance, *x_des, *mu;
  double delta, g;
  double *U_out;
  int *iters, *status;
  int *next;
} BatchJob;

static void *batch_worker(void *arg) {
  BatchJob *j = (BatchJob *)arg;
  const Sym *s = j->s;
  const int N = j->N;
  Work *w = osqpref_work_create(s);
  double *swing = (double *)malloc(sizeof(double) * 4 * N);
  double *xd = (double *)malloc(sizeof(double) * 13 * (N + 1));
  double *sol = (double *)malloc(sizeof(double) * s->n);
  for (;;) {
    const int b = __sync_fetch_and_add(j->next, 1);
    if (b >= j->B) break;
    for (int i = 0; i < N; ++i)
      for (int l = 0; l < 4; ++l) swing[l * N + i] = 1.0 - j->stance[((size_t)b * N + i) * 4 + l];
    for (int i = 0; i <= N; ++i)
      for (int k = 0; k < 13; ++k) xd[k * (N + 1) + i] = j->x_des[((size_t)b * (N + 1) + i) * 13 + k];
    w->rho = 0.1; /* every problem is its own fresh MPC instance */
    int st = osqpref_solve(w, j->x0 + (size_t)b * 13, j->r + (size_t)b * N * 12, swing, xd, j->mu[b],
                           j->delta, j->g, NULL, sol);
    memcpy(j->U_out + (size_t)b * 12 * N, sol, sizeof(double) * 12 * N);
    if (j->iters) j->iters[b] = w->last_iters;
    if (j->status) j->status[b] = st;
  }
  free(swing); free(xd); free(sol);
  osqpref_work_free(w);
  return NULL;
}

int osqpref_solve_batch(int N, int B, const double *x0, const double *r, const double *stance,
                        const double *x_des, const double *mu, double delta, double g,
                        double *U_out, int *iters, int *status, int nthreads) {
  /* symbolic analysis is setup work (the reference does it once in MPC.__init__): cached */
  static pthread_mutex_t mtx = PTHREAD_MUTEX_INITIALIZER;
  static Sym *cache[128] = {0};
  pthread_mutex_lock(&mtx);
  if (N < 128 && !cache[N]) cache[N] = osqpref_sym_create(N);
  Sym *s = N < 128 ? cache[N] : osqpref_sym_create(N);
  pthread_mutex_unlock(&mtx);
  if (nthreads <= 0) nthreads = (int)sysconf(_SC_NPROCESSORS_ONLN);
  if (nthreads > B) nthreads = B > 0 ? B : 1;
  if (nthreads > 256) nthreads = 256;
  int next = 0;
  BatchJob job = {s, N, B, x0, r, stance, x_des, mu, delta, g, U_out, iters, status, &next};
  pthread_t th[256];
  for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, batch_worker, &job);
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  if (N >= 128) osqpref_sym_free(s);
  return nthreads;
}
