"""CPU study (numpy): which parts of the wrench-space ADMM of the CUDA kernel need more than fp32
for the per-stage wrench to reach 1e-2 N of the tight optimum at long horizons.  Emulates the kernel's
iteration (residual-correction form, gradient carried by v += alpha P^-1 s, exact refresh every 5)
with selectable precision of (a) the refresh product M (G x) and the linear term h, (b) the carried
gradient."""
import sys, os, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, DT
from oracle import srbd_qp, wrench_form as wf, tight_ipm as ipm
from oracle.condensed_admm import project_frustum

def emulate(x0, r, st, xd, mu, N, K, rho, grad64, carry64, f32=np.float32, sigma=1e-6, alpha=1.6,
            adapt=25, tol=3.0, rho_lim=None):
    f = lambda a: np.float32(a).astype(np.float64)
    x0, r, xd = f(x0), f(r), f(xd)
    H, g, Sc, c0, idx = srbd_qp.condensed_qp(x0, r, st, xd, DT)
    G = wf.G_matrix(x0, r, st)                     # 6N x n
    M = wf.M_full(N, DT)
    n = G.shape[1]
    # h: G' h = g  ->  h = M-side linear term: g = 2 Sc' Q e0 = G' h with h = 2 T' Q e0; get it by lstsq-free route:
    # h = M G u* ... simpler: h solves G' h = g in the range; use h = pinv(G') g restricted (exact in fp64)
    h64 = np.linalg.lstsq(G.T, g, rcond=None)[0]
    rho_lim = rho_lim or (0.1 * rho, 300.0)
    G32 = G.astype(f32)
    def factor(rho_):
        d = 1.0 / (sigma + rho_)
        P = np.linalg.inv(M) + d * G @ G.T
        return np.linalg.inv(P).astype(f32), f32(d)
    Pinv, d = factor(rho)
    x = np.zeros(n, f32); y = np.zeros(n, f32)
    z = project_frustum(x.reshape(-1, 3), f32(mu)).reshape(-1).astype(f32)
    cdt = np.float64 if carry64 else f32
    def refresh():
        if grad64:
            return (M @ (G @ x.astype(np.float64)) + h64).astype(cdt)
        w = (G32 @ x).astype(f32)
        return ((M.astype(f32) @ w).astype(f32) + h64.astype(f32)).astype(cdt)
    vh = refresh()
    rho_c = f32(rho)
    for it in range(K):
        if it % 5 == 0 and it:
            vh = refresh()
        gr = (G.T.astype(cdt) @ vh).astype(f32)
        rp = x - z
        if adapt and it and it % adapt == 0:
            pri = np.abs(rp).max(); dua = np.abs(gr + y).max()
            nA = max(np.abs(x).max(), np.abs(z).max())
            nD = max(np.abs(gr - g.astype(f32)).max(), np.abs(y).max(), np.abs(g).max())
            rn = float(rho_c) * np.sqrt((pri / (nA + 1e-10)) / (dua / (nD + 1e-10) + 1e-10))
            rn = min(max(rn, rho_lim[0]), rho_lim[1])
            if rn > float(rho_c) * tol or rn * tol < float(rho_c):
                rho_c = f32(rn); Pinv, d = factor(float(rho_c))
        t = (-d * (gr + y + rho_c * rp)).astype(f32)
        s = (G32 @ t).astype(f32)
        q = (Pinv @ s).astype(f32)
        vh = (vh + cdt(alpha) * q.astype(cdt)).astype(cdt)
        dl = (t - d * (G32.T @ q)).astype(f32)
        xt = x + dl
        x = (x + f32(alpha) * dl).astype(f32)
        zh = f32(alpha) * xt + f32(1 - alpha) * z
        w3 = (zh + y / rho_c).astype(f32)
        z = project_frustum(w3.reshape(-1, 3), f32(mu)).reshape(-1).astype(f32)
        y = (rho_c * (w3 - z)).astype(f32)
    U = np.zeros((N, 12))
    for s_, (i, l) in enumerate(idx):
        U[i, 3 * l:3 * l + 3] = x[3 * s_:3 * s_ + 3]
    return srbd_qp.stage_wrench(U, r), c0 + (Sc @ x.astype(np.float64)).reshape(N + 1, 13).T

if __name__ == "__main__":
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    nb = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    pb = synthetic_batch(64, N=N, seed=0)
    for b in range(nb):
        x0, r, st, xd, mu = pb.problem(b)
        f = lambda a: np.float32(a).astype(np.float64)
        ref = ipm.solve_problem(f(x0), f(r), st, f(xd), mu, DT)
        out = {}
        for name, g64, c64 in (("fp32", False, False), ("grad64", True, False), ("grad64+carry64", True, True)):
            W, X = emulate(x0, r, st, xd, mu, N, K, 0.05 * N, g64, c64)
            dW = np.abs(W - ref["wrench"])
            out[name] = (float(dW.max()), float((dW / (1e-2 + 1e-3 * np.abs(ref["wrench"]))).max()),
                         float(np.abs(X - ref["X"]).max()))
        print(b, {k: ["%.2e" % v for v in vs] for k, vs in out.items()}, flush=True)
