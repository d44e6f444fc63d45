"""CPU study of ADMM iteration counts on config 2 (fp64, all problems at once).

Uses the eigen-decomposition of each H so that a rho change costs nothing here; the question is
only how many iterations each policy needs (mean and, above all, the maximum, which sets the
batch time of the one-CTA-per-problem kernel).  Test/experiment infrastructure: imports oracle/.
usage: python scripts/cpu_rho_study.py [B] [N] [seed]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpc_b200  # noqa: F401  (package alias)
from mpc_b200.problems import synthetic_batch
from oracle import batched_admm
from oracle.condensed_admm import project_frustum


def prepare(B, N, seed, gaits=("trot",), mu=(1.0, 1.0)):
    pb = synthetic_batch(B, N=N, gaits=gaits, seed=seed, mu=mu)
    H, g = batched_admm.build(pb)
    keep = np.stack([pb.problem(b)[2].reshape(-1).astype(bool) for b in range(B)])
    lam, V = np.linalg.eigh(H)
    mus = np.array([pb.problem(b)[4] for b in range(B)], dtype=float)
    return dict(H=H, g=g, keep=keep, lam=lam, V=V, mu=mus, B=B, n=g.shape[1])


def run(P, rho0=0.5, sigma=1e-6, alpha=1.6, eps=1e-3, max_iter=1000, check_every=5,
        adapt_at=lambda it: it > 0 and it % 25 == 0, tol=3.0, rho_lim=(0.05, 300.0), rho_stage=None,
        f_min=3.0, f_max=100.0, power=lambda it: 0.5):
    """rho0 scalar or (B,) ; rho_stage optional (B, n) multiplicative per-variable profile."""
    H, g, lam, V, B, n = P["H"], P["g"], P["lam"], P["V"], P["B"], P["n"]
    k3 = np.repeat(P["keep"], 3, axis=1)
    mu3 = P["mu"][:, None]
    L = n // 3
    rho = np.broadcast_to(np.asarray(rho0, float), (B,)).copy()
    proj = lambda v, idx: project_frustum(v.reshape(len(idx), L, 3), mu3[idx], f_min, f_max).reshape(len(idx), n) * k3[idx]
    x = np.zeros((B, n)); y = np.zeros((B, n)); z = proj(x, np.arange(B))
    iters = np.full(B, max_iter); done = np.zeros(B, bool); nup = np.zeros(B, int)
    ng = np.abs(g).max(1)
    for it in range(max_iter + 1):
        act = np.where(~done)[0]
        if len(act) == 0:
            break
        if it % check_every == 0 or adapt_at(it):
            Hx = np.einsum('bij,bj->bi', H[act], x[act])
            pri = np.abs(x[act] - z[act]).max(1)
            dua = np.abs(Hx + g[act] + y[act]).max(1)
            nA = np.maximum(np.abs(x[act]).max(1), np.abs(z[act]).max(1))
            nD = np.maximum(np.maximum(np.abs(Hx).max(1), np.abs(y[act]).max(1)), ng[act])
            if it % check_every == 0:
                ok = (pri <= eps + eps * nA) & (dua <= eps + eps * nD)
                iters[act[ok]] = it
                done[act[ok]] = True
            else:
                ok = np.zeros(len(act), bool)
            if adapt_at(it):
                rn = rho[act] * ((pri / (nA + 1e-10)) / (dua / (nD + 1e-10) + 1e-10)) ** power(it)
                rn = np.clip(rn, rho_lim[0], rho_lim[1])
                upd = ((rn > tol * rho[act]) | (rn * tol < rho[act])) & ~ok
                rho[act[upd]] = rn[upd]
                nup[act[upd]] += 1
            act = np.where(~done)[0]
            if len(act) == 0:
                break
        r_ = rho[act][:, None]
        rhs = (sigma * x[act] - g[act] + r_ * z[act] - y[act]) * k3[act]
        t = np.einsum('bji,bj->bi', V[act], rhs) / (lam[act] + sigma + r_)
        xt = np.einsum('bij,bj->bi', V[act], t) * k3[act]
        xn = alpha * xt + (1 - alpha) * x[act]
        zh = alpha * xt + (1 - alpha) * z[act]
        zn = proj(zh + y[act] / r_, act)
        yn = (y[act] + r_ * (zh - zn)) * k3[act]
        x[act], z[act], y[act] = xn, zn, yn
    return dict(iters=iters, done=done, rho=rho, nup=nup, x=x, y=y)


def report(name, out):
    it = out["iters"]
    print(f"{name:46s} mean {it.mean():6.1f} p50 {np.percentile(it,50):4.0f} p90 {np.percentile(it,90):4.0f} "
          f"p99 {np.percentile(it,99):4.0f} max {it.max():4d}  solved {out['done'].mean():.4f}  "
          f"rho-updates/problem {out['nup'].mean():.2f}", flush=True)


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    t0 = time.time()
    P = prepare(B, N, seed)
    print(f"prepared B={B} N={N} in {time.time()-t0:.1f}s", flush=True)
    report("current: rho .5, adapt every 25, tol 3", run(P))
    report("adapt at 10,25,50,...", run(P, adapt_at=lambda it: it in (10,) or (it > 0 and it % 25 == 0)))
    report("adapt every 10", run(P, adapt_at=lambda it: it > 0 and it % 10 == 0))
    report("adapt every 10, tol 2", run(P, adapt_at=lambda it: it > 0 and it % 10 == 0, tol=2.0))
    report("adapt every 15", run(P, adapt_at=lambda it: it > 0 and it % 15 == 0))
    report("adapt at 5,15,30,50,75,...", run(P, adapt_at=lambda it: it in (5, 15, 30, 50) or (it >= 75 and it % 25 == 0)))
