"""GPU diagnostic: how close does the fp32 CUDA ADMM get to the tight fp64 optimum (oracle/tight_ipm)
on the unique quantities (X, J, per-stage wrench) for each BASELINE config, as a function of the
iteration budget.  Prints one JSON line per (config, budget)."""
import json, sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, DT, GAIT_NAMES
from oracle import srbd_qp, tight_ipm as ipm

def run(name, pb, sel, budgets, **opts):
    refs = {}
    for b in sel:
        x0, r, st, xd, mu = pb.problem(b)
        f = lambda a: np.float32(a).astype(np.float64)
        refs[b] = ipm.solve_problem(f(x0), f(r), st, f(xd), float(np.float32(mu)), DT)
    dev = torch.device("cuda", 0)
    sub = [torch.from_numpy(a[sel]).to(dev) for a in pb.f32()]
    for K, eps in budgets:
        mpc = pkg.BatchedMPC(N=pb.N, max_batch=len(sel), max_iter=K, eps_abs=eps, eps_rel=eps, warm_mode=0,
                             check_every=5 if eps > 0 else 25, kernel_variant=int(os.environ.get("DIAG_VARIANT", "0")), **opts)
        U, X, st = mpc.solve(*sub)
        torch.cuda.synchronize()
        U = U.cpu().numpy().astype(np.float64); X = X.cpu().numpy().astype(np.float64)
        it = st.iters.cpu().numpy(); stat = st.status.cpu().numpy()
        eX, eW, eWn, eJ = [], [], [], []
        for i, b in enumerate(sel):
            ref = refs[b]
            x0, r, stn, xd, mu = pb.problem(b)
            W = srbd_qp.stage_wrench(U[i], np.float32(r).astype(np.float64))
            dW = np.abs(W - ref["wrench"])
            eW.append(dW.max())
            eWn.append((dW / (1e-2 + 1e-3 * np.abs(ref["wrench"]))).max())
            eX.append(np.abs(X[i].T - ref["X"]).max())
            eJ.append(abs(srbd_qp.objective(X[i].T, np.float32(xd).astype(np.float64)) / ref["J"] - 1))
        q = lambda v: [float(np.percentile(v, p)) for p in (50, 90, 100)]
        print(json.dumps(dict(cfg=name, K=K, eps=eps, opts=opts, iters=q(it), bad=int((stat < 0).sum()),
                              dX=q(eX), dW=q(eW), dW_over_tol=q(eWn), dJ=q(eJ))), flush=True)
        mpc.close()

if __name__ == "__main__":
    rng = np.random.default_rng(0)
    which = sys.argv[1:] or ["c2", "c3", "c4", "n60"]
    B = [(1000, 1e-3), (200, 0.0), (1000, 0.0), (3000, 0.0), (10000, 0.0)]
    if os.environ.get("DIAG_BUDGETS"):
        B = [(int(k), float(e)) for k, e in (t.split(":") for t in os.environ["DIAG_BUDGETS"].split(","))]
    if "c2" in which:
        pb = synthetic_batch(4096, N=10, seed=0)
        run("config2", pb, rng.choice(4096, 32, replace=False), B)
    if "c3" in which:
        pb = synthetic_batch(8192, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0))
        run("config3", pb, rng.choice(8192, 32, replace=False), B)
    if "c4" in which:
        pb = synthetic_batch(2048, N=30, seed=0)
        run("config4", pb, rng.choice(2048, 24, replace=False), B)
    if "n60" in which:
        pb = synthetic_batch(64, N=60, gaits=("pseudo_gallop", "trot"), seed=0)
        run("n60", pb, np.arange(8), B)
    if "eps" in which:       # which finite tolerances does the fp32 kernel certify, and how tight is the result
        for name, pb in (("config2", synthetic_batch(4096, N=10, seed=0)), ("config4", synthetic_batch(2048, N=30, seed=0)),
                         ("n60", synthetic_batch(64, N=60, gaits=("pseudo_gallop", "trot"), seed=0))):
            sel = rng.choice(pb.B, 24, replace=False) if pb.B > 64 else np.arange(8)
            run(name + "_eps", pb, sel, [(100000, 1e-4), (100000, 1e-5), (100000, 1e-6), (100000, 3e-7)])
    if "n20" in which:
        pb = synthetic_batch(64, N=20, gaits=GAIT_NAMES, seed=0)
        run("n20", pb, np.arange(12), B)
    if "t20" in which:       # the batch of tests/test_gpu_tight_parity.py::test_tight_parity_other_horizons[20]
        pb = synthetic_batch(64, N=20, gaits=GAIT_NAMES, seed=6, mu=(0.3, 1.0))
        tol = os.environ.get("DIAG_RHO_TOL")
        run("t20", pb, np.arange(12), B, **({"adaptive_rho_tolerance": float(tol)} if tol else {}))
    if "t30" in which:       # the sample of test_tight_parity_config4_long_horizon
        pb = synthetic_batch(16384, N=30, seed=0)
        run("t30", pb, np.random.default_rng(4).choice(pb.B, 24, replace=False), B)
