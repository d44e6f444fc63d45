import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
pb = synthetic_batch(4096, N=10, seed=0).slice(0, 1)
args = [torch.from_numpy(a).cuda() for a in pb.f32()]
for alpha, rho, refresh in ((1.0, 0.05, 1), (1.0, 0.05, 5), (0.5, 0.05, 1), (1.6, 0.3, 1), (1.6, 1.0, 1), (1.0, 0.3, 1)):
  print('alpha', alpha, 'rho', rho, 'refresh', refresh)
  for K in (1, 2, 3, 4, 5, 10, 25, 50, 100):
    mpc = pkg.BatchedMPC(N=10, max_batch=1, warm_mode=0, adaptive_rho_interval=0, rho=rho, alpha=alpha, max_iter=K, check_every=100000, eps_abs=0., eps_rel=0., refresh_every=refresh)
    U, X, st = mpc.solve(*args); torch.cuda.synchronize()
    xw, yw = mpc.get_warm(1)
    print('  ', K, 'max|x| %.4e max|y| %.4e' % (float(U.abs().max()), float(yw.abs().max())), flush=True)
