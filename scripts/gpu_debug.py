import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
pb = synthetic_batch(4096, N=10, seed=0)
args = [torch.from_numpy(a).cuda() for a in pb.f32()]
for opts in (dict(adaptive_rho_interval=25), dict(adaptive_rho_interval=25, refresh_every=1)):
    mpc = pkg.BatchedMPC(N=10, max_batch=4096, warm_mode=0, **opts)
    U, X, st = mpc.solve(*args); torch.cuda.synchronize()
    s = st.status.cpu().numpy(); it = st.iters.cpu().numpy(); pri = st.pri_res.cpu().numpy(); dua = st.dua_res.cpu().numpy()
    bad = np.where(s != 1)[0]
    print(opts, 'bad', len(bad))
    for b in bad[:12]:
        print('  ', b, 'status', s[b], 'iters', it[b], 'pri', pri[b], 'dua', dua[b], 'Umax', float(U[b].abs().max()), 'n_stance', int(pb.stance[b].sum()))
