"""GPU experiment (round 2): the double-buffered host path (cmpc_solve_host_async / cmpc_host_wait) on config 2:
host time per submission, steps per second with 2 and 3 buffer sets in flight, against the blocking call."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
pb = synthetic_batch(4096, N=10, seed=0)
B, N = pb.B, pb.N
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
def bufset():
    return ([torch.from_numpy(a).clone().pin_memory().numpy() for a in pb.f32()],
            (pin((B, N, 12), torch.float32), None, pin((B,), torch.int32), pin((B,), torch.float32), pin((B,), torch.float32), pin((B,), torch.int32)))
bufs = [bufset() for _ in range(3)]
mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0)
K = 60
for _ in range(5):
    mpc.solve_host(*bufs[0][0], want_X=False, out=bufs[0][1])
t0 = time.perf_counter()
for _ in range(K):
    mpc.solve_host(*bufs[0][0], want_X=False, out=bufs[0][1])
blocking = (time.perf_counter() - t0) / K
res = dict(blocking_ms=blocking * 1e3, blocking_solves_s=B / blocking)
for depth in (2, 3):
    for rep in range(2):
        for k in range(4):
            mpc.host_wait(mpc.solve_host_async(*bufs[k % depth][0], out=bufs[k % depth][1]))
        host = []; tickets = []
        t0 = time.perf_counter()
        for k in range(K):
            h0 = time.perf_counter()
            tickets.append(mpc.solve_host_async(*bufs[k % depth][0], out=bufs[k % depth][1]))
            host.append(time.perf_counter() - h0)
            if k >= depth - 1:
                mpc.host_wait(tickets[k - (depth - 1)])
        for tk in tickets[-(depth - 1):]:
            mpc.host_wait(tk)
        dt = (time.perf_counter() - t0) / K
        res[f"depth{depth}_run{rep}"] = dict(ms_per_step=dt * 1e3, solves_s=B / dt, host_us_per_submit_p50=float(np.median(host)) * 1e6, host_us_per_submit_max=float(np.max(host)) * 1e6)
print(json.dumps(res))
