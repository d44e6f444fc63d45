"""GPU experiment (round 2): where the end-to-end time of config 2 goes (cmpc_solve_host on page-locked buffers):
zero-copy (kernels read / write host memory in place) with and without the LPT pass, against staged copies."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
pb = synthetic_batch(4096, N=10, seed=0)
B, N = pb.B, pb.N
pinned = [torch.from_numpy(a).pin_memory() for a in pb.f32()]
hin = [t.numpy() for t in pinned]
pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory().numpy()
hout = (pin((B, N, 12), torch.float32), None, pin((B,), torch.int32), pin((B,), torch.float32), pin((B,), torch.float32), pin((B,), torch.int32))
dargs = [t.cuda() for t in pinned]
for name, opts in (("zero_copy", {}), ("zero_copy_no_lpt", dict(lpt_schedule=0)), ("staged_copies", dict(host_zero_copy=0))):
    mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0, time_kernel=1, **opts)
    for _ in range(5):
        mpc.solve_host(*hin, want_X=False, out=hout)
    t0 = time.perf_counter()
    K = 30
    for _ in range(K):
        mpc.solve_host(*hin, want_X=False, out=hout)
    e2e = (time.perf_counter() - t0) / K * 1e3
    kms = None      # solve_host launches are not kernel-timed
    out = mpc.alloc_outputs(B)
    for _ in range(3):
        mpc.solve(*dargs, out=out); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        mpc.solve(*dargs, out=out)
    torch.cuda.synchronize()
    dev = (time.perf_counter() - t0) / K * 1e3
    print(json.dumps(dict(mode=name, e2e_ms=e2e, e2e_solves_s=B / e2e * 1e3, solve_kernel_ms_in_e2e=kms, device_resident_ms=dev, device_kernel_ms=mpc.last_kernel_ms)), flush=True)
    mpc.close()
