import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch
flush = torch.empty(256*1024*1024, dtype=torch.uint8, device="cuda")
pb = synthetic_batch(4096, N=10, seed=0)
args = [torch.from_numpy(a).cuda() for a in pb.f32()]
ref=None
for opts in (dict(), dict(refresh_every=10), dict(refresh_every=8), dict(check_every=10, refresh_every=10), dict(refresh_every=5, check_every=10)):
    mpc = pkg.BatchedMPC(N=10, max_batch=4096, warm_mode=0, **opts)
    out = mpc.alloc_outputs(4096)
    for _ in range(3): mpc.solve(*args, out=out)
    torch.cuda.synchronize(); ts=[]
    for _ in range(10):
        flush.fill_(1); e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); mpc.solve(*args, out=out); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    it=out[2].cpu().numpy(); U=out[0].cpu().numpy(); X=out[1].cpu().numpy()
    if ref is None: ref=(U,X)
    print(opts, f"{np.median(ts):.4f} ms iters mean {it.mean():.1f} max {it.max()} solved {(out[5]==1).float().mean().item():.4f} max|dX| vs default {np.abs(X-ref[1]).max():.2e}", flush=True)
