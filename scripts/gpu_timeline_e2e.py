"""Per-CTA timeline of the solve kernel inside cmpc_solve_host (zero-copy, page-locked buffers), config 2."""
import sys, os, json, time
os.environ["CMPC_DEBUG_TIMELINE"] = "/tmp/cmpc_timeline.bin"
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
mpc = pkg.BatchedMPC(N=N, max_batch=B, warm_mode=0)
for _ in range(4):
    t0 = time.perf_counter(); mpc.solve_host(*hin, want_X=False, out=hout); wall = time.perf_counter() - t0
tl = np.fromfile("/tmp/cmpc_timeline.bin", dtype=np.int64).reshape(-1, 4)
t0 = tl[:, 0].min()
st, en, it = (tl[:, 0] - t0) / 1e3, (tl[:, 1] - t0) / 1e3, tl[:, 3]
span = en.max()
edges = np.linspace(0, span, 21)
occ = [float(np.clip(np.minimum(en, b) - np.maximum(st, a), 0, None).sum() / (b - a) / 148) for a, b in zip(edges[:-1], edges[1:])]
first = np.sort(st)[:888]
print(json.dumps(dict(wall_us_with_debug_sync=wall * 1e6, span_us=float(span), sum_cta_us=float((en - st).sum()), first_wave_start_us=[float(first[0]), float(np.median(first)), float(first[-1])],
                      mean_dur_first_wave=float((en - st)[np.argsort(st)[:888]].mean()))))
print("resident per SM over 20 bins:", [round(o, 2) for o in occ])
