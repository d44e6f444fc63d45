"""Per-CTA timeline of one solve-kernel launch (CMPC_DEBUG_TIMELINE): when each problem starts and ends, on which
SM, with how many iterations.  Prints the makespan, the SM occupancy over time and the critical CTAs."""
import sys, os, json
os.environ["CMPC_DEBUG_TIMELINE"] = "/tmp/cmpc_timeline.bin"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mpc_b200 as pkg
from mpc_b200.problems import synthetic_batch, GAIT_NAMES
which = sys.argv[1] if len(sys.argv) > 1 else "c2"
pb = synthetic_batch(4096, N=10, seed=0) if which == "c2" else synthetic_batch(8192, N=10, gaits=GAIT_NAMES, seed=0, mu=(0.3, 1.0))
opts = json.loads(sys.argv[2]) if len(sys.argv) > 2 else {}
args = [torch.from_numpy(a).cuda() for a in pb.f32()]
mpc = pkg.BatchedMPC(N=10, max_batch=pb.B, warm_mode=0, time_kernel=1, **opts)
out = mpc.alloc_outputs(pb.B)
for i in range(3):
    mpc.solve(*args, out=out); torch.cuda.synchronize()
tl = np.fromfile("/tmp/cmpc_timeline.bin", dtype=np.int64).reshape(-1, 4)
t0 = tl[:, 0].min()
st, en, sm, it = (tl[:, 0] - t0) / 1e3, (tl[:, 1] - t0) / 1e3, tl[:, 2], tl[:, 3]     # us
span = en.max()
print(json.dumps(dict(cfg=which, opts=opts, kernel_ms=mpc.last_kernel_ms, span_us=float(span), n_sm=int(len(np.unique(sm))),
                      sum_cta_us=float((en - st).sum()), mean_resident=float((en - st).sum() / span / 148))))
edges = np.linspace(0, span, 21)
occ = [float(np.clip(np.minimum(en, b) - np.maximum(st, a), 0, None).sum() / (b - a) / 148) for a, b in zip(edges[:-1], edges[1:])]
print("resident CTAs per SM over 20 time bins:", [round(o, 2) for o in occ])
last_start = [float(st[(st >= a) & (st < b)].size) for a, b in zip(edges[:-1], edges[1:])]
print("CTA starts per bin:", last_start)
order = np.argsort(-en)[:12]
print("last CTAs to finish (launch index, start us, end us, dur us, iters, us/iter after 27 us of setup):")
for i in order:
    print(f"  blk {i:5d} sm {sm[i]:3d} start {st[i]:7.1f} end {en[i]:7.1f} dur {en[i]-st[i]:7.1f} iters {it[i]:4d}  {(en[i]-st[i]-27)/max(it[i],1):.3f}")
# duration model: dur = a + b * iters for CTAs started in the first / last third
for name, sel in (("started in first 30%", st < 0.3 * span), ("started after 60%", st > 0.6 * span)):
    if sel.sum() > 10:
        A = np.stack([np.ones(sel.sum()), it[sel]], 1)
        c = np.linalg.lstsq(A, (en - st)[sel], rcond=None)[0]
        print(f"{name}: n={int(sel.sum())} dur ~ {c[0]:.1f} us + {c[1]:.3f} us/iter  (cycles at 1.965 GHz: {c[0]*1965:.0f} + {c[1]*1965:.0f}/iter)")
# slot turnover: for every CTA that starts after t = 0, the time since the latest earlier CTA end on the same SM
gaps = []
for s in np.unique(sm):
    m = sm == s
    ends = np.sort(en[m])
    for t in np.sort(st[m]):
        if t > 1.0:
            j = np.searchsorted(ends, t) - 1
            if j >= 0:
                gaps.append(t - ends[j])
gaps = np.array(gaps)
print("slot turnover (CTA start - latest CTA end on that SM), us: p10 %.2f p50 %.2f p90 %.2f mean %.2f n %d" % (
    np.percentile(gaps, 10), np.percentile(gaps, 50), np.percentile(gaps, 90), gaps.mean(), gaps.size))
# resident count seen by each SM at a few instants
for frac in (0.3, 0.5, 0.7):
    t = frac * span
    cnt = np.bincount(sm[(st <= t) & (en > t)].astype(int), minlength=148)
    print(f"t = {t:.0f} us: resident per SM histogram", np.bincount(cnt, minlength=9).tolist())
np.save(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", f"timeline_{which}.npy"), tl)
