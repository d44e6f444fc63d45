"""Small C-ABI entry points around the solve: event-timed kernel duration, rollout totals,
factorisation-cache errors."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import mpc_b200 as pkg                                   # noqa: E402
from mpc_b200 import _capi                               # noqa: E402
from mpc_b200.problems import synthetic_batch, GAIT_NAMES   # noqa: E402
from mpc_b200.solver import _ptr                         # noqa: E402


def test_last_kernel_ms_brackets_only_the_solve_kernel():
    pb = synthetic_batch(2048, N=10, seed=3)
    args = [torch.from_numpy(a).cuda() for a in pb.f32()]
    mpc = pkg.BatchedMPC(N=10, max_batch=2048, warm_mode=0, time_kernel=1)
    out = mpc.alloc_outputs(2048)
    for _ in range(3):
        mpc.solve(*args, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); mpc.solve(*args, out=out); e1.record(); e1.synchronize()
    k = mpc.last_kernel_ms
    assert 0.0 < k <= e0.elapsed_time(e1) + 1e-3            # the step also holds the two scheduling kernels
    plain = pkg.BatchedMPC(N=10, max_batch=2048, warm_mode=0)
    with pytest.raises(pkg.CmpcError):
        plain.last_kernel_ms                                 # not armed


def test_accumulate_stats_matches_torch_sums():
    pb = synthetic_batch(3000, N=10, gaits=GAIT_NAMES, seed=8)
    args = [torch.from_numpy(a).cuda() for a in pb.f32()]
    mpc = pkg.BatchedMPC(N=10, max_batch=3000, warm_mode=0, max_iter=30)      # some problems hit max_iter
    U, X, st = mpc.solve(*args)
    acc = torch.zeros(3, dtype=torch.int64, device="cuda")
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(2):
        _capi.check(_capi.lib().cmpc_accumulate_stats(mpc._h, 3000, 0, _ptr(st.iters), _ptr(st.status), _ptr(acc), s))
    torch.cuda.synchronize()
    assert int(acc[0]) == 2 * int(st.iters.sum()) and int(acc[1]) == 2 * int((st.status != 1).sum())
    assert int(acc[1]) > 0 and int(acc[2]) == 0             # no cache on this handle


def test_factorisation_cache_is_refused_where_it_is_not_compiled():
    with pytest.raises(pkg.CmpcError):
        pkg.BatchedMPC(N=20, max_batch=4, cache_factorization=1)
    with pytest.raises(pkg.CmpcError):
        pkg.BatchedMPC(N=10, max_batch=4).cache_meta(4)      # handle created without the cache
    pkg.BatchedMPC(N=30, max_batch=4, cache_factorization=1).close()
