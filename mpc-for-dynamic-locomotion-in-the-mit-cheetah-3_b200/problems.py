"""Batched gait plans and synthetic MPC workloads (numpy, fp64, vectorised over robots).

``BatchedGaitPlan`` is the array-of-robots form of :class:`gait.GaitPlan`: every robot
may have its own gait (first_swing, ss, ds), reference velocity and yaw rate, so the
unicycle integration of reference ``src/footstep_planner.py:72-177`` is carried out for
all robots at once.  ``synthetic_batch`` draws the workloads SURVEY.md section 8(d)
defines for BASELINE.json's configs 2-4 (randomised CoM states / velocity references
around a walking Lite3) and assembles the per-problem QP parameters exactly as the
reference's ``MPC.solve`` would (``src/mpc.py:178-255``).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .gait import GAITS, stance_bits

# pkl tick 0 of the reference's logged run
NOMINAL_FEET = np.array([[0.10629492, 0.1605, 0.01713467],
                         [0.10629492, -0.1605, 0.01713467],
                         [-0.24270508, 0.1605, 0.01713467],
                         [-0.24270508, -0.1605, 0.01713467]])
NOMINAL_COM = np.array([-0.0101835626879727, -0.00027996234635044087, 0.285])
GRAVITY = -9.81
DT = 0.01
_SGN_T = np.array([+1.0, +1.0, -1.0, -1.0])    # torso_displacement/2 sign, FL FR HL HR
_SGN_L = np.array([-1.0, +1.0, -1.0, +1.0])    # leg_displacement_y sign


@dataclass
class BatchedGaitPlan:
    pos: np.ndarray        # (B,S,4,3)
    feet_id: np.ndarray    # (B,S,4) stance mask of each step's single-support part
    ss: np.ndarray         # (B,) int
    ds: np.ndarray         # (B,) int
    step_height: float = 0.08

    @classmethod
    def build(cls, feet0, yaw0, first_swing, ss, ds, v_ref, omega_ref, total_steps=20,
              dt=DT, step_height=0.08) -> "BatchedGaitPlan":
        """feet0 (B,4,3) or (4,3); yaw0 (B,); first_swing (B,4); ss, ds (B,) ints;
        v_ref (B,3); omega_ref (B,)."""
        first_swing = np.atleast_2d(np.asarray(first_swing)).astype(np.int64)
        B = first_swing.shape[0]
        feet0 = np.broadcast_to(np.asarray(feet0, dtype=float), (B, 4, 3))
        ss = np.broadcast_to(np.asarray(ss, dtype=np.int64), (B,)).copy()
        ds = np.broadcast_to(np.asarray(ds, dtype=np.int64), (B,)).copy()
        v_ref = np.broadcast_to(np.asarray(v_ref, dtype=float), (B, 3))
        omega_ref = np.broadcast_to(np.asarray(omega_ref, dtype=float), (B,))
        theta = np.broadcast_to(np.asarray(yaw0, dtype=float), (B,)).copy()
        S = int(total_steps)
        # same arithmetic, in the same order, as reference src/footstep_planner.py:72-177 (and
        # gait.GaitPlan.from_initial): the 2x2 rotations go through np.matmul like the reference's
        # `Rm @ v`, so that the plan of every robot agrees with the reference planner to the last bit
        fl, fr, hl, hr = feet0[:, 0], feet0[:, 1], feet0[:, 2], feet0[:, 3]
        uni = (hl + hr + fl + fr) / 4.0                   # (B,3)
        dfl = fl[:, :2] - hl[:, :2]
        dlat = hr[:, :2] - hl[:, :2]

        def rot(th):
            c, s = np.cos(th), np.sin(th)
            return np.stack([np.stack([c, -s], -1), np.stack([s, c], -1)], -2)      # (B,2,2)

        def apply(Rm, v):                                  # per-robot Rm @ v (stacked matmul: same BLAS
            return np.matmul(Rm, v[:, :, None])[:, :, 0]   # kernel, same bits as the reference's 2x2 `@`)

        pos = np.zeros((B, S, 4, 3))
        feet_id = np.ones((B, S, 4), dtype=np.int64)
        support = first_swing.copy()
        period = ss + ds
        Rm = rot(theta)
        for j in range(S):
            if j >= 1:
                for i in range(int(period.max())):
                    live = i < period
                    theta = np.where(live, theta + omega_ref * dt, theta)
                    Rm = np.where(live[:, None, None], rot(theta), Rm)
                    step = apply(Rm, v_ref[:, :2]) * dt
                    uni[:, :2] = np.where(live[:, None], uni[:, :2] + step, uni[:, :2])
            torso = apply(Rm, dfl)
            lat = apply(Rm, dlat) / 2.0
            half = torso / 2
            base = uni[:, None, :2] + _SGN_T[None, :, None] * half[:, None, :]     # u +- t/2 (exact sign flip)
            new = np.empty((B, 4, 3))
            new[:, :, :2] = base + _SGN_L[None, :, None] * lat[:, None, :]          # (u +- t/2) +- lat
            new[:, :, 2] = uni[:, None, 2]
            if j >= 1:
                keep = support.astype(bool)[:, :, None]
                pos[:, j] = np.where(keep, pos[:, j - 1], new)
                feet_id[:, j] = support
                support = 1 - support
            else:
                pos[:, j] = new
        return cls(pos, feet_id, ss, ds, step_height)

    @property
    def n_steps(self):
        return self.pos.shape[1]

    def step_index(self, t):
        """t (B,) or (B,K) ticks -> step index, reference src/footstep_planner.py:226-231."""
        t = np.asarray(t)
        period = (self.ss + self.ds).reshape((-1,) + (1,) * (t.ndim - 1))
        return np.minimum(t // period, self.n_steps - 1)

    def stance_mask(self, t):
        """(B,K) ticks -> (B,K,4) stance mask, reference src/footstep_planner.py:239-246."""
        t = np.asarray(t)
        shp = (-1,) + (1,) * (t.ndim - 1)
        step = self.step_index(t)
        tin = t - step * (self.ss + self.ds).reshape(shp)
        single = tin < self.ss.reshape(shp)
        b = np.arange(t.shape[0]).reshape(shp)
        return np.where(single[..., None], self.feet_id[b, step], 1)

    def foot_position(self, t):
        """(B,K) ticks -> (B,K,4,3) look-ahead foot position (reference src/mpc.py:306-318,
        src/foot_trajectory_generator.py:27-96)."""
        t = np.asarray(t)
        shp = (-1,) + (1,) * (t.ndim - 1)
        step = self.step_index(t)
        b = np.arange(t.shape[0]).reshape(shp)
        tin = (t - step * (self.ss + self.ds).reshape(shp)).astype(float)[..., None, None]
        start = self.pos[b, step]
        target = self.pos[b, np.minimum(step + 1, self.n_steps - 1)]
        ts = (0.80 * self.ss.astype(float)).reshape(shp)[..., None, None]
        out = start + (target - start) * (-2 / ts ** 3 * tin ** 3 + 3 / ts ** 2 * tin ** 2)
        h = self.step_height
        tz, tsz = tin[..., 0], ts[..., 0]
        out[..., 2] = (16 * h / tsz ** 4 * tz ** 4 - 32 * h / tsz ** 3 * tz ** 3
                       + 16 * h / tsz ** 2 * tz ** 2 + start[..., 2])
        out = np.where(tin >= ts, target, out)
        out = np.where((step == 0)[..., None, None], start, out)
        stance = self.stance_mask(t).astype(bool)[..., None]
        return np.where(stance, start, out)


@dataclass
class ProblemBatch:
    """fp64 problem data in the layout of the C ABI."""
    x0: np.ndarray        # (B,13)
    r: np.ndarray         # (B,N,4,3)
    stance: np.ndarray    # (B,N,4) {0,1}
    x_des: np.ndarray     # (B,N+1,13)
    mu: np.ndarray        # (B,)
    gait_id: np.ndarray   # (B,)
    tick: np.ndarray      # (B,)

    @property
    def B(self):
        return self.x0.shape[0]

    @property
    def N(self):
        return self.r.shape[1]

    @property
    def mask_bits(self):
        return stance_bits(self.stance)

    def f32(self):
        """(x0, r, mask, x_des, mu) as contiguous fp32/uint8 numpy arrays."""
        c = np.ascontiguousarray
        return (c(self.x0, dtype=np.float32), c(self.r, dtype=np.float32), c(self.mask_bits),
                c(self.x_des, dtype=np.float32), c(self.mu, dtype=np.float32))

    def problem(self, b):
        """(x0, r, stance, x_des(13,N+1), mu) of problem b for the oracle."""
        return self.x0[b], self.r[b], self.stance[b], self.x_des[b].T, float(self.mu[b])

    def slice(self, lo, hi):
        return ProblemBatch(self.x0[lo:hi], self.r[lo:hi], self.stance[lo:hi], self.x_des[lo:hi],
                            self.mu[lo:hi], self.gait_id[lo:hi], self.tick[lo:hi])


GAIT_NAMES = ("trot", "pronk", "amble", "pseudo_gallop")


def synthetic_batch(B, N=10, gaits=("trot",), seed=0, mu=(1.0, 1.0), tick_range=(20, 380),
                    noise=1.0, total_steps=20, tick_shift=0) -> ProblemBatch:
    """Workload of BASELINE.json configs 2-4: robot b walks with gait
    ``gaits[b % len(gaits)]``... drawn uniformly; tick ~ U{tick_range}; velocity references
    v_x ~ U[-0.3,0.3], v_y ~ U[-0.1,0.1], yaw rate ~ U[-0.5,0.5]; measured state = reference
    state at that tick + N(0, sigma) noise (rpy 0.05 rad, com 0.02 m, omega 0.2 rad/s,
    v 0.1 m/s); stage-0 feet = planned feet + N(0, 0.005 m); mu ~ U[mu].
    ``tick_shift`` moves every robot's tick by that many ticks with all random draws unchanged
    (``tick_shift=-1`` is "the same problems one tick earlier", the warm-start source of config 2)."""
    rng = np.random.default_rng(seed)
    gid = rng.integers(0, len(gaits), size=B)
    table = [GAITS[g] for g in gaits]
    first_swing = np.array([table[i][0] for i in gid], dtype=np.int64)
    ss = np.array([table[i][1] for i in gid], dtype=np.int64)
    ds = np.array([table[i][2] for i in gid], dtype=np.int64)
    tick = rng.integers(tick_range[0], tick_range[1], size=B) + int(tick_shift)
    v_ref = np.stack([rng.uniform(-0.3, 0.3, B), rng.uniform(-0.1, 0.1, B), np.zeros(B)], 1)
    om_ref = rng.uniform(-0.5, 0.5, B)
    mu_b = rng.uniform(mu[0], mu[1], B)
    plan = BatchedGaitPlan.build(NOMINAL_FEET, np.zeros(B), first_swing, ss, ds, v_ref, om_ref,
                                 total_steps=total_steps)
    # reference (src/mpc.py:181-183): references are zeroed during the last planned step
    last = plan.step_index(tick) == total_steps - 1
    v_use = np.where(last[:, None], 0.0, v_ref)
    om_use = np.where(last, 0.0, om_ref)
    # reference accumulators after `tick` solves (src/mpc.py:261-262), all before the last step
    yaw_start = om_ref * DT * tick
    com_start = NOMINAL_COM[None] + v_ref * DT * tick[:, None]
    k = np.arange(N + 1)
    x_des = np.zeros((B, N + 1, 13))
    x_des[:, :, 2] = yaw_start[:, None] + om_use[:, None] * DT * k
    x_des[:, :, 3:6] = com_start[:, None, :] + v_use[:, None, :] * DT * k[None, :, None]
    x_des[:, :, 8] = om_use[:, None]
    x_des[:, :, 9:12] = v_use[:, None, :]
    x_des[:, :, 12] = GRAVITY
    x0 = x_des[:, 0, :].copy()
    x0[:, 0:3] += noise * rng.normal(0, 0.05, (B, 3))
    x0[:, 3:6] += noise * rng.normal(0, 0.02, (B, 3))
    x0[:, 6:9] += noise * rng.normal(0, 0.2, (B, 3))
    x0[:, 9:12] += noise * rng.normal(0, 0.1, (B, 3))
    ticks = tick[:, None] + np.arange(N)[None]
    feet = plan.foot_position(ticks)                                   # (B,N,4,3)
    feet[:, 0] += noise * rng.normal(0, 0.005, (B, 4, 3))
    r = np.empty((B, N, 4, 3))
    r[:, 0] = feet[:, 0] - x0[:, None, 3:6]                            # measured (src/mpc.py:223-226)
    r[:, 1:] = feet[:, 1:] - x_des[:, 1:N, None, 3:6]                  # planned - desired com
    stance = plan.stance_mask(ticks)
    return ProblemBatch(x0, r, stance, x_des, mu_b, gid, tick)
