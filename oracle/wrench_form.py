"""ORACLE (test infrastructure, not product code).

Wrench-space factorisation of the condensed QP that the CUDA path relies on, restated
in numpy fp64 so tests can check it against the plain recursion of
:mod:`oracle.srbd_qp` (which follows reference ``src/mpc.py:64-136`` literally).

With tau^_j = Rz I^-1 [r_jl]x f_l summed over stance legs and a_j = sum f_l / m, the
horizon cost depends on the forces only through the 6N wrench sequence w = G u:

    H = G' M G,   M = blockdiag over the 6 axes (Theta_x,y,z ; p_x,y,z) of N x N matrices
    M_a[j,j'] = 2 * sum_{k>max(j,j')}^{N} ( w_pos,a d^4 (k-1-j)(k-1-j') + w_vel,a d^2 )

valid when the angular-velocity weights of x and y agree (reference: 1e4, 1e4,
src/mpc.py:128-129), because ||Rz' v||_W = ||v||_W then.
"""
import numpy as np
from . import srbd_qp


def axis_gram(N, delta, w=srbd_qp.W_STATE):
    """M (6,N,N): axes 0-2 rotated angular (weights w[0:3] on Theta, w[6:9] on omega),
    axes 3-5 linear (w[3:6] on p, w[9:12] on v)."""
    assert w[6] == w[7], "wrench form needs isotropic xy angular-velocity weights"
    M = np.zeros((6, N, N))
    k = np.arange(1, N + 1)
    for a in range(6):
        wp, wv = (w[a], w[6 + a]) if a < 3 else (w[a], w[6 + a])
        for j in range(N):
            for jp in range(N):
                kk = k[k > max(j, jp)]
                M[a, j, jp] = 2.0 * np.sum(wp * delta ** 4 * (kk - 1 - j) * (kk - 1 - jp)
                                           + wv * delta ** 2)
    return M


def leg_maps(x0, r, mass=srbd_qp.MASS, ibody_inv=srbd_qp.IBODY_INV):
    """Ghat (N,4,3,3) = Rz I^-1_hat [r]x  (torque rows of G in rotated coordinates)."""
    Rz = srbd_qp.rot_z(x0[2])
    Ihat = Rz @ np.diag(ibody_inv) @ Rz.T
    N = r.shape[0]
    Gh = np.zeros((N, 4, 3, 3))
    for j in range(N):
        for l in range(4):
            Gh[j, l] = Rz @ Ihat @ srbd_qp.skew(r[j, l])
    return Gh


def G_matrix(x0, r, stance, mass=srbd_qp.MASS):
    """G (6N x n) on the compact stance unknowns; wrench index = 6*j + a."""
    N = r.shape[0]
    idx = srbd_qp.stance_index(stance)
    Gh = leg_maps(x0, r)
    G = np.zeros((6 * N, 3 * len(idx)))
    for s, (j, l) in enumerate(idx):
        G[6 * j:6 * j + 3, 3 * s:3 * s + 3] = Gh[j, l]
        G[6 * j + 3:6 * j + 6, 3 * s:3 * s + 3] = np.eye(3) / mass
    return G


def M_full(N, delta, w=srbd_qp.W_STATE):
    """6N x 6N with wrench index 6*j + a."""
    Ma = axis_gram(N, delta, w)
    M = np.zeros((6 * N, 6 * N))
    for a in range(6):
        M[a::6, a::6] = Ma[a]
    return M
