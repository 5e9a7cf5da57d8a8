"""oracle/extensions_oracle.py — float64 numpy restatements of the solver parts that north_star names but the reference
does not contain: 3-D shape matching (K12), XSPH viscosity and vorticity confinement (K13).

TEST INFRASTRUCTURE ONLY (tests/test_gpu_extensions.py).  PARITY UNPINNED: there is no reference implementation of
these — the reference GPU solver's rigid_body_functor is an empty stub (gpu/src/cuda/solver_kernel.cuh:289-312) and
XSPH / vorticity exist nowhere in it — so these follow the papers the reference implements:
  shape matching          Mueller et al. 2005 / Macklin et al. 2014 §5.1: goal_i = c + R r_i, R = polar(sum m (x-c) r^T)
  XSPH, vorticity         Macklin & Mueller 2013, eqs. 15-17, with the solver's own kernels (gpu/src/cuda/integration_kernel.cuh:20-35)
The rotation here comes from an SVD (U diag(1,1,det) V^T), independent of the kernel's quaternion iteration."""
import numpy as np

H, H2 = 2.0, 4.0
POLY6 = 0.00305992474   # integration_kernel.cuh: 315 / (64 pi H^9)
SPIKY = 0.22381163872   # 45 / (pi H^6)


def polar_rotation(A):
    U, _, Vt = np.linalg.svd(A)
    d = np.sign(np.linalg.det(U @ Vt))
    return U @ np.diag([1.0, 1.0, d]) @ Vt


def shape_match(x, rest, mass, stiffness=1.0):
    """x: current positions (n,3); rest: rest offsets from the rest centre of mass; returns (new x, R, c)"""
    x, rest, mass = np.asarray(x, np.float64), np.asarray(rest, np.float64), np.asarray(mass, np.float64)
    c = (mass[:, None] * x).sum(0) / mass.sum()
    A = ((mass[:, None] * (x - c))[:, :, None] * rest[:, None, :]).sum(0)
    R = polar_rotation(A)
    goal = c + rest @ R.T
    return x + stiffness * (goal - x), R, c


def quat_to_mat(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def viscosity(pos, vel, fluid, c_xsph, vorticity_eps, dt):
    """brute force over all pairs with |r| < H; returns the new velocities (non-fluid particles unchanged)"""
    pos, vel = np.asarray(pos, np.float64)[:, :3], np.asarray(vel, np.float64)[:, :3]
    n = pos.shape[0]
    r = pos[:, None, :] - pos[None, :, :]            # r_ij = p_i - p_j
    r2 = (r * r).sum(-1)
    nb = (r2 < H2) & ~np.eye(n, dtype=bool) & fluid[:, None] & np.ones(n, bool)[None, :]
    rl = np.sqrt(np.where(nb, r2, 1.0))
    ok = nb & (rl >= 1e-4)
    dv_ij = vel[None, :, :] - vel[:, None, :]         # v_j - v_i
    # grad_pj W(p_i - p_j) = +SPIKY (H - r)^2 r/|r|
    g = np.where(ok[:, :, None], (SPIKY * (H - rl) ** 2 / rl)[:, :, None] * r, 0.0)
    omega = np.cross(dv_ij, g).sum(1)
    om_len = np.linalg.norm(omega, axis=1)
    W = np.where(nb, POLY6 * (H2 - r2) ** 3, 0.0)
    xs = (dv_ij * W[:, :, None]).sum(1)
    eta = (-(g) * om_len[None, :, None]).sum(1)      # sum_j |omega_j| grad_pi W
    en = np.linalg.norm(eta, axis=1)
    dv = c_xsph * xs
    if vorticity_eps != 0.0:
        N = np.where((en > 1e-6)[:, None], eta / np.maximum(en, 1e-30)[:, None], 0.0)
        dv = dv + vorticity_eps * dt * np.cross(N, omega)
    out = vel.copy()
    out[fluid] += dv[fluid]
    return out, omega
