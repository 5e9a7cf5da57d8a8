"""oracle/extensions_oracle.py — float64 numpy restatements of the solver parts that north_star names but the reference
does not contain: 3-D shape matching (K12), XSPH viscosity and vorticity confinement (K13), SDF contacts between rigid bodies in K5.

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


# ---------------------------------------------------------------------------------------------------------------------------------
# K5 with SDF contacts between rigid bodies (ps_set_rigid_body_sdf): float64 all-pairs restatement of the whole contact pass —
# gpu/src/cuda/integration_kernel.cuh:303-462 (collideD / collideCell: neighbour count, Jacobi average, exp(-y) mass scaling,
# friction with its [sic]s) with the pair rule of the reference CPU app's RigidContactConstraint (cpu/src/constraint/
# rigidcontactconstraint.cpp:13-66) for pairs that both carry SDF data, lifted to 3-D.  Independent of the C oracle's traversal.
CLOTH, SOLID = 2, 3
EPS, S_FRICTION, K_FRICTION = 0.001, 0.005, 0.0002


def sdf_pair(si, sj, i_first, r, dist, diam):
    """(depth, normal i -> j) of one SDF contact, or None; si / sj = (gx, gy, gz, depth) world frame, r = x_i - x_j"""
    if si[3] < sj[3] or (si[3] == sj[3] and i_first):
        d, e = si[3], np.array(si[:3], np.float64)
    else:
        d, e = sj[3], -np.array(sj[:3], np.float64)
    if d < diam + EPS:                                   # initBoundary (:13-27)
        d = diam - dist
        if d < EPS:
            return None
        # direction from i to j; the reference's 2-D code takes p1 - p2 (the opposite one), see sdf_contact in ps_neighbor_kernels.cu
        x = -r / dist if dist > EPS else np.array([0.0, 1.0 if i_first else -1.0, 0.0])
        dp = float(x @ e)
        e = x - 2.0 * dp * e if dp < 0 else x
    return d, e


def contact_pass(pos, prev, w, phase, radius, sdf_world=None, same_body_skip=None):
    """one K5 launch over all particles; returns (new positions, contact counts).  same_body_skip(i, j) -> bool for equal phases > SOLID
    (default: always skip, the reference)"""
    pos, prev, w = np.asarray(pos, np.float64)[:, :3], np.asarray(prev, np.float64)[:, :3], np.asarray(w, np.float64)
    n = pos.shape[0]
    cd = radius * 2.001
    out, counts = pos.copy(), np.zeros(n, np.int64)
    scaled = lambda wi, y: 1.0 / ((1.0 / wi) * np.exp(-y)) if wi != 0 else wi
    for i in range(n):
        if phase[i] < CLOTH:
            continue
        nb = []
        for j in range(n):
            if j == i:
                continue
            if phase[i] > SOLID and phase[i] == phase[j] and (same_body_skip is None or same_body_skip(i, j)):
                continue
            if np.linalg.norm(pos[i] - pos[j]) < cd:
                nb.append(j)
        counts[i] = len(nb)
        delta = np.zeros(3)
        for j in nb:
            r = pos[i] - pos[j]
            dist = np.linalg.norm(r)
            both = phase[i] >= SOLID and phase[j] >= SOLID
            cw, cw2 = (scaled(w[i], pos[i, 1]), scaled(w[j], pos[i, 1])) if both else (w[i], w[j])   # [sic] both use y of particle i
            wsum = cw + cw2
            dp = r / dist * ((dist - cd) / wsum)
            fn, fd = r, dist
            if both and sdf_world is not None and sdf_world[i][3] >= 0 and sdf_world[j][3] >= 0:
                c = sdf_pair(sdf_world[i], sdf_world[j], i < j, r, dist, 2 * radius)
                if c is None:
                    continue
                depth, fn = c
                dp = fn * (depth / wsum)
            dp1, dp2 = -cw * dp / len(nb), cw2 * dp / len(nb)
            delta += dp1
            if not both:
                continue
            nf = fn / np.linalg.norm(fn)
            rel = (pos[i] + dp1 - prev[i]) - (prev[i] + dp2 - prev[j])   # [sic] integration_kernel.cuh:447
            t = rel - (rel @ nf) * nf
            lt = np.linalg.norm(t)
            if lt < EPS:
                continue
            if lt < S_FRICTION * fd:
                delta -= t * cw / wsum
            else:
                delta -= t * min(K_FRICTION * fd / lt, 1.0)
        out[i] = pos[i] + delta
    return out, counts
