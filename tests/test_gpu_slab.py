"""Slab decomposition on the GPU: the ps_slab_* kernels (ordered halo packing, ghost unpacking, migration with stable
compaction) driven by particlesolver_b200.slab, several slabs as several contexts on one B200 (in-process hand-over of
the record buffers), against (a) one undecomposed context and (b) the CPU oracle engine doing the same decomposition.
The multi-process NCCL path is in test_gpu_slab_nccl.py."""
import numpy as np
import pytest
import torch

import helpers as H
import oracle_py as orc
import particlesolver_b200 as psb
from particlesolver_b200 import slab
from slab_oracle_engine import HALO_DT, MIGR_DT, OracleEngine
from test_slab_cpu import _match, _scene, _split

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0


def _params(p_or):
    p = psb.default_params()
    p.grid_size[:] = tuple(p_or.grid)
    p.min_bounds[:] = tuple(p_or.min_b)
    p.max_bounds[:] = tuple(p_or.max_b)
    return p


def _ctx_engine(p, arrays, cap=20000):
    pos, vel, w, phase, ros = arrays
    sol = psb.Solver(p, max_particles=cap)
    sol.append(pos, vel, w, ros, phase)
    return slab.CtxEngine(sol, halo_capacity=cap, migrant_capacity=cap)


def _records(t, dt):
    return np.ascontiguousarray(t.cpu().numpy()).reshape(-1).view(dt)


def test_pack_kernels_match_the_cpu_engine():
    """halo and migrant records: same particles, same order, same bytes as the numpy restatement"""
    p_or, pos, vel, w, phase, ros = _scene()
    rng = np.random.default_rng(0)
    pos[:, 0] += rng.uniform(-0.3, 0.3, pos.shape[0]).astype(np.float32)  # ragged faces
    eng = _ctx_engine(_params(p_or), (pos, vel, w, phase, ros))
    ref = OracleEngine(p_or, pos, vel, w, phase, ros)
    x_lo, x_hi = 8.0, 15.0
    for width in (0.7, 4.5):
        gl, gr = eng.pack_halo(x_lo, x_hi, width)
        rl, rr = ref.pack_halo(x_lo, x_hi, width)
        assert gl.shape[0] > 0 and gr.shape[0] > 0
        assert np.array_equal(gl.cpu().numpy(), rl.numpy()) and np.array_equal(gr.cpu().numpy(), rr.numpy())
    # migration: prev/vel travel, the stayers keep their order
    eng.sol.predict(DT)
    ref.pos, ref.prev = eng.sol.download(psb.ARR_POS).copy(), eng.sol.download(psb.ARR_PREV).copy()  # identical inputs for the split
    gl, gr = eng.pack_migrants(x_lo, x_hi)
    rl, rr = ref.pack_migrants(x_lo, x_hi)
    assert gl.shape[0] > 0 and gr.shape[0] > 0
    assert np.array_equal(gl.cpu().numpy(), rl.numpy()) and np.array_equal(gr.cpu().numpy(), rr.numpy())
    assert eng.n_owned == ref.n_owned < pos.shape[0]
    assert np.array_equal(eng.sol.download_owned(psb.ARR_POS), ref.pos)
    assert np.array_equal(eng.sol.download_owned(psb.ARR_PREV), ref.prev)
    assert np.array_equal(eng.sol.download_owned(psb.ARR_VEL), ref.vel)
    # and back in: appended behind the stayers, left neighbour's first
    eng.append_migrants(gr, gl); ref.append_migrants(rr, rl)
    assert eng.n_owned == pos.shape[0]
    assert np.array_equal(eng.sol.download_owned(psb.ARR_VEL), ref.vel)
    assert np.array_equal(eng.sol.download_owned(psb.ARR_PHASE), ref.phase)
    eng.sol.close()


def test_lambda_exchange_kernels():
    """ps_slab_pack_lambda answers the halo pack record by record (lambda lives by sorted slot, the records by particle);
    ps_slab_set_ghost_lambda drops the received values into the ghosts' sorted slots, left neighbour's first"""
    p_or, pos, vel, w, phase, ros = _scene()
    rng = np.random.default_rng(1)
    pos[:, 0] += rng.uniform(-0.3, 0.3, pos.shape[0]).astype(np.float32)
    eng = _ctx_engine(_params(p_or), (pos, vel, w, phase, ros))
    sol = eng.sol
    n = pos.shape[0]
    with pytest.raises(psb.PsError) as e:   # nothing to answer yet
        eng.pack_lambda()
    assert e.value.code == psb.PS_ERR_STATE
    x_lo, x_hi, width = 8.0, 15.0, 2.25
    gl, gr = eng.pack_halo(x_lo, x_hi, width)
    nl, nr = gl.shape[0], gr.shape[0]
    assert nl > 0 and nr > 0
    # pretend the particle's own halo records come back as ghosts (right buffer as the left neighbour's, and vice versa)
    eng.set_ghosts(gr, gl)
    assert sol.n == n + nl + nr
    sol.build_grid()
    sol.solve_fluid_lambda()
    index = sol.download(psb.ARR_INDEX)
    lam_sorted = sol.download(psb.ARR_LAMBDA)
    lam = np.empty(n + nl + nr, np.float32); lam[index] = lam_sorted       # by particle
    ll, lr = eng.pack_lambda()
    sol.sync()   # the pack runs on the solver's stream and reports no counts, so nothing has waited for it yet
    assert (ll.shape[0], lr.shape[0]) == (nl, nr)
    x = pos[:, 0]
    sel_l, sel_r = x < np.float32(x_lo + width), x >= np.float32(x_hi - width)
    assert np.array_equal(ll.cpu().numpy().view(np.float32).ravel(), lam[:n][sel_l])
    assert np.array_equal(lr.cpu().numpy().view(np.float32).ravel(), lam[:n][sel_r])
    # hand recognisable values to the ghosts
    fl = torch.arange(nr, dtype=torch.float32, device="cuda").add_(1000.0).view(torch.uint8).reshape(-1, 4)
    fr = torch.arange(nl, dtype=torch.float32, device="cuda").add_(5000.0).view(torch.uint8).reshape(-1, 4)
    eng.set_ghost_lambda(fl, fr)
    got = np.empty_like(lam); got[index] = sol.download(psb.ARR_LAMBDA)
    assert np.array_equal(got[:n], lam[:n])                                # owned lambdas untouched
    assert np.array_equal(got[n:n + nr], 1000.0 + np.arange(nr, dtype=np.float32))
    assert np.array_equal(got[n + nr:], 5000.0 + np.arange(nl, dtype=np.float32))
    with pytest.raises(psb.PsError):                                       # one value per ghost, no more, no fewer
        sol.slab_set_ghost_lambda(fl.data_ptr(), nr - 1, fr.data_ptr(), nl)
    sol.close()


def test_lambda_sinks_fill_the_messages_from_the_lambda_pass():
    """ps_slab_set_lambda_sinks: the fused lambda kernel writes the lambda of every halo member into the outgoing messages itself, and
    ps_slab_pack_lambda on the same buffers only reports the counts — same bytes as the separate pack pass; other buffers, the queue
    walk or a cleared sink fall back to the pack kernel"""
    p_or, pos, vel, w, phase, ros = _scene()
    rng = np.random.default_rng(2)
    pos[:, 0] += rng.uniform(-0.3, 0.3, pos.shape[0]).astype(np.float32)
    assert (phase == 0).all()                                              # all fluid: what the sinks are for
    eng = _ctx_engine(_params(p_or), (pos, vel, w, phase, ros))
    sol = eng.sol
    x_lo, x_hi, width = 8.0, 15.0, 2.25
    gl, gr = eng.pack_halo(x_lo, x_hi, width)
    nl, nr = gl.shape[0], gr.shape[0]
    eng.set_ghosts(gr, gl)
    sol.build_grid()
    sol.solve_fluid_lambda()
    ll, lr = eng.pack_lambda()
    sol.sync()
    want_l, want_r = ll.cpu().numpy().copy(), lr.cpu().numpy().copy()      # the pack kernel's messages
    assert nl > 0 and nr > 0 and want_l.shape[0] == nl
    # now as sinks: poison the buffers, run the lambda pass, "pack" without a launch
    for buf in eng.lam_send:
        buf.fill_(0xFF)
    sol.slab_set_lambda_sinks(eng.lam_send[0].data_ptr(), eng.lam_send[1].data_ptr(), eng.halo_cap)
    sol.solve_fluid_lambda()
    sol.sync()
    assert np.array_equal(eng.lam_send[0][:nl].cpu().numpy(), want_l) and np.array_equal(eng.lam_send[1][:nr].cpu().numpy(), want_r)
    assert (eng.lam_send[0][nl:nl + 4].cpu().numpy() == 0xFF).all()        # nothing beyond the records
    ll2, lr2 = eng.pack_lambda()                                           # counts only; the bytes stay
    sol.sync()
    assert (ll2.shape[0], lr2.shape[0]) == (nl, nr) and np.array_equal(ll2.cpu().numpy(), want_l) and np.array_equal(lr2.cpu().numpy(), want_r)
    # a different destination is packed by the kernel as before
    other = [torch.full_like(b, 0xEE) for b in eng.lam_send]
    sol.solve_fluid_lambda()
    c = sol.slab_pack_lambda(other[0].data_ptr(), other[1].data_ptr(), eng.halo_cap)
    sol.sync()
    assert c == (nl, nr) and np.array_equal(other[0][:nl].cpu().numpy(), want_l) and np.array_equal(other[1][:nr].cpu().numpy(), want_r)
    # sinks cleared: the lambda pass leaves the buffers alone
    sol.slab_set_lambda_sinks(None, None, 0)
    for buf in eng.lam_send:
        buf.fill_(0xFF)
    sol.solve_fluid_lambda()
    sol.sync()
    assert (eng.lam_send[0][:nl].cpu().numpy() == 0xFF).all()
    ll3, _ = eng.pack_lambda()
    sol.sync()
    assert np.array_equal(ll3.cpu().numpy(), want_l)
    sol.close()


@pytest.mark.parametrize("exchange_lambda", [True, False], ids=["lambda-exchanged", "lambda-local"])
@pytest.mark.parametrize("nranks", [2, 3])
def test_slabs_on_one_gpu_match_one_context(nranks, exchange_lambda):
    p_or, pos, vel, w, phase, ros = _scene()
    p = _params(p_or)
    steps = 4
    whole = psb.Solver(p, max_particles=pos.shape[0] + 16)
    whole.append(pos, vel, w, ros, phase)
    for _ in range(steps):
        whole.step(DT)
    cuts = slab.quantile_cuts(pos[:, 0], nranks)
    engines = [_ctx_engine(p, part) for part in _split(cuts, pos, vel, w, phase, ros)]
    cl = slab.LocalCluster(engines, cuts, exchange_lambda=exchange_lambda)
    for _ in range(steps):
        cl.step(DT)
    got_pos = np.concatenate([e.sol.download_owned(psb.ARR_POS) for e in engines])
    got_vel = np.concatenate([e.sol.download_owned(psb.ARR_VEL) for e in engines])
    assert got_pos.shape[0] == pos.shape[0]
    assert sum(d.stats["migrated_out"] for d in cl.doms) > 0 and all(d.stats["ghosts"] > 0 for d in cl.doms)
    # the same wall-jitter stream on every rank (each context seeds XORWOW with 1234 like the reference)
    assert np.array_equal(engines[0].sol.download(psb.ARR_RANDS), whole.download(psb.ARR_RANDS))
    _match(whole.download(psb.ARR_POS), whole.download(psb.ARR_VEL), got_pos, got_vel, tol=5e-5)
    for e in engines:
        e.sol.close()
    whole.close()


def test_slab_context_rejects_constraints():
    ps = psb.ParticleSystem.scene("2")  # cloth: distance + point constraints
    sol = ps.solver
    buf = torch.empty((1024, 32), dtype=torch.uint8, device="cuda")
    with pytest.raises(psb.PsError) as e:
        sol.slab_pack_halo(0.0, 1.0, 0.5, buf.data_ptr(), buf.data_ptr(), 1024)
    assert e.value.code == psb.PS_ERR_STATE
    ps.close()


def test_record_buffer_overflow_is_reported():
    p_or, pos, vel, w, phase, ros = _scene()
    eng = _ctx_engine(_params(p_or), (pos, vel, w, phase, ros))
    buf = torch.empty((8, 32), dtype=torch.uint8, device="cuda")
    with pytest.raises(psb.PsError) as e:
        eng.sol.slab_pack_halo(8.0, 15.0, 4.5, buf.data_ptr(), buf.data_ptr(), 8)
    assert e.value.code == psb.PS_ERR_CAPACITY
    eng.sol.close()


def test_x_histogram_kernel_matches_numpy():
    p_or, pos, vel, w, phase, ros = _scene()
    eng = _ctx_engine(_params(p_or), (pos, vel, w, phase, ros))
    ref = OracleEngine(p_or, pos, vel, w, phase, ros)
    for lo, hi, bins in ((0.0, 40.0, 512), (5.0, 12.0, 64), (0.0, 40.0, 65536)):
        got, want = eng.x_histogram(lo, hi, bins), ref.x_histogram(lo, hi, bins)
        assert got.sum() == pos.shape[0] == want.sum()
        assert np.abs(np.cumsum(got) - np.cumsum(want)).max() <= 2   # float32 vs float64 binning: a particle on a bin edge may move by one bin
    with pytest.raises(psb.PsError):
        eng.x_histogram(1.0, 1.0, 16)
    eng.sol.close()


def test_recut_on_the_gpu_keeps_the_result_and_balances():
    p_or, pos, vel, w, phase, ros = _scene(nx=36)
    p = _params(p_or)
    steps = 6
    whole = psb.Solver(p, max_particles=pos.shape[0] + 16)
    whole.append(pos, vel, w, ros, phase)
    for _ in range(steps):
        whole.step(DT)
    x = pos[:, 0]
    cuts = [-np.inf, float(np.quantile(x, 0.66)), float(np.quantile(x, 0.83)), np.inf]   # deliberately unbalanced
    engines = [_ctx_engine(p, part) for part in _split(cuts, pos, vel, w, phase, ros)]
    n0 = [e.n_owned for e in engines]
    cl = slab.LocalCluster(engines, cuts, recut_every=1, recut_range=(0.0, 40.0), recut_bins=512)
    for _ in range(steps):
        cl.step(DT)
    n1 = [e.n_owned for e in engines]
    got_pos = np.concatenate([e.sol.download_owned(psb.ARR_POS) for e in engines])
    got_vel = np.concatenate([e.sol.download_owned(psb.ARR_VEL) for e in engines])
    # 6 steps (the other slab tests run 4 at 5e-5): ghosts enter the neighbour sums in another order, the differences compound
    _match(whole.download(psb.ARR_POS), whole.download(psb.ARR_VEL), got_pos, got_vel, tol=1e-4)
    assert cl.doms[0].stats["recuts"] == steps - 1
    assert max(n1) / (sum(n1) / 3) < max(n0) / (sum(n0) / 3) - 0.3, (n0, n1)
    for e in engines:
        e.sol.close()
    whole.close()
