"""ps_step_streamed (include/psolver.h; csrc/ps_stream_io.cu): a host that owns positions and velocities gets the same steps as a
resident run — the double-buffered transfers change when bytes move, not what is computed.  Reference behaviour: the host reads
the state back after update() (gpu/src/particlesystem.cpp:122-142, 248-262)."""
import numpy as np
import pytest
import torch

import particlesolver_b200 as psb

pytestmark = pytest.mark.gpu
DT = 1.0 / 60.0


def _pinned(n):
    return torch.empty((n, 4), dtype=torch.float32).pin_memory()


@pytest.mark.parametrize("scene", ["7", "8"])
def test_streamed_steps_equal_resident_steps(scene):
    steps = 6
    ref = psb.ParticleSystem.scene(scene)
    states = []
    for _ in range(steps):
        ref.update(DT)
        states.append((ref.solver.download(psb.ARR_POS).copy(), ref.solver.download(psb.ARR_VEL).copy()))
    ref.close()

    ps = psb.ParticleSystem.scene(scene)
    sol = ps.solver
    n = sol.n
    hin = [(_pinned(n), _pinned(n)) for _ in range(2)]
    hout = [(_pinned(n), _pinned(n)) for _ in range(2)]
    hin[0][0].numpy()[:] = sol.download(psb.ARR_POS)
    hin[0][1].numpy()[:] = sol.download(psb.ARR_VEL)
    for k in range(steps):
        (ip, iv), (op, ov) = hin[k & 1], hout[k & 1]
        sol.step_streamed(DT, ip.data_ptr(), iv.data_ptr(), op.data_ptr(), ov.data_ptr())
        sol.io_wait(0)
        assert np.array_equal(op.numpy(), states[k][0]) and np.array_equal(ov.numpy(), states[k][1]), f"step {k}"
        # the host hands the delivered state back as the next step's input (a host-driven loop)
        hin[(k + 1) & 1][0].numpy()[:] = op.numpy()
        hin[(k + 1) & 1][1].numpy()[:] = ov.numpy()
    ps.close()


def test_streamed_pipeline_delivers_every_frame_in_order():
    """inputs submitted ahead of the results (the pipelined use): outputs of call k land in the k-th buffer pair, one call late at most"""
    ps = psb.ParticleSystem.scene("7")
    sol = ps.solver
    n = sol.n
    ref = psb.ParticleSystem.scene("7")
    frames = []
    for _ in range(5):
        ref.update(DT)
        frames.append(ref.solver.download(psb.ARR_POS).copy())
    ref.close()
    outs = [(_pinned(n), _pinned(n)) for _ in range(2)]
    got = []
    for k in range(5):
        op, ov = outs[k & 1]
        sol.step_streamed(DT, None, None, op.data_ptr(), ov.data_ptr())   # state stays resident, results are delivered
        if k >= 1:
            sol.io_wait(1)
            got.append(outs[(k - 1) & 1][0].numpy().copy())
    sol.io_wait(0)
    got.append(outs[4 & 1][0].numpy().copy())
    for k in range(5):
        assert np.array_equal(got[k], frames[k]), f"frame {k}"
    ps.close()


def test_begin_prefetch_end_around_explicit_steps():
    """ps_io_begin / ps_io_prefetch / ps_io_end around a step the caller issues itself (what a slab-decomposed run does): step k's inputs
    are the recorded state k of a resident run, the upload of step k+1's inputs is started before step k is issued, and every delivered
    frame equals the resident run's next state"""
    steps = 5
    ref = psb.ParticleSystem.scene("7")
    n = ref.solver.n
    states = [(ref.solver.download(psb.ARR_POS).copy(), ref.solver.download(psb.ARR_VEL).copy())]
    for _ in range(steps):
        ref.update(DT)
        states.append((ref.solver.download(psb.ARR_POS).copy(), ref.solver.download(psb.ARR_VEL).copy()))
    ref.close()
    ps = psb.ParticleSystem.scene("7")
    sol = ps.solver
    hin = [(_pinned(n), _pinned(n)) for _ in range(steps)]
    for k in range(steps):
        hin[k][0].numpy()[:] = states[k][0]
        hin[k][1].numpy()[:] = states[k][1]
    hout = [(_pinned(n), _pinned(n)) for _ in range(2)]
    # perturb the resident state: every step must really start from the host's frame
    sol.upload(psb.ARR_POS, states[0][0] + np.float32(0.01))
    for k in range(steps):
        sol.io_begin(hin[k][0].data_ptr(), hin[k][1].data_ptr())
        if k + 1 < steps:
            sol.io_prefetch(hin[k + 1][0].data_ptr(), hin[k + 1][1].data_ptr())
        sol.step(DT)
        op, ov = hout[k & 1]
        sol.io_end(op.data_ptr(), ov.data_ptr())
        sol.io_wait(0)
        assert np.array_equal(op.numpy(), states[k + 1][0]) and np.array_equal(ov.numpy(), states[k + 1][1]), f"step {k}"
    ps.close()
