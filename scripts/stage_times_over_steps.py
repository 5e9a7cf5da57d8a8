"""per-stage device times (ps_step_profiled) of the C3 bench scene as the simulation advances: python scripts/stage_times_over_steps.py [side]"""
import sys
sys.path.insert(0, ".")
import numpy as np
import particlesolver_b200 as psb
side = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ps = psb.ParticleSystem.scene("c3", grid=256, max_particles=side ** 3 + 1024, side=side)
sol = ps.solver
print("n", sol.n)
done = 0
for target in (3, 10, 20, 30, 40, 60, 90, 120):
    while done < target:
        sol.step(1 / 60); done += 1
    st, ln = sol.step_profiled(1 / 60); done += 1
    nn = sol.download(psb.ARR_NUM_NEIGHBORS)
    y = sol.download(psb.ARR_POS)[:, 1]
    rows = sol.download(psb.ARR_NEIGHBOR_ROWS)
    ovf = int((rows == 0xFFFFFFFF).sum()); ok = rows[rows != 0xFFFFFFFF]
    print(f"step {done:4d}: lambda {st['lambda'] / 5:.3f} delta_p {st['delta_p'] / 5:.3f} sort {st['sort'] / 5:.3f} total {sum(st.values()):.2f} ms | neighbours mean {nn.mean():.1f} max {nn.max()} | y min {y.min():.2f} | rows mean {ok.mean():.0f} p99 {np.percentile(ok, 99):.0f} max {ok.max()} overflowed warps {ovf} | padding {(ok.sum() * 32.0) / max(nn.sum(), 1):.2f}x")
