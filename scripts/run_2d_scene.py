"""tiny driver for profiling the 2-D path: python scripts/run_2d_scene.py KEY TICKS"""
import sys, time
sys.path.insert(0, ".")
import particlesolver_b200 as psb
key, ticks = sys.argv[1], int(sys.argv[2])
sim = psb.Simulation2D.scene(key)
for _ in range(5):
    sim.tick(.01)
t0 = time.perf_counter()
for _ in range(ticks):
    sim.tick(.01)
dt = time.perf_counter() - t0
print(f"scene {key}: n={sim.getNumParticles()} {1e3 * dt / ticks:.3f} ms/tick launches/tick={sim.launches_per_tick} levels={sim.num_levels} contacts={sim.num_contact_constraints}")
