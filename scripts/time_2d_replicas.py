import os, sys, time
import particlesolver_b200 as psb
for key, ticks in (("6x2", 60), ("6x4", 12), ("6x4", 60)):
    sim = psb.Simulation2D.scene(key)
    t0 = time.perf_counter()
    marks = []
    for t in range(ticks):
        sim.tick(.01)
        marks.append(time.perf_counter())
    n = sim.getNumParticles()
    steady = (marks[-1] - marks[len(marks)//2]) / (len(marks) - 1 - len(marks)//2)
    print(os.environ.get("PS_LIBRARY", "default")[-24:], key, n, ticks, "mean ms %.3f" % (1e3*(marks[-1]-t0)/ticks), "second-half ms %.3f" % (1e3*steady), "first tick ms %.1f" % (1e3*(marks[0]-t0)), "launches", sim.launches_per_tick)
    sim.close()
