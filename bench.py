#!/usr/bin/env python
"""bench.py — particle-steps/s of the unified particle solver step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (config.workload): "c3" = the reference's GPU scene 7 (fluid blob with artificial-pressure "surface
tension", gpu/src/particleapp.cpp:172-176) scaled to 100^3 = 1,000,000 PBF particles on a 256^3 grid, 5 solver
iterations, dt = 1/60 — BASELINE.json configs[2], SURVEY.md §8 C3.  One "step" = one ParticleSystem::update.

Own arm (default):
  value     whole-job particle-steps/s, state resident in HBM, K steps timed with CUDA events on the solver's
            stream (CUDA-graph replay), max over ranks.
  e2e       same metric through the public host API with HOST buffers: every step uploads positions+velocities
            from pinned host memory, steps, and downloads positions+velocities.
  roofline  dominant kernel (largest share of the step) against the measured HBM peak, from per-stage CUDA
            events of an instrumented (eager) pass over the same state.
  cpu_baseline  the oracle port (oracle/gpu_step_oracle.c, OpenMP) on a bounded sample of the workload.
Reference arm (--impl reference): the reference algorithm on the host cores — the oracle port with all threads
on a bounded sample (the reference's CPU app is 2-D only; its GPU sources need a GPU), plus, as extra fields, the
reference's own unmodified CPU solver (oracle/_ref/ref_cpu, scene 6) when that binary is present.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DT = 1.0 / 60.0
SIDE = 100          # c3: SIDE^3 fluid particles
GRID = 256
ITERS = 5

# algorithmic (compulsory) HBM bytes per particle per launch of each stage — SURVEY.md §8(d) / BASELINE.md §5.
# sort: 4 + 16 P with P = radix passes (3 for 2^24 cells); cell_table: 4 B per CELL (one write of the dense table) + 4 B per key.
STAGE_BYTES = {"predict": 64, "hash": 24, "sort": 52, "reorder": 56, "contacts": 60, "lambda": 36, "delta_p": 64, "world": 52,
               "velocity": 48}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def c3_sample(side, seed=1234):
    """CPU-side sample of the c3 workload: side^3 lattice, spacing 2.5 r, jitter +-0.01 r, rho0 1.5, mass 1."""
    rng = np.random.default_rng(seed)
    ext = int(np.ceil(side * 0.625))
    l = -(ext // 2)
    z, y, x = np.meshgrid(np.arange(side), np.arange(side), np.arange(side), indexing="ij")
    pos = np.ones((side ** 3, 4), np.float32)
    pos[:, 0] = l + x.ravel() * np.float32(0.625)
    pos[:, 1] = 6 + y.ravel() * np.float32(0.625)
    pos[:, 2] = l + z.ravel() * np.float32(0.625)
    pos[:, :3] += rng.uniform(-0.0025, 0.0025, size=(side ** 3, 3)).astype(np.float32)
    return pos


def time_oracle_port(side, steps):
    """particle-steps/s of the CPU port on a side^3 sample of the workload, all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_py as orc
    pos = c3_sample(side)
    n = pos.shape[0]
    p = orc.make_params(grid=(GRID, GRID, GRID))
    o = orc.OracleSystem(p, pos, np.zeros((n, 4), np.float32), np.ones(n, np.float32), np.zeros(n, np.int32), np.full(n, 1.5, np.float32),
                         iterations=ITERS)
    rands = np.full((ITERS, 6), 0.5, np.float32)
    o.step(DT, rands)  # warm-up (page faults, thread pool)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(DT, rands)
    dt = time.perf_counter() - t0
    return n * steps / dt, n, dt / steps


def time_reference_cpu_solver():
    """The reference's own unmodified 2-D CPU solver on its scene 6 (oracle/_ref/ref_cpu), if the binary is here."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_cpu")
    if not os.path.exists(exe):
        return None
    try:
        out = subprocess.run([exe, "--scene", "6", "--ticks", "200", "--json"], capture_output=True, text=True, timeout=120).stdout
        for line in out.splitlines():
            if line.startswith("{"):
                return json.loads(line)
    except Exception as e:
        return {"error": str(e)}
    return None


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    side = 58  # 195,112 particles: a few seconds per step on a multi-core host
    steps = max(1, min(args.steps, 3))
    for _ in range(max(0, min(args.warmup, 1))):
        pass  # time_oracle_port warms up once itself
    v, n, sec = time_oracle_port(side, steps)
    line = {"metric": "particle-steps/sec", "value": v, "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": "c3: reference GPU scene 7 scaled, PBF fluid, 5 solver iterations, dt=1/60, 256^3 grid",
                       "sample": f"{side}^3 = {n} particles of the 1,000,000 (same lattice, density, parameters)", "steps_timed": steps},
            "cpu_baseline": {"value": v, "unit": "particle-steps/s", "cores": cores, "kind": "port",
                             "sample": f"oracle/gpu_step_oracle.c (C restatement of the reference GPU step, OpenMP) on {n} particles x {steps} steps"},
            "e2e": {"value": v, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    ref_cpu = time_reference_cpu_solver()
    if ref_cpu:
        line["reference_cpu_solver"] = ref_cpu
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import particlesolver_b200 as psb
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the solver has no CPU path")
    torch.cuda.set_device(local_rank)
    os.environ["PS_DEVICE"] = str(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_target = SIDE ** 3
    ps = psb.ParticleSystem.scene("c3", grid=GRID, max_particles=n_target + 4096, iterations=ITERS, side=SIDE, seed=1 + rank)
    sol = ps.solver
    n = ps.getNumParticles()
    assert n == n_target, n

    # ---------------- resident: K graph-replayed steps, CUDA events on the solver stream ----------------
    for _ in range(max(args.warmup, 3)):
        ps.update(DT)
    sol.sync()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    sol.timer_start()
    for _ in range(args.steps):
        ps.update(DT)
    ms = sol.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    launches = sol.launches_per_step * args.steps
    value = n * world * args.steps / (ms * 1e-3)

    if args.quick:
        if rank == 0:
            stages = {}
            if not os.environ.get("PS_NO_GRAPH"):  # under ncu (PS_NO_GRAPH=1) keep the launch list to the plain steps
                for _ in range(3):
                    st, _ln = sol.step_profiled(DT)
                    for k, v in st.items():
                        stages[k] = round(stages.get(k, 0.0) + v / 3 / (1 if k in ("predict", "velocity") else ITERS), 4)
            print(json.dumps({"metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
                              "ms_per_step": ms / args.steps, "gpu_launches": launches, "quick": True, "lib": os.environ.get("PS_LIBRARY", "default"),
                              "stage_ms_per_launch": stages, "clocks": clocks}), flush=True)
        ps.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---------------- end to end: host buffers in, host buffers out, every step ----------------
    hpos = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    hvel = torch.empty((n, 4), dtype=torch.float32).pin_memory()
    hpos.numpy()[:] = sol.download(psb.ARR_POS)
    hvel.numpy()[:] = sol.download(psb.ARR_VEL)
    e2e_steps = args.steps

    def e2e_step():
        sol.upload_async(psb.ARR_POS, hpos.data_ptr(), 4 * n)
        sol.upload_async(psb.ARR_VEL, hvel.data_ptr(), 4 * n)
        ps.update(DT)
        sol.download_async(psb.ARR_POS, hpos.data_ptr(), 4 * n)
        sol.download_async(psb.ARR_VEL, hvel.data_ptr(), 4 * n)
        sol.sync()  # the host owns the result before it submits the next step

    for _ in range(3):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    sol.timer_start()
    for _ in range(e2e_steps):
        e2e_step()
    e2e_ms_dev = sol.timer_stop()
    e2e_wall = (time.perf_counter() - t0) * 1e3
    barrier()
    e2e_ms = max_over_ranks(max(e2e_ms_dev, e2e_wall))
    e2e_value = n * world * e2e_steps / (e2e_ms * 1e-3)

    # ---------------- per-stage device times (instrumented eager pass over the same state) ----------------
    prof_steps = min(args.steps, 10)
    acc, launches_by_stage = {}, {}
    for _ in range(prof_steps):
        st, ln = sol.step_profiled(DT)
        for k, v in st.items():
            acc[k] = acc.get(k, 0.0) + v
        launches_by_stage = ln
    peak, peak_kind = measured_peaks()
    cells = GRID ** 3
    kernels = {}
    step_ms_prof = sum(acc.values()) / prof_steps
    for k, total in acc.items():
        per_step = total / prof_steps
        if per_step <= 0:
            continue
        calls = 1 if k in ("predict", "velocity") else ITERS
        if k == "cell_table":
            b = 4.0 * cells + 4.0 * n  # write the dense table once, read the sorted keys once
        elif k in STAGE_BYTES:
            b = STAGE_BYTES[k] * n
        else:
            continue
        gbs = b * calls / (per_step * 1e-3) / 1e9
        kernels[k] = {"ms_per_step": round(per_step, 4), "share": round(per_step / step_ms_prof, 4), "alg_bytes_per_launch": b,
                      "achieved_gbs": round(gbs, 1), "frac_hbm": round(gbs / peak, 4)}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")  # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(dom)
        except Exception:
            traffic = None
    roofline = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s",
                "frac": kernels[dom]["frac_hbm"], "traffic": traffic, "peak_kind": peak_kind,
                "note": "neighbour kernels (contacts/lambda/delta_p) are FP32-issue/L1 bound, ~370 candidate tests per particle against "
                        "36-64 compulsory bytes (SURVEY 8d); streaming kernels are the HBM-bound ones, see 'kernels'"}

    if rank == 0:
        cpu_v, cpu_n, cpu_sec = time_oracle_port(46, 1)
        line = {"metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "c3: reference GPU scene 7 scaled to 100^3 = 1,000,000 PBF particles (addFluid((-31,6,-31),(32,69,32),1,1.5)), "
                                       "256^3 grid, 5 solver iterations, dt=1/60", "particles_per_gpu": n,
                           "multi_gpu": "independent replicas" if world > 1 else "single",
                           "l2": "working set 1M x ~150 B + 2 x 64 MB cell tables > 126 MB L2; no flush"},
                "particle_iterations_per_s": value * ITERS,
                "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 32 * n,
                        "ms_per_step": e2e_ms / e2e_steps},
                "gpu_launches": launches,
                "roofline": roofline, "kernels": kernels,
                "cpu_baseline": {"value": cpu_v, "unit": "particle-steps/s", "cores": os.cpu_count() or 1, "kind": "port",
                                 "sample": f"oracle port (OpenMP) on a 46^3 = {cpu_n} particle block of the same lattice, 1 step"},
                "clocks": clocks}
        print(json.dumps(line), flush=True)
    ps.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--quick", action="store_true", help="resident timing only (for runs under ncu): no e2e, no per-stage pass, no CPU baseline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
