"""torchrun worker of test_gpu_slab_nccl.py: every rank owns one slab of a moving fluid block on its own GPU, the
exchange goes over NCCL; rank 0 also runs the undecomposed system on its GPU and compares."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import particlesolver_b200 as psb  # noqa: E402
from particlesolver_b200 import slab  # noqa: E402
from test_slab_cpu import _match, _scene, _split  # noqa: E402

DT = 1.0 / 60.0


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    recut = int(sys.argv[2]) if len(sys.argv) > 2 else 0   # re-cut the slabs every that many steps (0: fixed planes)
    exchange = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True   # ghost lambdas from their owners (1) or computed locally (0)
    c_abi = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False      # the exchange behind the C ABI (ps_comm_*, NCCL C API) instead of slab.py
    p_or, pos, vel, w, phase, ros = _scene(nx=40)
    p = psb.default_params()
    p.grid_size[:] = tuple(p_or.grid); p.min_bounds[:] = tuple(p_or.min_b); p.max_bounds[:] = tuple(p_or.max_b)
    cuts = slab.quantile_cuts(pos[:, 0], world)
    if recut:  # start unbalanced: the first slab holds more than its share
        q = np.quantile(pos[:, 0], [min(0.95, (r + 0.6) / world) for r in range(1, world)])
        cuts = [-np.inf, *[float(v) for v in q], np.inf]
    mine = _split(cuts, pos, vel, w, phase, ros)[rank]
    sol = psb.Solver(p, max_particles=pos.shape[0], device=local)
    sol.append(mine[0], mine[1], mine[2], mine[4], mine[3])
    eng = slab.CtxEngine(sol, halo_capacity=pos.shape[0], migrant_capacity=pos.shape[0])
    dom = slab.SlabDomain(eng, rank, world, cuts, comm=slab.DistComm(eng), recut_every=recut, recut_range=(0.0, 40.0), recut_bins=1024,
                          exchange_lambda=exchange)
    n_start = sol.n_owned
    if c_abi:
        # rank 0's ncclGetUniqueId travels over the process group that launched us; from here on the step is ps_comm_step alone
        ident = [psb.Solver.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ident, src=0)
        sol.comm_init(ident[0], rank, world)
        sol.comm_set_slab(cuts[rank], cuts[rank + 1], drift=0.25, exchange_lambda=exchange, halo_capacity=pos.shape[0], migrant_capacity=pos.shape[0])
        if recut:
            sol.comm_set_recut(cuts, recut, 0.0, 40.0, 1024)
        for _ in range(steps):
            sol.comm_step(DT)
        st = sol.comm_stats()
        dom.stats["migrated_out"], dom.stats["ghosts"] = st["migrated_out"], st["ghosts"]
        dom.comm.bytes_sent = st["bytes_sent"]
        if recut:
            new_cuts, dom.stats["recuts"] = sol.comm_cuts(world)
            dom.cuts = [float(v) for v in new_cuts]
        total = sol.comm_allreduce_sum([sol.n_owned])
        assert int(total[0]) == pos.shape[0], (total, pos.shape)
    else:
        for _ in range(steps):
            dom.step(DT)
    sol.sync()
    # gather the owned particles on rank 0
    n = torch.tensor([sol.n_owned], dtype=torch.int64, device="cuda")
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    cap = max(ns)
    buf = torch.zeros((cap, 8), dtype=torch.float32, device="cuda")
    buf[:ns[rank], :4] = torch.from_numpy(sol.download_owned(psb.ARR_POS)).cuda()
    buf[:ns[rank], 4:] = torch.from_numpy(sol.download_owned(psb.ARR_VEL)).cuda()
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    stats = torch.tensor([dom.stats["migrated_out"], dom.stats["ghosts"], dom.comm.bytes_sent], dtype=torch.int64, device="cuda")
    dist.all_reduce(stats)
    if rank == 0:
        got = np.concatenate([b[:k].cpu().numpy() for b, k in zip(bufs, ns)])
        assert got.shape[0] == pos.shape[0], (got.shape, pos.shape)
        whole = psb.Solver(p, max_particles=pos.shape[0] + 16, device=local)
        whole.append(pos, vel, w, ros, phase)
        for _ in range(steps):
            whole.step(DT)
        _match(whole.download(psb.ARR_POS), whole.download(psb.ARR_VEL), got[:, :4], got[:, 4:], tol=5e-5 if not recut else 1e-4)
        if recut:
            assert dom.stats["recuts"] == (steps - 1) // recut and max(ns) / (sum(ns) / world) < 1.3, (ns, dom.cuts)
        assert int(stats[0]) > 0 and int(stats[1]) > 0 and int(stats[2]) > 0
        print(f"SLAB_NCCL_OK c_abi={int(c_abi)} recut={recut} exchange_lambda={exchange} ghosts={int(stats[1])} cuts={[round(c, 3) for c in dom.cuts[1:-1]]} world={world} particles={pos.shape[0]} owned={ns} migrated={int(stats[0])} bytes_sent={int(stats[2])}", flush=True)
    dist.barrier()
    sol.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
