#!/usr/bin/env python
"""Summarise ncu output into small text/JSON files for profiles/ (the .ncu-rep files themselves are scratch).

  python scripts/ncu_summary.py launches gpurun_out/launches.csv  > profiles/<round>_launches.txt
  python scripts/ncu_summary.py full gpurun_out/prof.ncu-rep      > profiles/<round>_ncu_full.txt
  python scripts/ncu_summary.py traffic gpurun_out/prof.ncu-rep   > profiles/ncu_traffic.json   (dram bytes per launch by stage)
"""
import collections
import csv
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STAGE_OF = {"k_find_lambdas": "lambda", "k_find_lambdas_fused": "lambda", "k_find_lambdas_staged": "lambda", "k_calc_hash_hist": "hash", "k_solve_fluids": "delta_p", "k_solve_fluids_list": "delta_p", "k_radix_pass": "sort", "k_radix_hist": "sort", "k_reorder": "reorder",
            "k_cell_begin": "cell_table", "k_cell_fill": "cell_table", "k_collide_world": "world", "k_collide": "contacts", "k_calc_hash": "hash",
            "k_predict": "predict", "k_velocity": "velocity", "k_lambda": "lambda", "k_delta": "delta_p"}


def short(name):
    n = name.split("(")[0]
    return n.replace("<unnamed>::", "").replace("void ", "").strip()


def to_float(v, unit, want):
    v = float(v.replace(",", ""))
    scale = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "s": 1.0, "second": 1.0, "nsecond": 1e-9,
             "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return v * scale.get(unit, 1.0)


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        a = agg.setdefault(short(r[ki]), [0, 0.0])
        a[0] += 1
        a[1] += to_float(r[vi], r[ui], "s") * 1e6
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none ; {len(rows) - 1} launches, {tot / 1e3:.3f} ms of kernel time")
    print(f"# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"{'kernel':48s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}")
    for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:48]:48s} {a[0]:8d} {a[1]:10.1f} {a[1] / a[0]:9.1f} {a[1] / tot:7.3f}")


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def full(rep):
    hdr, units, rows = raw_rows(rep)
    idx = [(k, hdr.index(k)) for k in KEYS if k in hdr]
    ni = hdr.index("Kernel Name")
    print(f"# ncu --set full --clock-control none --import-source on ; {rep} ; one block per captured launch")
    for r in rows:
        print(f"== {short(r[ni])}")
        for k, i in idx:
            print(f"   {k:66s} {r[i]:>16s} {units[i]}")


def traffic(rep):
    hdr, units, rows = raw_rows(rep)
    ni, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    acc = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows:
        k = short(r[ni]).split("<")[0]
        b = to_float(r[ri], units[ri], "B") + to_float(r[wi], units[wi], "B")
        acc[STAGE_OF.get(k, k)][k].append(b)
    # per stage call: sum over the stage's kernels of (mean bytes per launch) x (launches of that kernel per stage call); the
    # multiplicity is taken from the launch counts in the capture (e.g. k_radix_pass runs P times per k_radix_hist)
    out = {}
    for stage, ks in acc.items():
        fewest = min(len(v) for v in ks.values())
        per_kernel = {k: {"bytes_per_launch": sum(v) / len(v), "launches_per_stage_call": round(len(v) / fewest, 2)} for k, v in ks.items()}
        out[stage] = sum(d["bytes_per_launch"] * d["launches_per_stage_call"] for d in per_kernel.values())
        out[stage + "__kernels"] = per_kernel
    print(json.dumps(out, indent=1))


def inst(rep):
    """warp-instructions per stage call (smsp__inst_executed.sum), same aggregation as traffic(): profiles/ncu_inst.json"""
    hdr, units, rows = raw_rows(rep)
    ni, ii = hdr.index("Kernel Name"), hdr.index("smsp__inst_executed.sum")
    acc = collections.defaultdict(lambda: collections.defaultdict(list))
    for r in rows:
        k = short(r[ni]).split("<")[0]
        acc[STAGE_OF.get(k, k)][k].append(float(r[ii].replace(",", "")))
    out = {}
    for stage, ks in acc.items():
        fewest = min(len(v) for v in ks.values())
        out[stage] = sum(sum(v) / len(v) * round(len(v) / fewest, 2) for v in ks.values())
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic, "inst": inst}[sys.argv[1]](sys.argv[2])
