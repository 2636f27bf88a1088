#!/usr/bin/env python3
"""profiles/ncu_counts.json from an ncu metrics pass over a few steps of `bench.py --device-only`.

  ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,\
smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum \
      --clock-control none -k regex:hope -c 400 --csv --log-file gpurun_out/counts.csv python bench.py --steps 3 --warmup 3 --device-only --no-cpu-baseline
  python profiles/tools/ncu_counts.py gpurun_out/counts.csv <commit> > profiles/ncu_counts.json

Per kernel: launches seen, mean duration, mean DRAM bytes per launch, float64 flops per launch (DADD + DMUL + 2 DFMA thread
instructions).  bench.py reads `fp64_flops_per_step_65536` (the step's kernels, one launch each) and `dram_bytes_per_launch`.
Only launches over the full 65 536 envs are counted (the reset step and small set-up launches are dropped by grid size).
"""
import collections
import csv
import json
import sys

STEP_KERNELS = ("k_advance", "k_observe", "k_rs_enumerate", "k_rs_walk", "k_rs_check", "k_rs_select")


def main():
    path, commit = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    h = rows[0]
    ki, mi, vi, ui, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    per_launch = collections.OrderedDict()
    for r in rows[1:]:
        key = (r[idi], r[ki].split("(")[0].replace("void ", "").replace("hope::", "").split("<")[0])
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        if unit in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        if r[mi] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)  # -> microseconds
        per_launch.setdefault(key, {})[r[mi]] = v
    agg = collections.defaultdict(list)
    for (lid, name), m in per_launch.items():
        agg[name].append(m)
    out = {"commit": commit, "source": path, "per_kernel": {}, "dram_bytes_per_launch": {}}
    flops_step = 0.0
    for name, ms in agg.items():
        # the steady-state launches: drop the first (the reset step) when there are several
        use = ms[1:] if len(ms) > 1 else ms
        mean = lambda k: sum(m.get(k, 0.0) for m in use) / len(use)
        flops = mean("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum") + mean("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum") + \
            2 * mean("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum")
        dram = mean("dram__bytes_read.sum") + mean("dram__bytes_write.sum")
        out["per_kernel"][name] = {"launches": len(ms), "us_per_launch_under_ncu": mean("gpu__time_duration.sum"), "dram_bytes_per_launch": dram,
                                   "fp64_flops_per_launch": flops, "warp_instructions_per_launch": mean("smsp__inst_executed.sum")}
        out["dram_bytes_per_launch"][name] = dram
        if name in STEP_KERNELS:
            flops_step += flops
    out["fp64_flops_per_step_65536"] = flops_step
    out["note"] = "fp64 flops = DADD + DMUL + 2 x DFMA executed thread instructions (predicated-on), per launch, mean over the steady-state launches; " \
                  "the step = one launch each of " + ", ".join(STEP_KERNELS)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
