#!/usr/bin/env python
"""Distils an `ncu --set full` capture of step_kernel into profiles/step_kernel_profile.json (what bench.py quotes).

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > profiles/r2_xxx_ncu_raw.csv
    python tools/ncu_to_profile.py profiles/r2_xxx_ncu_raw.csv --task 1 --envs 65536 --csrc-hash <so100_build_id of the profiled library>

Per launch (averaged over the captured launches of the kernel): DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum),
executed FP32 work (FADD + FMUL + 2 FFMA thread instructions, predicated-on) per env step, FMA-pipe and issue-slot
utilisation over elapsed cycles, duration.  Entries are keyed by (task, envs, csrc_hash); bench.py marks a figure stale
when the loaded library's so100_build_id() differs.
"""
import argparse
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0,
        "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--task", type=int, required=True)
    ap.add_argument("--envs", type=int, required=True)
    ap.add_argument("--csrc-hash", required=True)
    ap.add_argument("--kernel", default="step_kernel")
    ap.add_argument("--note", default="")
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if args.kernel in r[rows[0].index("Kernel Name")]]
    if not data:
        raise SystemExit(f"no launches of {args.kernel} in {args.csv}")
    col = {h: i for i, h in enumerate(hdr)}

    def val(name):
        i = col[name]
        scale = UNIT.get(units[i], 1.0)
        return sum(float(r[i].replace(",", "")) for r in data) / len(data) * scale

    cycles = val("smsp__cycles_elapsed.avg") if "smsp__cycles_elapsed.avg" in col else val("sm__cycles_elapsed.avg")
    per_cyc = lambda op: val(f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed")  # noqa: E731
    fadd, fmul, ffma = per_cyc("fadd"), per_cyc("fmul"), per_cyc("ffma")
    entry = {
        "csrc_hash": args.csrc_hash, "task": args.task, "envs": args.envs, "launches_averaged": len(data),
        "kernel": data[0][col["Kernel Name"]],
        "duration_us": val("gpu__time_duration.sum") * 1e6,
        "traffic_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
        "executed_flop_per_env_step": (fadd + fmul + 2.0 * ffma) * cycles / args.envs,
        "executed_fp32_inst_per_env_step": {"fadd": fadd * cycles / args.envs, "fmul": fmul * cycles / args.envs, "ffma": ffma * cycles / args.envs},
        "pipe_fma_pct": val("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        "issue_active_pct": val("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "registers_per_thread": val("launch__registers_per_thread") if "launch__registers_per_thread" in col else None,
        "waves_per_sm": val("launch__waves_per_multiprocessor") if "launch__waves_per_multiprocessor" in col else None,
        "source": os.path.relpath(os.path.abspath(args.csv), ROOT), "note": args.note,
    }
    path = os.path.join(ROOT, "profiles", "step_kernel_profile.json")
    try:
        doc = json.load(open(path))
    except Exception:  # noqa: BLE001
        doc = {"entries": []}
    doc["entries"] = [e for e in doc["entries"] if (e["csrc_hash"], e["task"], e["envs"]) != (args.csrc_hash, args.task, args.envs)] + [entry]
    json.dump(doc, open(path, "w"), indent=1)
    print(json.dumps(entry, indent=1))


if __name__ == "__main__":
    main()
