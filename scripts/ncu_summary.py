#!/usr/bin/env python3
"""Turns `ncu -i X.ncu-rep --page raw --csv` into the small JSON kept under profiles/.

usage: ncu_summary.py raw.csv out.json "what was captured" [metric-prefix ...]
"""
import csv, json, sys

KEEP = ("dram__", "gpu__time", "sm__cycles", "sm__throughput", "sm__inst", "sm__warps", "smsp__inst",
        "smsp__issue", "smsp__average_warp", "smsp__warp_issue_stalled", "smsp__thread_inst",
        "sm__pipe", "sm__inst_executed_pipe", "l1tex__t_bytes", "l1tex__data_bank", "lts__t_bytes", "lts__t_sector",
        "launch__", "smsp__cycles_active", "sm__maximum_warps", "gpc__cycles_elapsed")


def main():
    raw, out, what = sys.argv[1:4]
    prefixes = tuple(sys.argv[4:]) or KEEP
    rows = list(csv.reader(open(raw)))
    names, units, vals = rows[0], rows[1], rows[2]
    metrics = {}
    for n, u, v in zip(names, units, vals):
        if n.startswith(prefixes):
            metrics[n] = (v + " " + u).strip()
    json.dump({"what": what, "metrics": metrics}, open(out, "w"), indent=1)
    print(len(metrics), "metrics ->", out)


if __name__ == "__main__":
    main()
