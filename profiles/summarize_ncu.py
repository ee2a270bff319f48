#!/usr/bin/env python
"""Turn the scratch ncu outputs under gpurun_out/ into the tracked summaries under profiles/.

  python profiles/summarize_ncu.py <tag> <launches.csv> <full.ncu-rep> <fish per profiled launch>

Writes profiles/<tag>_launches.csv (the launch list as ncu wrote it), profiles/<tag>_launch_shares.md (per-kernel
share of the step), profiles/<tag>_passage_metrics.json (what bench.py reads: DRAM traffic per launch, executed
thread instructions per fish, pipe utilisations, stall reasons) and profiles/<tag>_passage_sass_hot.txt (SASS with
executed counts per tile of 32 fish).  Needs `ncu` on PATH (reads the .ncu-rep; no GPU).
"""
import csv
import json
import os
import re
import shutil
import subprocess
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))


def launch_shares(path):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[h + 1:]:
        if len(r) < 5:
            continue
        name = re.sub(r"\(.*", "", r[4])[:90]
        agg[name][0] += 1
        agg[name][1] += float(r[-1].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    lines = ["| kernel | launches | total ms (ncu, serialised, cold) | share |", "|---|---|---|---|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {v[1] / tot:.4f} |")
    return "\n".join(lines), {k: dict(launches=v[0], ms=v[1] / 1e6, share=v[1] / tot) for k, v in agg.items()}


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, vals = rows[0], rows[2]
    get = lambda k: float(vals[hdr.index(k)].replace(",", "")) if k in hdr and vals[hdr.index(k)] else None
    unit = lambda k: rows[1][hdr.index(k)] if k in hdr else ""
    m = {}
    for k in ("gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
              "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
        v = get(k)
        if v is not None:
            m[k] = {"value": v, "unit": unit(k)}
    stalls = {}
    for i, hname in enumerate(hdr):
        mm = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", hname)
        if mm and vals[i]:
            stalls[mm.group(1)] = float(vals[i])
    m["stall_reasons_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1]))
    return m


def sass_hot(rep, fish, path):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else ""
    hdr = rows[1]
    iS, iE, iA, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Address"), hdr.index("# Samples")
    tiles = fish / 32.0
    ops, tot = defaultdict(float), 0.0
    with open(path, "w") as fh:
        fh.write(f"# {kernel}\n# addr  executed-per-tile(32 fish)  stall-samples  SASS\n")
        for r in rows[2:]:
            if len(r) <= iE:
                continue
            try:
                e, s = float(r[iE]), float(r[iN])
            except ValueError:
                continue
            mm = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[iS])
            ops[(mm.group(2) if mm else "?").split(".")[0]] += e
            tot += e
            if e / tiles >= 0.5:
                fh.write(f"{r[iA][-5:]} {e / tiles:7.1f} {s:8.0f}  {r[iS][:110]}\n")
    return tot / tiles, {k: v / tiles for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:24]}


def main():
    tag, launches, rep, fish = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    shutil.copy(launches, os.path.join(HERE, f"{tag}_launches.csv"))
    table, shares = launch_shares(launches)
    m = raw_metrics(rep)
    per_tile, ops = sass_hot(rep, fish, os.path.join(HERE, f"{tag}_passage_sass_hot.txt"))
    dram = m.get("dram__bytes_read.sum", {}).get("value", 0.0)
    if m.get("dram__bytes_read.sum", {}).get("unit", "").lower().startswith("k"):
        dram *= 1e3
    elif m.get("dram__bytes_read.sum", {}).get("unit", "").lower().startswith("m"):
        dram *= 1e6
    dw = m.get("dram__bytes_write.sum", {})
    dram_w = dw.get("value", 0.0) * ({"k": 1e3, "m": 1e6}.get(dw.get("unit", " ")[:1].lower(), 1.0))
    thread_inst = m.get("smsp__thread_inst_executed.sum", {}).get("value")
    summary = {"tag": tag, "fish_in_profiled_launch": fish, "dram_bytes_per_launch": dram + dram_w,
               "warp_inst_per_tile": per_tile, "thread_inst_per_fish": (thread_inst / fish) if thread_inst else per_tile,
               "opcode_mix_per_tile": ops, "launch_shares": shares, "metrics": m,
               "how": "ncu --set full --clock-control none --import-source on (one launch of the passage kernel); "
                      "launch list: ncu --metrics gpu__time_duration.sum --clock-control none of the same bench command"}
    with open(os.path.join(HERE, f"{tag}_passage_metrics.json"), "w") as fh:
        json.dump(summary, fh, indent=1)
    with open(os.path.join(HERE, f"{tag}_launch_shares.md"), "w") as fh:
        fh.write(f"# {tag}: per-kernel share of `python bench.py` (ncu launch list; cold-cache, serialised)\n\n{table}\n")
    print(json.dumps({k: summary[k] for k in ("dram_bytes_per_launch", "warp_inst_per_tile", "thread_inst_per_fish")}))


if __name__ == "__main__":
    main()
