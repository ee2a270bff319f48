#!/usr/bin/env python
"""Round-2 profile summaries: gpurun_out/r02_<w>_step_metrics.csv (ncu, one timed step of `bench.py --workload <w> --steps 1`,
metrics per launch of count_kernel / passage_spec) + gpurun_out/r02_<w>_step.json (the bench line of that same run: the
fish simulated in the step) -> profiles/r02_<w>_passage_metrics.json (what bench.py reads for roofline.issue_util and
roofline.traffic) and profiles/r02_<w>_launches.md (per-launch table).

  python profiles/summarize_r02.py c2 c3 c5
"""
import csv
import json
import os
import sys
from collections import defaultdict

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def read_metrics(path, warm_launches):
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = defaultdict(dict)
    order = []
    for r in rows:
        key = int(r["ID"])
        if key not in per:
            order.append(key)
        per[key]["kernel"] = r["Kernel Name"].split("(")[0].replace("void ", "")
        try:
            per[key][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            pass
    launches = [per[k] for k in order]
    return launches


def main(names):
    for w in names:
        mpath = os.path.join(ROOT, "gpurun_out", f"r02_{w}_step_metrics.csv")
        jpath = os.path.join(ROOT, "gpurun_out", f"r02_{w}_step.json")
        if not (os.path.isfile(mpath) and os.path.isfile(jpath)):
            print("missing", w)
            continue
        line = json.loads([ln for ln in open(jpath) if ln.startswith("{")][-1])
        fish = line["config"]["fish_passages_per_step"]
        per_step = int(round(line["roofline"]["kernel_launches_per_step"]))
        launches = read_metrics(mpath, 0)
        passage = [x for x in launches if "passage" in x["kernel"]]
        count = [x for x in launches if "count_kernel" in x["kernel"]]
        # the LAST step of the run is the timed one (warm-up steps come first; the e2e calls after it launch through another plan
        # with the same kernels: take the last `per_step` launches before them = launches [3 * per_step, 4 * per_step))
        p_step = passage[3 * per_step:4 * per_step]
        c_step = count[3 * per_step:4 * per_step]
        tot = lambda xs, k: sum(x.get(k, 0.0) for x in xs)
        inst = tot(p_step, "smsp__inst_executed.sum")
        t_ns = tot(p_step, "gpu__time_duration.sum")
        wavg = lambda xs, k: sum(x.get(k, 0.0) * x.get("gpu__time_duration.sum", 0.0) for x in xs) / max(tot(xs, "gpu__time_duration.sum"), 1.0)
        out = {
            "workload": line["config"]["workload"], "source": f"ncu --metrics ... -k regex:passage_spec|count_kernel, one step of bench.py --workload {w} --steps 1",
            "fish_passages_per_step": fish, "passage_launches_per_step": len(p_step),
            "warp_inst_per_fish": inst / fish, "thread_inst_per_fish": 32.0 * inst / fish,
            "passage_ms_per_step_ncu": t_ns / 1e6, "count_ms_per_step_ncu": tot(c_step, "gpu__time_duration.sum") / 1e6,
            "count_share_of_kernels": tot(c_step, "gpu__time_duration.sum") / max(t_ns + tot(c_step, "gpu__time_duration.sum"), 1.0),
            "issue_active_pct": wavg(p_step, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "pipe_alu_pct": wavg(p_step, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
            "pipe_fma_pct": wavg(p_step, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
            "pipe_xu_pct": wavg(p_step, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
            "pipe_lsu_pct": wavg(p_step, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
            "pipe_fp64_pct": wavg(p_step, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": p_step[0].get("launch__registers_per_thread") if p_step else None,
            "dram_bytes_per_launch": tot(p_step, "dram__bytes_read.sum") + tot(p_step, "dram__bytes_write.sum")
                                     + tot(c_step, "dram__bytes_read.sum") + tot(c_step, "dram__bytes_write.sum"),
            "dram_bytes_passage": tot(p_step, "dram__bytes_read.sum") + tot(p_step, "dram__bytes_write.sum"),
            "dram_bytes_count": tot(c_step, "dram__bytes_read.sum") + tot(c_step, "dram__bytes_write.sum"),
            "algorithmic_bytes_per_step": line["roofline"]["hbm"]["algorithmic_bytes_per_step"],
            "note": "dram_bytes_per_launch = DRAM bytes of ALL count + passage launches of one step (the name bench.py reads)",
        }
        with open(os.path.join(HERE, f"r02_{w}_passage_metrics.json"), "w") as fh:
            json.dump(out, fh, indent=1)
        md = [f"# r02 {w}: launches of one timed step (ncu, serialised; `bench.py --workload {w} --steps 1`)", "",
              "| # | kernel | us | warp inst | issue active % | DRAM read MB | DRAM write MB |", "|---|---|---|---|---|---|---|"]
        for i, x in enumerate(sorted(c_step + p_step, key=lambda z: launches.index(z))):
            md.append(f"| {i} | `{x['kernel']}` | {x.get('gpu__time_duration.sum', 0) / 1e3:.1f} | {x.get('smsp__inst_executed.sum', 0):.4g} | "
                      f"{x.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):.1f} | {x.get('dram__bytes_read.sum', 0) / 1e6:.2f} | "
                      f"{x.get('dram__bytes_write.sum', 0) / 1e6:.2f} |")
        md += ["", f"fish-passages in the step: {fish:.6g}; warp instructions per fish (passage): {out['warp_inst_per_fish']:.3f} "
                   f"(= {out['thread_inst_per_fish']:.1f} lane-ops); count kernels' share of the kernel time: {out['count_share_of_kernels']:.3f}; "
                   f"DRAM bytes per step {out['dram_bytes_per_launch'] / 1e6:.1f} MB vs algorithmic {out['algorithmic_bytes_per_step'] / 1e6:.1f} MB"]
        with open(os.path.join(HERE, f"r02_{w}_launches.md"), "w") as fh:
            fh.write("\n".join(md) + "\n")
        print(w, json.dumps({k: out[k] for k in ("warp_inst_per_fish", "issue_active_pct", "count_share_of_kernels", "dram_bytes_per_launch", "algorithmic_bytes_per_step")}))


if __name__ == "__main__":
    main(sys.argv[1:] or ["c2", "c3", "c5"])
