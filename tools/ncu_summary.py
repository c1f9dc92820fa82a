#!/usr/bin/env python
"""Summarise ncu artefacts from gpurun_out/ into profiles/ (tracked): launch-list shares and the key
raw metrics of a --set full capture.  Usage: tools/ncu_summary.py <tag> <launches.csv|-> <rep|-> <workload>"""
from __future__ import annotations

import collections
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__waves_per_multiprocessor", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def launch_shares(path: Path) -> list[dict]:
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    agg = collections.defaultdict(list)
    for r in rows:
        agg[r["Kernel Name"].split("(")[0]].append(float(r["Metric Value"]))
    tot = sum(sum(v) for v in agg.values())
    ours = sum(sum(v) for k, v in agg.items() if "sb::" in k)
    return [{"kernel": k, "launches": len(v), "mean_us": sum(v) / len(v) / 1e3, "share_of_all": sum(v) / tot,
             "share_of_our_kernels": (sum(v) / ours if "sb::" in k else None)}
            for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1]))]


def raw_metrics(rep: Path) -> list[dict]:
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in txt.splitlines() if l.startswith('"')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")][:80]}
        for k in KEYS:
            if k in hdr:
                d[k] = f"{r[hdr.index(k)]} {units[hdr.index(k)]}".strip()
        out.append(d)
    return out


def to_bytes(s: str) -> float:
    v, u = s.split()
    return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]


def main():
    tag, launches, rep, workload = sys.argv[1:5]
    prof = ROOT / "profiles"
    prof.mkdir(exist_ok=True)
    summary = {"tag": tag, "workload": workload}
    if launches != "-":
        summary["launch_list"] = launch_shares(Path(launches))
        (prof / f"{tag}_launches.csv").write_text(Path(launches).read_text())
    if rep != "-":
        m = raw_metrics(Path(rep))
        summary["full_capture"] = m
        tf = prof / "k1_dram_traffic.json"
        traffic = json.loads(tf.read_text()) if tf.exists() else {}
        traffic[workload] = sum(to_bytes(x["dram__bytes_read.sum"]) + to_bytes(x["dram__bytes_write.sum"]) for x in m) / len(m)
        tf.write_text(json.dumps(traffic, indent=1))
    (prof / f"{tag}_summary.json").write_text(json.dumps(summary, indent=1))
    print(json.dumps(summary, indent=1)[:2500])


if __name__ == "__main__":
    main()
