#!/usr/bin/env python
"""Copy the evidence of tools/validate_round.sh from gpurun_out/ (scratch) into profiles/ (tracked): bench lines, test
logs, ncu launch-list shares and the key raw metrics of the full captures.  Usage: tools/collect_profiles.py <round tag>"""
import json
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
for f in sorted(OUT.glob("val_bench_*.json")):
    lines = [q for q in f.read_text().strip().splitlines() if q.startswith("{")]
    if lines:
        (PROF / f"{tag}_{f.name[4:]}").write_text(lines[-1] + "\n")
for name in ("val_pytest.log", "val_smoke.log"):
    if (OUT / name).exists():
        shutil.copyfile(OUT / name, PROF / f"{tag}_{name[4:]}")
jobs = [("default_c5", "val_launches_default.csv", "val_prof_k1_default.ncu-rep", "c5_weak"),
        ("c3_512_ade_slab", "val_launches_ade_slab.csv", "val_prof_k1ade_slab.ncu-rep", "c3_512_ade_slab_k1ade"),
        ("c3_512_ade_sphere", "val_launches_ade_sphere.csv", None, "c3_512_ade"),
        ("c1_100_k5", None, "val_prof_k5_c1.ncu-rep", "c1_100_k5")]
for name, csv, rep, workload in jobs:
    a = str(OUT / csv) if csv and (OUT / csv).exists() else "-"
    b = str(OUT / rep) if rep and (OUT / rep).exists() else "-"
    if a == "-" and b == "-":
        continue
    if rep and b == "-" and (PROF / f"{tag}_{name}_summary.json").exists():
        continue                             # a pass without the full ncu capture must not drop the one already kept
    subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_summary.py"), f"{tag}_{name}", a, b, workload], check=False,
                   stdout=subprocess.DEVNULL)
print(sorted(p.name for p in PROF.glob(f"{tag}_*")))
