#!/bin/bash
# Round-2 GPU pass: full suite, ADE shape sweep, ncu evidence for the fused ADE kernel, host-loop check at C1.
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --tb=line ) > gpurun_out/r02_pytest_full.log 2>&1; tail -4 gpurun_out/r02_pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
for shape in "0 0" "8 8" "32 8" "16 4" "8 4" "4 8"; do
  set -- $shape
  timeout 300 python bench.py --workload c3_512_ade_slab --steps 50 --warmup 5 --no-cpu-baseline --ade-chunk $1 --ade-warps $2 > gpurun_out/r02_slab_c$1_w$2.json 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02_slab_c$1_w$2.json").read().strip().splitlines()[-1])
    print("slab chunk $1 warps $2:", round(d["value"], 1), "Gcell/s", round(d["ms_per_step"], 4), "ms")
except Exception as e:
    print("slab chunk $1 warps $2: failed", e)
PY
done
for w in c3_512_ade c3_512 c1_100 c2_200; do
  timeout 300 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_$w.json 2> gpurun_out/r02_bench_$w.err; cut -c1-200 gpurun_out/r02_bench_$w.json
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_ade_slab.csv python bench.py --workload c3_512_ade_slab --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_step_march_ade -s 20 -c 2 -o gpurun_out/r02_prof_k1ade_slab python bench.py --workload c3_512_ade_slab --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -5
