#!/bin/bash
# 1 GPU: K1-ADE with the uniform fast paths and K5 with branch-free z faces: parity, benches, ncu evidence
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --tb=line ) > gpurun_out/r02_pytest_run7.log 2>&1; tail -4 gpurun_out/r02_pytest_run7.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], ":", round(d["value"], 1), "Gcell/s  e2e", round(d["e2e"]["value"], 1), " ms/step", round(d["ms_per_step"], 5), "frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[2], "failed", e, open(sys.argv[1]).read()[-300:])
PY
}
for w in c3_512_ade_slab c3_512_ade c3_512 c2_200 c4_enclosure; do
  timeout 300 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_$w.json 2>&1; show gpurun_out/r02d_$w.json "$w"
done
timeout 300 python bench.py --workload c3_512_ade --ade-layout 3 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02d_sphere_fused.json 2>&1; show gpurun_out/r02d_sphere_fused.json "sphere fused"
for k in 1000 4000; do
  timeout 300 python bench.py --workload c1_100 --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r02d_c1_$k.json 2>&1; show gpurun_out/r02d_c1_$k.json "c1_100 steps $k"
done
timeout 300 python tools/resident_bench.py > gpurun_out/r02d_resident_bench.jsonl 2>&1; tail -8 gpurun_out/r02d_resident_bench.jsonl | cut -c1-200
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_step_march_ade -s 20 -c 1 -o gpurun_out/r02d_prof_k1ade_slab python bench.py --workload c3_512_ade_slab --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02d_launches_sphere.csv python bench.py --workload c3_512_ade --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02d_launches_slab.csv python bench.py --workload c3_512_ade_slab --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:k5_resident -s 2 -c 1 -o gpurun_out/r02d_prof_k5_c1 python bench.py --workload c1_100 --steps 256 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls gpurun_out | grep r02d | head -30
