#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --tb=line ) > gpurun_out/r02_pytest_run6.log 2>&1; tail -4 gpurun_out/r02_pytest_run6.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], ":", round(d["value"], 1), "Gcell/s  e2e", round(d["e2e"]["value"], 1), " ms/step", round(d["ms_per_step"], 4))
except Exception as e:
    print(sys.argv[2], "failed", e, open(sys.argv[1]).read()[-300:])
PY
}
for lay in 0 1 3; do
  timeout 300 python bench.py --workload c3_512_ade --steps 100 --warmup 5 --no-cpu-baseline --ade-layout $lay > gpurun_out/r02c_sphere_l$lay.json 2>&1; show gpurun_out/r02c_sphere_l$lay.json "sphere layout $lay"
done
for lay in 0 1 2; do
  timeout 300 python bench.py --workload c3_512_ade_slab --steps 100 --warmup 5 --no-cpu-baseline --ade-layout $lay > gpurun_out/r02c_slab_l$lay.json 2>&1; show gpurun_out/r02c_slab_l$lay.json "slab layout $lay"
done
timeout 300 python bench.py --workload c3_512 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_c3_512.json 2>&1; show gpurun_out/r02c_c3_512.json "c3_512 no material"
for k in 200 1000 4000; do
  timeout 300 python bench.py --workload c1_100 --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r02c_c1_$k.json 2>&1; show gpurun_out/r02c_c1_$k.json "c1_100 steps $k"
done
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r02c_default.json 2>&1; show gpurun_out/r02c_default.json "default c5"
