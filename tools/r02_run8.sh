#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_config_gates_gpu.py -q -m gpu --tb=line -k "ade or c3" ) > gpurun_out/r02_pytest_run8.log 2>&1; tail -3 gpurun_out/r02_pytest_run8.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], ":", round(d["value"], 1), "Gcell/s  e2e", round(d["e2e"]["value"], 1), " ms/step", round(d["ms_per_step"], 5), "frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print(sys.argv[2], "failed", e, open(sys.argv[1]).read()[-300:])
PY
}
for occ in 2 3; do for ch in 0 16; do
  timeout 300 python bench.py --workload c3_512_ade_slab --steps 100 --warmup 5 --no-cpu-baseline --ade-occ $occ --ade-chunk $ch > gpurun_out/r02e_slab_o${occ}_c$ch.json 2>&1; show gpurun_out/r02e_slab_o${occ}_c$ch.json "slab occ $occ chunk $ch"
done; done
timeout 300 python bench.py --workload c3_512_ade --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_sphere.json 2>&1; show gpurun_out/r02e_sphere.json "sphere auto (lists beside K1, priority stream)"
timeout 300 python bench.py --workload c3_512_ade --steps 100 --warmup 5 --no-cpu-baseline --ade-layout 3 > gpurun_out/r02e_sphere_fused.json 2>&1; show gpurun_out/r02e_sphere_fused.json "sphere fused"
timeout 300 python bench.py --workload c3_512 --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_c3_512.json 2>&1; show gpurun_out/r02e_c3_512.json "c3_512"
