#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --tb=line -x ) > gpurun_out/r02_pytest_run5.log 2>&1; tail -4 gpurun_out/r02_pytest_run5.log
for w in c3_512_ade_slab c3_512_ade c3_512; do
  timeout 300 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_$w.json 2> gpurun_out/r02b_bench_$w.err; cut -c1-200 gpurun_out/r02b_bench_$w.json
done
for k in 200 1000 4000; do
  timeout 300 python bench.py --workload c1_100 --steps $k --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_c1_$k.json 2> gpurun_out/r02b_bench_c1_$k.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r02b_bench_c1_$k.json").read().strip().splitlines()[-1])
print("c1 steps $k: device", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "us/step", round(d["ms_per_step"] * 1e3, 3))
PY
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_step_march_ade -s 20 -c 1 -o gpurun_out/r02b_prof_k1ade_slab python bench.py --workload c3_512_ade_slab --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -3
