#!/bin/bash
set -u
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -q -m gpu --tb=line -rs ) > gpurun_out/r02_pytest_run10.log 2>&1; tail -12 gpurun_out/r02_pytest_run10.log | cut -c1-300
timeout 300 python tools/resident_bench.py > gpurun_out/r02f_resident_bench.jsonl 2>&1; grep -E '"n": (64|100), "geometry": false, "kernel": "resident_nosplit"' gpurun_out/r02f_resident_bench.jsonl | cut -c1-200
for w in c1_100 c3_512_ade c3_512_ade_slab; do python bench.py --workload $w --steps 1000 --warmup 5 --no-cpu-baseline 2>/dev/null | python tools/show_bench.py "$w"; done
