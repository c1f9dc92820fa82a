#!/bin/bash
set -u
show() { python tools/show_bench.py "$1"; }
for rep in 1 2 3; do python bench.py --workload c3_512_ade --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | show "sphere auto rep$rep"; done
python bench.py --workload c3_512_ade --steps 100 --warmup 5 --no-cpu-baseline --rows 1 --warps-j 4 --warps-k 1 --chunk-i 16 2>/dev/null | show "sphere r1 wj4 wk1 c16"
python bench.py --workload c3_512_ade --steps 100 --warmup 5 --no-cpu-baseline --rows 1 --warps-j 2 --warps-k 2 --chunk-i 16 2>/dev/null | show "sphere r1 wj2 wk2 c16"
python bench.py --workload c3_512_ade --steps 100 --warmup 5 --no-cpu-baseline --rows 2 --warps-j 8 --warps-k 1 --chunk-i 16 2>/dev/null | show "sphere r2 wj8 wk1 c16"
python bench.py --workload c3_512 --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | show "c3_512 auto"
python bench.py --workload c3_512 --steps 100 --warmup 5 --no-cpu-baseline --rows 1 --warps-j 4 --warps-k 1 --chunk-i 16 2>/dev/null | show "c3_512 r1 wj4 wk1 c16"
python bench.py --workload c3_512_ade --ade-layout 3 --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | show "sphere fused"
for rep in 1 2; do python bench.py --workload c3_512_ade_slab --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | show "slab auto rep$rep"; done
