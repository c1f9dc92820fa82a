#!/bin/bash
# Round-end style validation on one B200: full GPU test suite, smoke(), benches, ncu launch lists and captures.
set -u
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -x -q -m gpu ) > gpurun_out/val_pytest.log 2>&1; tail -3 gpurun_out/val_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1; tail -6 gpurun_out/val_smoke.log
timeout 600 python bench.py > gpurun_out/val_bench_default.json 2> gpurun_out/val_bench_default.err; cut -c1-400 gpurun_out/val_bench_default.json
timeout 300 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/val_bench_reference.json 2>&1; cut -c1-300 gpurun_out/val_bench_reference.json
for w in c1_100 c2_200 c3_512 c3_512_ade c3_512_ade_slab c4_enclosure; do
  timeout 300 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/val_bench_$w.json 2> gpurun_out/val_bench_$w.err; cut -c1-330 gpurun_out/val_bench_$w.json
done
# launch lists (kernel shares) and one full capture of the two new kernels
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c1_100_r01b.csv python bench.py --workload c1_100 --steps 64 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_200_r01b.csv python bench.py --workload c2_200 --steps 32 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:k5_resident -s 2 -c 1 -o gpurun_out/prof_k5_c1_final python bench.py --workload c1_100 --steps 256 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:k6_pipeline -s 2 -c 1 -o gpurun_out/prof_k6_c2_final python bench.py --workload c2_200 --steps 64 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | tail -12
