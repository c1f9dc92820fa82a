#!/bin/bash
# Round-end style validation on one B200: full GPU test suite, smoke(), benches of every BASELINE configuration, ncu launch
# lists and full captures of the kernels this round changed.  Everything lands in gpurun_out/ (tools/collect_profiles.py
# copies the summaries into profiles/).
set -u
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -q -m gpu --tb=line -rs ) > gpurun_out/val_pytest.log 2>&1; tail -8 gpurun_out/val_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/val_smoke.log 2>&1; tail -6 gpurun_out/val_smoke.log
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/val_bench_reference.json 2> gpurun_out/val_bench_reference.err; cut -c1-300 gpurun_out/val_bench_reference.json
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/val_bench_default.json 2> gpurun_out/val_bench_default.err; cut -c1-300 gpurun_out/val_bench_default.json
for w in c1_100 c2_200 c3_512 c3_512_ade c3_512_ade_slab c4_enclosure; do
  timeout 300 python bench.py --workload $w --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/val_bench_$w.json 2> gpurun_out/val_bench_$w.err; cut -c1-230 gpurun_out/val_bench_$w.json
done
timeout 300 python bench.py --workload c1_100 --steps 4000 --warmup 5 --no-cpu-baseline > gpurun_out/val_bench_c1_100_4000.json 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/val_launches_default.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/val_launches_ade_slab.csv python bench.py --workload c3_512_ade_slab --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/val_launches_ade_sphere.csv python bench.py --workload c3_512_ade --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
if [ -n "${SKIP_NCU_FULL:-}" ]; then ls -la gpurun_out | grep val_ | awk '{print $5, $9}'; exit 0; fi
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_step_march_ade -s 20 -c 1 -o gpurun_out/val_prof_k1ade_slab python bench.py --workload c3_512_ade_slab --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k1_step_march -s 45 -c 1 -o gpurun_out/val_prof_k1_default python bench.py --steps 4 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 250 ncu --set full --clock-control none --import-source on -k regex:k5_resident -s 8 -c 1 -o gpurun_out/val_prof_k5_c1 python bench.py --workload c1_100 --steps 256 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out | grep val_ | awk '{print $5, $9}'
