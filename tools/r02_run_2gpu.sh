#!/bin/bash
# Two B200s: the real multi-rank tests (NCCL send/recv with the overlapped schedule, peer-store halo) and bench --gpus 2
# with its decomposition self-check, in both halo modes.  Logs are kept under profiles/ ("keep the log").
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/r02_2gpu_devices.txt
( time timeout 1200 python -m pytest tests/test_multi_gpu.py -q -m gpu -rs --tb=short -k "two_ranks or two_processes" ) > gpurun_out/r02_2gpu_pytest.log 2>&1; tail -8 gpurun_out/r02_2gpu_pytest.log
for halo in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 --halo $halo > gpurun_out/r02_scale2_$halo.json 2> gpurun_out/r02_scale2_$halo.err
  tail -1 gpurun_out/r02_scale2_$halo.json | cut -c1-300; grep -o '"parity_n": {[^}]*}' gpurun_out/r02_scale2_$halo.json
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale1.json 2> gpurun_out/r02_scale1.err; tail -1 gpurun_out/r02_scale1.json | cut -c1-200
# the unchanged reference script on two GPUs through the module runner
mkdir -p /tmp/two && cd /tmp/two && PYTHONPATH=$GRAFT_REPO_ROOT timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29563 -m strata_fdtd_b200 $GRAFT_REPO_ROOT/oracle/_ref/examples/basic_pulse.py > $GRAFT_REPO_ROOT/gpurun_out/r02_2gpu_basic_pulse.log 2>&1; cd $GRAFT_REPO_ROOT
python - <<'PY'
import json, numpy as np
z = np.load("/tmp/two/results.h5"); g = np.load("tests/golden/script_basic_pulse.npz")
print("basic_pulse.py on 2 GPUs: trace equals the reference's:", bool(np.array_equal(z["probes/downstream"], g["probe_downstream"])),
      json.loads(str(z["__attrs__"])).get("metadata@num_gpus"))
PY
