#!/bin/bash
# N B200s of one box: bench --gpus N in both halo modes (with the decomposition self-check) and the N = 1 line of the same box
set -u
N=$1
mkdir -p gpurun_out
for halo in p2p nccl; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$N bench.py --gpus $N --steps 20 --warmup 5 --halo $halo > gpurun_out/r02_scale${N}_$halo.json 2> gpurun_out/r02_scale${N}_$halo.err
  tail -1 gpurun_out/r02_scale${N}_$halo.json | python tools/show_bench.py "N=$N $halo"; grep -o '"parity_n": {[^}]*}' gpurun_out/r02_scale${N}_$halo.json
done
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale${N}_n1.json 2>/dev/null; tail -1 gpurun_out/r02_scale${N}_n1.json | python tools/show_bench.py "N=1 same box"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2958$N bench.py --impl reference --gpus $N --steps 5 --warmup 2 > gpurun_out/r02_scale${N}_reference.json 2>/dev/null; cut -c1-400 gpurun_out/r02_scale${N}_reference.json
