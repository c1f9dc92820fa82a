#!/bin/bash
# Builds tools/bin/k1_ab_r01 (round-1 kernels, from git history) and tools/bin/k1_ab_r02 (current kernels): the same
# harness (tools/k1_ab.cu) against two revisions of sb_kernels.cuh, to be run back to back on ONE box.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin/ab_r01
git show 696b0a4:strata_fdtd_b200/csrc/sb_kernels.cuh > tools/bin/ab_r01/sb_kernels.cuh
F="-gencode arch=compute_100a,code=sm_100a -O3 --fmad=false -std=c++17 -DKV=4"
nvcc $F -Itools/bin/ab_r01 -o tools/bin/k1_ab_r01 tools/k1_ab.cu
nvcc $F -Istrata_fdtd_b200/csrc -o tools/bin/k1_ab_r02 tools/k1_ab.cu
# register-target experiments (DESIGN.md 6, k1_min_blocks in sb_kernels.cuh): the current kernels with ptxas aiming at an
# unspecified number / 2 / 3 blocks of 256 threads per SM, and the table as shipped.  Run e.g.
#   tools/bin/k1_ab_m2 1024 512 512 20 1 16 <variant 0-9> 2 2      (nx ny nz steps rows chunk variant warps_j warps_k)
for m in 0 2 3; do nvcc $F -DSB_K1_MINB=$m -DSB_K1_RTBOX_NONE -Istrata_fdtd_b200/csrc -o tools/bin/k1_ab_m$m tools/k1_ab.cu; done
nvcc $F -Istrata_fdtd_b200/csrc -o tools/bin/k1_ab_fin tools/k1_ab.cu
ls -la tools/bin
