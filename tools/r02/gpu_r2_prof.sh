#!/bin/bash
# round 2 ncu captures (--set full, one launch each): SuperPoint candidate kernel (MODE 4) + one-column re-rank,
# RANSAC under 50 % outliers, the XOR/popc Hamming kernel
mkdir -p gpurun_out
B="python bench.py --images 48 --steps 1 --warmup 1 --no-stages --no-e2e --no-cpu-baseline --no-configs"
NCU="ncu --set full --clock-control none --import-source on"
$B --kind superpoint > gpurun_out/r2_prof_plain_sp.log 2>&1 && $NCU -k regex:"l2_i8x2_kernel|l2f_rerank1" -s 20 -c 2 -f -o gpurun_out/r2_prof_sp $B --kind superpoint > gpurun_out/r2_prof_sp.log 2>&1; echo "sp $?"
$B --kind sift --outlier-frac 0.5 > gpurun_out/r2_prof_plain_rs.log 2>&1 && $NCU -k regex:fmat_ransac -s 10 -c 1 -f -o gpurun_out/r2_prof_ransac $B --kind sift --outlier-frac 0.5 > gpurun_out/r2_prof_ransac.log 2>&1; echo "ransac $?"
$B --kind orb --debug-flags 1024 > gpurun_out/r2_prof_plain_ham.log 2>&1 && $NCU -k regex:hamming_top2 -s 10 -c 1 -f -o gpurun_out/r2_prof_hamming $B --kind orb --debug-flags 1024 > gpurun_out/r2_prof_hamming.log 2>&1; echo "hamming $?"
ls -la gpurun_out/*.ncu-rep
