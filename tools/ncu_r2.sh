#!/bin/bash
# Round-2 ncu captures (one GPU; numbers printed under ncu are never bench values).  Reports -> gpurun_out/ncu_r2/, read with
# tools/ncu_summary.py on the build container and summarised under profiles/r2_*.
set -u
out=gpurun_out/ncu_r2; mkdir -p $out
export PNPADMM_NO_CALIBRATE=1   # the planner calibration launches its own K1 / K2 kernels first: keep them out of the captures
NCU="ncu --set full --clock-control none --import-source on"
# K1 with the fused prologue: one wave (28 images = 14 planes = 14 clusters), 10 iterations
$NCU -k regex:cluster256_kernel -s 1 -c 1 -o $out/k1_fused -f python tools/prof_reconstruct.py random 28 10 cluster > $out/k1_fused.log 2>&1
# K3 row-separable kernel: B = 64 (the bench batch), 50 iterations
$NCU -k regex:rowsepN_kernel -s 1 -c 1 -o $out/k3_rowsepN256 -f python tools/prof_reconstruct.py cartesian 64 50 > $out/k3_rowsep.log 2>&1
# K2 at N = 1024, B = 64: one rows pass + one columns pass of the iteration (DRAM traffic for roofline_streaming)
$NCU -k regex:"cols2_tma_kernel|rows2_kernel" -s 4 -c 2 -o $out/k2_1024 -f python tools/k2_bench.py 1024 64 4 > $out/k2_1024.log 2>&1
# K5 one 64->64 layer at B = 256
$NCU -k regex:conv64_tc_kernel -s 2 -c 1 -o $out/k5_conv64 -f python tools/conv64_time.py > $out/k5_conv64.log 2>&1
# launch list of the bench (shares, not absolutes)
unset PNPADMM_NO_CALIBRATE
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/launches_bench.csv python bench.py --steps 3 --warmup 3 > $out/bench_under_ncu.log 2>&1
ls -la $out
