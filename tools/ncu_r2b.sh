#!/bin/bash
# Round-2 (second half) ncu captures of the K5 variants: a dilated 64->64 layer, and the launches of one FFDNet forward.
set -u
out=gpurun_out/ncu_r2b; mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:conv64_tc_kernel -s 2 -c 1 -o $out/k5_conv64_dil3 -f python tools/conv64_time.py 256 3 3 > $out/k5_conv64_dil3.log 2>&1
# second forward: pack, thin first layer, 13 middle layers, tail = 16 launches; capture pack + first layer + one middle layer + tail by name / order
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second --clock-control none -k regex:"conv64_tc_kernel|ffdnet_pack_kernel" -s 16 -c 16 --csv --log-file $out/ffdnet_forward_launches.csv python tools/ffdnet_time.py 256 > $out/ffdnet_forward.log 2>&1
# launch list of the final bench (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/launches_bench.csv python bench.py --steps 3 --warmup 3 > $out/bench_under_ncu.log 2>&1
ls -la $out
