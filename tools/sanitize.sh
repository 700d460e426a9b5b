#!/bin/bash
# compute-sanitizer over the hand-rolled synchronisation of K1 (mbarrier + st.async DSMEM transposes, release/acquire
# hand-off through L2) and K2 (cp.async.bulk / TMA + mbarrier).  Logs -> gpurun_out/sanitizer/ (copied to profiles/).
set -u
# the planner calibration (first library call of a process) runs ~6 ms of kernels: hours under racecheck. The production kernels are
# exercised by sanitize_run.py itself.
export PNPADMM_NO_CALIBRATE=1
out=gpurun_out/sanitizer; mkdir -p $out
for job in memcheck:k5 memcheck:k1 memcheck:hybrid memcheck:fused memcheck:k3 memcheck:k2 racecheck:k1 racecheck:fused racecheck:k3 racecheck:k2 synccheck:k2 synccheck:k3 synccheck:k1; do
    tool=${job%%:*}; w=${job##*:}
    timeout 180 compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 python tools/sanitize_run.py $w > $out/${tool}_$w.log 2>&1
    echo "$tool $w rc=$?" | tee -a $out/summary.txt
    tail -4 $out/${tool}_$w.log | sed 's/^/    /' | tee -a $out/summary.txt
done
