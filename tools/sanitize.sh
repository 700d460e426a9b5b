#!/bin/bash
# compute-sanitizer over the hand-rolled synchronisation of K1 (mbarrier + st.async DSMEM transposes, release/acquire
# hand-off through L2) and K2 (cp.async.bulk / TMA + mbarrier).  Logs -> gpurun_out/sanitizer/ (copied to profiles/).
set -u
out=gpurun_out/sanitizer; mkdir -p $out
for job in memcheck:k1 memcheck:hybrid memcheck:fused memcheck:k2 racecheck:k1 racecheck:fused racecheck:k2 synccheck:k2 synccheck:k1; do
    tool=${job%%:*}; w=${job##*:}
    timeout 360 compute-sanitizer --tool $tool --error-exitcode 1 --print-limit 20 python tools/sanitize_run.py $w > $out/${tool}_$w.log 2>&1
    echo "$tool $w rc=$?" | tee -a $out/summary.txt
    tail -4 $out/${tool}_$w.log | sed 's/^/    /' | tee -a $out/summary.txt
done
