timeout 1500 python -m pytest tests/test_parity_gpu.py -q -x -s -k "rowsep" 2>&1 | grep -E "rowsep|passed|failed|Error|error|FAILED" | tail -25
PNPADMM_K3_K1CODE=1 timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "rowsep_kernel_kat or rowsep_kernel_refuses" 2>&1 | tail -2
bash tools/ncu_r2.sh > gpurun_out/ncu_r2.log 2>&1; tail -3 gpurun_out/ncu_r2.log
