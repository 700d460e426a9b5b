timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "rowsep or hybrid or fused" 2>&1 | tail -3
python - <<'PY'
import ctypes, sys
sys.path.insert(0, '.')
from pnp_admm_cnc_mri_b200 import _abi
lib = _abi.load()
c = (ctypes.c_double * 4)(); cal = ctypes.c_int()
lib.pnpadmm_debug_plan_constants.argtypes = [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
import torch; torch.cuda.init()
import time; t0 = time.time()
print('rc', lib.pnpadmm_debug_plan_constants(c, cal), 'constants tau1, k2_a, k2_b, k2_pro:', list(c), 'calibrated', cal.value, 'first-call s', time.time() - t0)
p = [ctypes.c_int() for _ in range(6)]
lib.pnpadmm_plan_info(64, 256, 0, 50, 0, *p); print('plan B=64:', [v.value for v in p])
lib.pnpadmm_plan_info(1024, 256, 0, 50, 0, *p); print('plan B=1024:', [v.value for v in p])
PY
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo bench rc=$?; tail -3 gpurun_out/r2_bench4.err
PNPADMM_NO_CALIBRATE=1 timeout 600 python bench.py --steps 20 --warmup 5 --legs headline > gpurun_out/r2_bench4_nocal.json 2> gpurun_out/r2_bench4_nocal.err; echo bench rc=$?
