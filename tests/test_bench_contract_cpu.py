"""CPU: the reference arm of bench.py (`--impl reference`: the oracle restatement of the reference's NumPy loop on the host
cores) prints one JSON line with the keys the bench contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1'],
                         check=True, capture_output=True, text=True, timeout=600, cwd=ROOT).stdout
    line = json.loads(out.strip().splitlines()[-1])
    assert line['impl'] == 'reference'
    assert line['metric'] == 'admm_cnc_iterations_per_s' and line['unit'] == 'iterations/s'
    assert line['higher_is_better'] is True and line['value'] > 0 and line['steps'] == 1
    assert line['config']['workload'].startswith('BASELINE config 2')
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == line['value'] and 'sample' in cb
    e2e = line['e2e']
    assert e2e['value'] == line['value'] and e2e['h2d_bytes_per_step'] == 0 and e2e['d2h_bytes_per_step'] == 0
    assert line['gpu_launches'] == 0 and line['vs_baseline'] is None
