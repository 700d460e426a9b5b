#!/usr/bin/env python
"""Refresh profiles/traffic.json from `ncu --set full` reports (run HERE, no GPU needed to read a report):

    python tools/update_traffic.py k2 gpurun_out/ncu_r2c/k2_1024.ncu-rep      # one rows2_kernel<1024> + one cols2_tma_kernel<1024> launch
    python tools/update_traffic.py k5 gpurun_out/ncu_r2c/k5_conv64.ncu-rep    # one conv64_tc_kernel<64> launch at B = 256

The reports come from tools/ncu_r2.sh (K2: `-k regex:"cols2_tma_kernel|rows2_kernel" -s 4 -c 2 python tools/k2_bench.py 1024 64 4`,
K5: `-k regex:conv64_tc_kernel -s 2 -c 1 python tools/conv64_time.py 256 3`).  The record stores dram__bytes_read.sum + dram__bytes_write.sum
per launch and a hash of the kernel sources it was captured on; bench.py reports `traffic` only while that hash matches (a capture must be
redone after ANY edit of those files, even one that cannot change the traffic).  The tool stamps the CURRENT sources' hash: only feed it
reports captured from the library built from the current sources."""
import csv, hashlib, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, 'pnp_admm_cnc_mri_b200', 'csrc')
KEYS = {'k2': ('k2_n1024_b64_rows2_plus_cols2', ('rows2_kernel<1024', 'cols2_tma_kernel<1024')),
        'k5': ('k5_conv64_b256', ('conv64_tc_kernel<64>',))}


def source_sha(files):
    h = hashlib.sha256()
    for f in files:
        with open(os.path.join(CSRC, f), 'rb') as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def to_bytes(value, unit):
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]
    return float(value.replace(',', '')) * scale


def main():
    which, rep = sys.argv[1], sys.argv[2]
    key, kernels = KEYS[which]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ik, ir, iw, it = (hdr.index(n) for n in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum'))
    total, parts, seen = 0.0, [], set()
    for r in rows[2:]:
        for k in kernels:
            if k in r[ik] and k not in seen:                      # the first launch of each kernel in the report
                seen.add(k)
                rd, wr = to_bytes(r[ir], units[ir]), to_bytes(r[iw], units[iw])
                total += rd + wr
                parts.append(f'{r[ik].split("(")[0].replace("void ", "")}: read {rd / 1e6:.1f} MB + write {wr / 1e6:.1f} MB, {r[it]} {units[it]}')
    if len(seen) != len(kernels):
        sys.exit(f'{rep}: expected launches of {kernels}, found {sorted(seen)}')
    path = os.path.join(ROOT, 'profiles', 'traffic.json')
    rec = json.load(open(path))
    rec[key]['dram_bytes'] = total
    rec[key]['sources_sha16'] = source_sha(rec[key]['sources'])
    rec[key]['capture'] = f'{os.path.relpath(rep, ROOT)} (tools/update_traffic.py): ' + '; '.join(parts)
    json.dump(rec, open(path, 'w'), indent=1)
    print(key, total, rec[key]['sources_sha16'])


if __name__ == '__main__':
    main()
