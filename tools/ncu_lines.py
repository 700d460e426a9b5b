"""Attribute ncu samples / executed instructions of a kernel to CUDA source lines (inlined-at outermost
kernel-file line) using nvdisasm -g line info.  usage: ncu_lines.py rep.ncu-rep lib.so kernel_substr file_substr"""
import csv, collections, re, subprocess, sys, io, os, tempfile, glob
rep, lib, kname, fsub = sys.argv[1:5]
tmp = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, capture_output=True)
dis = subprocess.run(['nvdisasm', '-g', '-c', glob.glob(tmp + '/*.cubin')[0]], capture_output=True, text=True).stdout
lines = dis.split('\n')
# walk the function: keep current line for file fsub (outermost inline level preferred), list of (line, innermost)
insts = []
on = False; cur_outer = None; cur_inner = None
for ln in lines:
    if ln.startswith('.text.') or '.section' in ln and '.text.' in ln:
        on = kname in ln
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        f, l, rest = m.group(1), int(m.group(2)), m.group(3)
        cur_inner = (os.path.basename(f), l)
        # "inlined at" chain follows on the same line
        chain = re.findall(r'inlined at "([^"]+)", line (\d+)', rest)
        allf = [(os.path.basename(f), l)] + [(os.path.basename(a), int(b)) for a, b in chain]
        outer = [x for x in allf if fsub in x[0]]
        cur_outer = outer[-1] if outer else allf[-1]
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m:
        insts.append((int(m.group(1), 16), cur_outer, cur_inner, m.group(2)))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
if rows[0][0] == 'Kernel Name': rows = rows[1:]
h = rows[0]; ce = h.index('Instructions Executed'); cs = h.index('# Samples')
data = [r for r in rows[1:] if len(r) > cs and r[cs].isdigit()]
assert len(data) == len(insts), (len(data), len(insts))
agg = collections.defaultdict(lambda: [0, 0]); tot = [0, 0]
for r, (addr, outer, inner, txt) in zip(data, insts):
    agg[outer][0] += int(r[cs]); agg[outer][1] += int(r[ce]); tot[0] += int(r[cs]); tot[1] += int(r[ce])
src = {}
for k in agg:
    if k and k[0] not in src:
        for root in ('pnp_admm_cnc_mri_b200/csrc',):
            p = os.path.join(root, k[0])
            if os.path.exists(p): src[k[0]] = open(p).read().split('\n')
print('samples', tot[0], 'instructions', tot[1])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 40]:
    text = src.get(k[0], [''] * 100000)[k[1] - 1].strip()[:90] if k else ''
    print(f'{100*v[0]/tot[0]:5.1f}% smp {100*v[1]/tot[1]:5.1f}% inst  {k[0] if k else None}:{k[1] if k else 0:4d}  {text}')
