"""Summarise an .ncu-rep (read here, no GPU needed): curated raw metrics + top stall locations.
usage: python tools/ncu_summary.py gpurun_out/k1_r1.ncu-rep [n_hot_lines]"""
import csv, io, subprocess, sys, collections

rep = sys.argv[1]
nhot = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__cluster', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fmaheavy', 'sm__inst_executed_pipe_fmalite',
        'smsp__sass_thread_inst_executed_op_fadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_ffma_pred_on.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_dshared_op_st.sum', 'l1tex__m_l1tex2xbar_write_sectors_mem_dshared_op_st.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct', 'lts__t_bytes.sum',
        'lts__throughput.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled', 'smsp__average_warps_issue_stalled', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__pcsamp_warps_issue_stalled']
for r in rows[2:]:
    print('=' * 100)
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(w) or w in h for w in WANT) and 'peak_sustained.' not in h:
            if 'pcsamp' in h and (v in ('0', '') ):
                continue
            print(f'{h:90s} {v:>16s} {u}')

# source page: top lines by stall samples
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if rows:
    if rows[0] and rows[0][0] == 'Kernel Name':
        rows = rows[1:]
    h = rows[0]
    def col(name):
        for i, c in enumerate(h):
            if c.strip() == name:
                return i
        return None
    ci_src, ci_samp, ci_inst = col('Source'), col('# Samples'), col('# Instructions Executed')
    ci_stall = {c: i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c}
    data = []
    for r in rows[1:]:
        try:
            data.append((int(r[ci_samp]), r))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    agg = collections.Counter()
    for s, r in data:
        for c, i in ci_stall.items():
            try:
                agg[c] += int(r[i])
            except Exception:
                pass
    print('total samples', tot)
    print('stall totals:', ', '.join(f'{k}={v} ({100*v/tot:.1f}%)' for k, v in agg.most_common(12)))
    print('--- hottest SASS lines ---')
    for s, r in sorted(data, key=lambda d: -d[0])[:nhot]:
        st = sorted(((int(r[i]) if r[i].isdigit() else 0, c) for c, i in ci_stall.items()), reverse=True)[:2]
        print(f'{s:7d} {100*s/tot:5.1f}%  {r[ci_src][:90]:90s} {st}')
