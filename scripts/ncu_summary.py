#!/usr/bin/env python
"""Summarise an ncu report: key raw metrics of each kernel + per-CUDA-line instruction / stall-sample shares.
usage: ncu_summary.py report.ncu-rep [min_share_pct]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg',
        'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.max']
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d.get('Kernel Name'))
    for k in KEYS:
        if k in d:
            print('   %-72s %s %s' % (k, d[k], rows[1][hdr.index(k)]))
    for k in hdr:
        if 'warp_issue_stalled' in k and k.endswith('_per_warp_active.pct') and 'not_issued' not in k:
            try:
                if float(d[k]) > 3: print('   stall %-66s %s' % (k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_warp_active.pct', ''), d[k]))
            except ValueError:
                pass
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = None
for i, r in enumerate(rows):
    if 'Instructions Executed' in r:
        h = i
        break
if h is not None:
    hdr = rows[h]
    iI, iT, iS = hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
    agg = [r for r in rows[h + 1:] if len(r) > iI and r[2] == '-']
    tot = sum(float(r[iI] or 0) for r in agg) or 1
    ts = sum(float(r[iS] or 0) for r in agg) or 1
    print('-- per CUDA line: total warp instructions %.0f, samples %.0f' % (tot, ts))
    for r in agg:
        v, s = float(r[iI] or 0), float(r[iS] or 0)
        if 100 * v / tot > thr or 100 * s / ts > thr:
            print('%5s inst %5.2f%% thr/inst %5.1f samp %5.2f%% | %s' % (r[0], 100 * v / tot, float(r[iT] or 0) / max(v, 1), 100 * s / ts, r[1][:120]))
