#!/usr/bin/env python
"""Per-phase instruction mix by issue pipe from an ncu report: ncu_pipes.py report.ncu-rep src.cu 'name:first-last' ... (line ranges of src.cu; everything else -> 'other')"""
import csv, io, subprocess, sys, collections
rep, srcfile = sys.argv[1], sys.argv[2]
phases = []
for a in sys.argv[3:]:
    n, r = a.split(':'); lo, hi = r.split('-'); phases.append((n, int(lo), int(hi)))
ALU = {'LOP3','SHF','ISETP','VIADD','SEL','LEA','VIADDMNMX','VIMNMX','IADD3','PRMT','IABS','FLO','BREV','IMNMX','VIMNMX3','PLOP3','P2R','R2P','BMSK','SGXT','VABSDIFF','ICMP','FSEL','FSETP','I2FP','F2FP','LOP','SHL','SHR','IADD','MOV','CS2R','VOTE'}
def pipe(op):
    b = op.split('.')[0]
    if b.startswith('IMAD') or b in ('FFMA','FMUL','FADD','IDP'): return 'fma'
    if b in ALU: return 'alu'
    if b in ('LDS','STS','LDG','STG','ATOMS','ATOM','ATOMG','RED','LDL','STL','LD','ST','LDSM','SHFL','MATCH','REDUX'): return 'lsu/shfl'
    if b in ('POPC','MUFU','F2I','I2F','F2F','FRND','DADD','DMUL','DFMA','DSETP'): return 'xu/fp64'
    if b in ('BRA','BSSY','BSYNC','EXIT','WARPSYNC','NOP','CALL','RET','BAR','YIELD','JMP','BRX','BREAK'): return 'ctrl'
    if b.startswith('U') or b in ('S2UR','LDCU','S2R','LDC','R2UR'): return 'unif/const'
    return 'misc'
src = subprocess.run(['ncu','-i',rep,'--page','source','--print-source','cuda,sass','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur_file = None; cur_line = None; iI = None
agg = collections.defaultdict(collections.Counter)
for r in rows:
    if r and r[0] == 'File Path': cur_file = r[1]; continue
    if 'Instructions Executed' in r: iI = r.index('Instructions Executed'); iS = 3; continue
    if iI is None or len(r) <= iI: continue
    if r[2] == '-':
        try: cur_line = int(r[0])
        except ValueError: cur_line = None
        continue
    s = r[3].strip()
    if s.startswith('@'): s = s.split(None, 1)[1] if ' ' in s else s
    op = s.split()[0] if s else '?'
    ph = 'other'
    if cur_file and cur_file.endswith(srcfile) and cur_line:
        for n, lo, hi in phases:
            if lo <= cur_line <= hi: ph = n; break
    elif cur_file: ph = 'hdr:' + cur_file.split('/')[-1]
    try: val = float(r[iI] or 0)
    except ValueError: continue
    agg[ph][pipe(op)] += val
tot = sum(sum(c.values()) for c in agg.values())
E = float(__import__('os').environ.get('EDGES', '1'))
print('%-26s %8s %8s %8s %8s %8s %8s %8s   (warp instr / edge)' % ('phase', 'alu', 'fma', 'lsu/shfl', 'xu/fp64', 'ctrl', 'unif', 'total'))
sumc = collections.Counter()
for ph, c in sorted(agg.items(), key=lambda kv: -sum(kv[1].values())):
    t = sum(c.values()); sumc.update(c)
    print('%-26s %8.2f %8.2f %8.2f %8.2f %8.2f %8.2f %8.2f' % (ph[:26], c['alu']/E, c['fma']/E, c['lsu/shfl']/E, c['xu/fp64']/E, c['ctrl']/E, (c['unif/const']+c['misc'])/E, t/E))
print('%-26s %8.2f %8.2f %8.2f %8.2f %8.2f %8.2f %8.2f' % ('TOTAL', sumc['alu']/E, sumc['fma']/E, sumc['lsu/shfl']/E, sumc['xu/fp64']/E, sumc['ctrl']/E, (sumc['unif/const']+sumc['misc'])/E, tot/E))
