#!/usr/bin/env python
"""Static per-source-line opcode counts of one kernel: sass_lines.py file.o kernel_substring src.cu first_line last_line"""
import collections, re, subprocess, sys, os, tempfile
obj, sub, src, lo, hi = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
d = tempfile.mkdtemp()
subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [f for f in os.listdir(d) if f.endswith('.cubin')][0]
txt = subprocess.run(['nvdisasm', '--print-line-info', os.path.join(d, cub)], capture_output=True, text=True).stdout
secs = txt.split('\t.section\t.text.')
sec = [s for s in secs if s.startswith(sub) or sub in s.split('\n')[0]][0]
cur = None
per = collections.defaultdict(collections.Counter)
base = os.path.basename(src)
for l in sec.split('\n'):
    if '//## File' in l:
        m = re.search(r'File "([^"]+)", line (\d+)', l)
        cur = int(m.group(2)) if m and m.group(1).endswith(base) else None
        continue
    m = re.match(r'\s+(?:/\*[0-9a-f]+\*/)?\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]+)', l)
    if m and cur:
        op = m.group(1)
        per[cur][op if op.startswith('IMAD') else op.split('.')[0]] += 1
lines = open(src).read().split('\n')
tot = collections.Counter()
for ln in range(lo, hi + 1):
    if ln in per:
        print(ln, sum(per[ln].values()), dict(per[ln]), '|', lines[ln - 1].strip()[:70])
        tot.update(per[ln])
print('total', sum(tot.values()), dict(tot))
