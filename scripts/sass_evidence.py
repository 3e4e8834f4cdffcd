#!/usr/bin/env python
"""SASS evidence for the bulk-async (TMA 1-D) staging claims: per kernel of libporrt_b200.so, how often the mnemonics that prove
them occur (UBLKCP = cp.async.bulk global -> shared, SYNCS = mbarrier arrive / try_wait, ATOMS = shared-memory atomics of the
column solver, LDG.E.256 = the 256-bit bitmap loads of the edge kernel).  Writes profiles/sass_<tag>.txt.
usage: python scripts/sass_evidence.py [round tag, default r2]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "po_rrt_b200", "libporrt_b200.so")
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").split("(")[0]
PAT = {"UBLKCP": r"\bUBLKCP", "SYNCS": r"\bSYNCS", "ATOMS": r"\bATOMS", "ATOMG/RED": r"\b(ATOMG|RED)\b", "LDG.E.256": r"LDG\.E\.(ENL2\.)?256",
       "LDS.128": r"LDS\.128", "DADD": r"\bDADD\b", "DSETP": r"\bDSETP", "SHFL": r"\bSHFL", "BAR": r"\bBAR\.", "total": r"^\s+/\*[0-9a-f]{4}\*/"}
cur, counts = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = demangle(m.group(1))
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for k, p in PAT.items():
        if re.search(p, line):
            counts[cur][k] += 1
arch = re.findall(r"arch = (sm_\w+)", sass)
out = ["# SASS mnemonic counts per kernel of po_rrt_b200/libporrt_b200.so (cuobjdump -sass), arch %s" % sorted(set(arch)),
       "# columns: " + " ".join(PAT)]
groups = [("edge kernel (edge3.cu): class plane staged by cp.async.bulk + mbarrier, 256-bit bitmap loads", "edge_validity_v3"),
          ("NN tiles (nn_tile.cu): vertex tiles staged by cp.async.bulk + mbarrier (double buffered)", "nt_"),
          ("column solver (colsolve.cu): 64-bit shared-memory atomics, double loads as LDS.128", "colsolve_"),
          ("frontier relaxation (sssp_frontier.cu): global 64-bit atomic min", "sf_"),
          ("refiner (refine.cu): on-device trial loop", "shortcut_loop")]
for title, key in groups:
    out.append("\n== " + title)
    for name, c in counts.items():
        if key in name:
            out.append("%-90s %s" % (name[:90], " ".join("%s=%d" % (k, c[k]) for k in PAT)))
path = os.path.join(ROOT, "profiles", "sass_%s.txt" % tag)
open(path, "w").write("\n".join(out) + "\n")
print("wrote", path)
