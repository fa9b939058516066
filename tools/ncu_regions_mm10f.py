#!/usr/bin/env python
"""Executed instructions / FP64 instructions / stall samples of k_mm10f per region of
w3d_mm10_fast.h, from an ncu source-page CSV and nvdisasm -g of the SAME build.
usage: ncu_regions_mm10f.py <source.csv> <nvdisasm -g output> [kernel substring]"""
import collections, csv, re, sys
src_csv, sass = sys.argv[1:3]
kern = sys.argv[3] if len(sys.argv) > 3 else 'k_mm10fILb0ELb1ELi19'
cur = None; infn = False; off2 = {}
for line in open(sass):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\.text\.(\S+):', line)
    if m: infn = kern in m.group(1); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
    if m and infn: off2[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
base = int(rows[2][0], 16)
import os
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'warp3d_b200', 'csrc', 'w3d_mm10_fast.h')).read().split('\n')
def find(s): return [i + 1 for i, l in enumerate(src) if s in l][0]
marks = [('helpers', 1), ('setup', find('// ---- mm10_setup: B6 = RE E6')), ('strup_init', find('// ---- mm10_solve_strup')),
         ('looptop', find('if (!SYNC::any(!done)) break;')), ('begin_sub', find('if (ct.begin_sub) {')),
         ('residual', find('// ================= residual at xt')), ('stage logic', find('if (ct.stage == ST_ELASTIC) {')),
         ('jacobian', find('// ================= Jacobian at x')), ('lu+solve', find('const int jstop = lu7_nopivot')),
         ('after loop', find('if (idle) return false;')), ('output', find('// ---- mm10_output (mm10_a.f:3493)')),
         ('rotation', find('// ---- mm10_update_rotation')), ('gauss_point', find('W3DF_HD inline bool gauss_point_update_fast')),
         ('host', find('// ---- host-side set-up'))]
inst = collections.Counter(); samp = collections.Counter(); fp = collections.Counter()
stl = collections.defaultdict(collections.Counter)
KEYS = ['stall_barrier', 'stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_math', 'stall_no_inst', 'stall_mio', 'stall_lg']
nw = int(rows[2][ix['Instructions Executed']])
for r in rows[2:]:
    if len(r) < 10: continue
    off = int(r[0], 16) - base
    if off not in off2: continue
    (cur, op) = off2[off]
    f, l = cur if cur else ('?', 0)
    if f == 'w3d_mm10_fast.h':
        reg = [n for n, lo in marks if lo <= l][-1]
        if reg == 'helpers': reg = 'h:' + str(l)
    else: reg = f
    n = int(r[ix['Instructions Executed']]); s = int(r[ix['# Samples']])
    inst[reg] += n; samp[reg] += s
    if re.match(r'(@!?P\d+\s+)?(DFMA|DMUL|DADD)', op): fp[reg] += n
    for k in KEYS: stl[reg][k] += int(r[ix[k]])
T = sum(samp.values()); TI = sum(inst.values())
print("warps", nw, "instr/warp %.0f" % (TI / nw), "fp64/warp %.0f" % (sum(fp.values()) / nw))
for k, v in inst.most_common(26):
    print(f"{k:14s} exec/warp {v/nw:8.0f} fp64/warp {fp[k]/nw:7.0f} samples {100*samp[k]/T:5.1f}%  " +
          " ".join(f"{s[6:]} {stl[k][s]/max(samp[k],1):.2f}" for s in KEYS))
