"""ncu source-page CSV (per SASS instruction) bucketed by source file and line range, using
nvdisasm --print-line-info output of the SAME build.
usage: python tools/ncu_samples_by_line.py <source.csv> <nvdisasm.sass> <kernel-mangled-substr> [bucket=20]"""
import collections, csv, re, sys
src_csv, sass, kern = sys.argv[1:4]
bucket = int(sys.argv[4]) if len(sys.argv) > 4 else 20
cur = None; infn = False; off2 = {}
for line in open(sass):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\.text\.(\S+):', line)
    if m: infn = kern in m.group(1); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
    if m and infn: off2[int(m.group(1), 16)] = cur
print('static instructions', len(off2))
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
base = int(rows[2][0], 16)
samp = collections.Counter(); inst = collections.Counter(); stat = collections.Counter()
st = collections.defaultdict(collections.Counter); miss = 0
KEYS = ('stall_long_sb', 'stall_no_inst', 'stall_wait', 'stall_short_sb', 'stall_math')
for r in rows[2:]:
    if len(r) < 10: continue
    off = int(r[0], 16) - base
    if off not in off2: miss += 1; continue
    f, l = off2[off] if off2[off] else ('?', 0)
    key = (f, (l // bucket) * bucket)
    s = int(r[ix['# Samples']]); samp[key] += s; inst[key] += int(r[ix['Instructions Executed']]); stat[key] += 1
    for k in KEYS: st[key][k] += int(r[ix[k]])
print('unmapped', miss)
T = sum(samp.values()); nw = int(rows[2][ix['Instructions Executed']])
for k, v in samp.most_common(30):
    print(f"{k[0]:26s} {k[1]:4d}  samples {100*v/T:5.1f}%  exec/warp {inst[k]/nw:6.0f}  static {stat[k]:5d}  "
          + " ".join(f"{s[6:]} {st[k][s]/max(v,1):.2f}" for s in KEYS))
