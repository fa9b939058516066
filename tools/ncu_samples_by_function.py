"""Attributes an ncu source-page CSV (per SASS instruction: samples, instructions executed, stall
reasons) to source functions, using nvdisasm line info of the SAME build of the kernel.
usage: python tools/ncu_samples_by_function.py <source.csv> <nvdisasm --print-line-info output> <kernel-substr> <header>"""
import bisect, collections, csv, re, sys

src_csv, sass_file, kern, header = sys.argv[1:5]
cur = None; infn = False; off2line = {}
for line in open(sass_file):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\.text\.(\S+):', line)
    if m: infn = kern in m.group(1); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);', line)
    if m and infn: off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {n: i for i, n in enumerate(hdr)}
base = int(rows[2][0], 16)
starts = []
for i, l in enumerate(open(header).read().split('\n'), 1):
    m = re.match(r'W3D_HD\s+(?:inline\s+)?(?:constexpr\s+)?[\w:<>]+\s+(\w+)\(', l)
    if m: starts.append((i, m.group(1)))
S = [s for s, _ in starts]
samp = collections.Counter(); inst = collections.Counter(); mism = 0
stall = collections.defaultdict(collections.Counter)
KEYS = ('stall_long_sb', 'stall_no_inst', 'stall_wait')
for r in rows[2:]:
    if len(r) < 10: continue
    off = int(r[0], 16) - base
    if off not in off2line: mism += 1; continue
    loc = off2line[off]
    if loc is None: fn = '?'
    elif loc[0] == header.split('/')[-1]:
        k = bisect.bisect_right(S, loc[1]) - 1; fn = starts[k][1] if k >= 0 else 'top'
    else: fn = loc[0]
    samp[fn] += int(r[ix['# Samples']]); inst[fn] += int(r[ix['Instructions Executed']])
    for st in KEYS: stall[fn][st] += int(r[ix[st]])
nwarp = int(rows[2][ix['Instructions Executed']])
print('unmapped', mism, 'warps', nwarp)
T = sum(samp.values()); TI = sum(inst.values())
for k, v in samp.most_common(24):
    print(f"{k:26s} samples {100*v/T:5.1f}%  inst {100*inst[k]/TI:5.1f}% ({inst[k]/nwarp:7.0f}/warp)  "
          + "  ".join(f"{s[6:]} {stall[k][s]/max(v,1):.2f}" for s in KEYS))
