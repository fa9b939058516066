#!/usr/bin/env python
"""Attribute local-memory (spill) instructions and FP64 instructions of one kernel to source lines.
usage: sass_spills_by_line.py <object-or-cubin> <mangled-kernel-prefix> [source-file-to-mark]
Needs -lineinfo at compile time; uses cuobjdump -xelf + nvdisasm -g (no GPU needed)."""
import collections, os, re, subprocess, sys, tempfile

obj, prefix = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
if obj.endswith(".cubin"):
    cub = os.path.abspath(obj)
else:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cub = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
txt = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(txt) if l.startswith(".text." + prefix)][0]
end = [i for i, l in enumerate(txt) if l.startswith(".text.") and i > start]
end = end[0] if end else len(txt)
cur = None
spill, allc, fp, ops = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
for l in txt[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", l)
    if m:
        op = m.group(1)
        allc[cur] += 1
        ops[op] += 1
        if op in ("STL", "LDL"):
            spill[cur] += 1
        if op in ("DFMA", "DMUL", "DADD"):
            fp[cur] += 1
print("instructions", sum(allc.values()), "fp64", sum(fp.values()), "local ld/st", sum(spill.values()))
print("top opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(14)))
print("local-memory instructions by source line:")
for k, v in sorted(spill.items(), key=lambda kv: -kv[1])[:25]:
    print("  ", k, v)
