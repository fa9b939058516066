#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of a built library (cuobjdump -sass): static instruction
counts per mnemonic for every kernel whose (mangled) name contains one of the given substrings.
usage: sass_opcode_histogram.py warp3d_b200/libw3d_engine.so k_tanstf k_eps_sig_ifvILi1 k_mm10f ..."""
import collections
import re
import subprocess
import sys


def main():
    lib, keys = sys.argv[1], sys.argv[2:]
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    name, hist = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", line)
        if m and name and any(k in name for k in keys):
            toks = m.group(1).split()
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            hist.setdefault(name, collections.Counter())[op.split(".")[0]] += 1
    for name in sorted(hist):
        c = hist[name]
        tot = sum(c.values())
        demangled = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        print(f"{demangled}\n  {tot} instructions ({tot * 16 / 1024:.1f} KB): " +
              ", ".join(f"{op} {n}" for op, n in c.most_common(16)))


if __name__ == "__main__":
    main()
