"""SASS instruction count of one kernel attributed to source functions (via -lineinfo).
usage: python tools/sass_by_function.py <lib.so> <cubin-name-substring> <kernel-substring> <source-header>"""
import bisect, collections, os, re, subprocess, sys, tempfile

lib, cubin_key, kern, src = sys.argv[1:5]
with tempfile.TemporaryDirectory(dir="build") as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
    cub = [f for f in os.listdir(d) if cubin_key in f][0]
    sass = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(d, cub)], capture_output=True, text=True).stdout
cur = None; fn = None; cnt = collections.Counter()
for line in sass.split("\n"):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r'\.text\.(\S+):', line)
    if m:
        fn = m.group(1); continue
    if re.match(r'\s+/\*[0-9a-f]{4,6}\*/', line) and fn and kern in fn:
        cnt[cur] += 1
starts = []
for i, l in enumerate(open(src).read().split("\n"), 1):
    m = re.match(r'W3D_HD\s+(?:inline\s+)?(?:constexpr\s+)?[\w:<>]+\s+(\w+)\(', l)
    if m: starts.append((i, m.group(1)))
byfn = collections.Counter()
for (f, ln), c in cnt.items():
    if f == os.path.basename(src):
        k = bisect.bisect_right([s for s, _ in starts], ln) - 1
        byfn[starts[k][1] if k >= 0 else "top"] += c
    else:
        byfn[f] += c
print("total", sum(cnt.values()))
for k, v in byfn.most_common(25): print(f"{v:7d} {k}")
