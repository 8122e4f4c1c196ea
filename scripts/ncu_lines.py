"""Per-source-line view of an ncu capture: joins the SASS rows of `ncu --page source --csv` (instruction counts,
stall samples) with the line table of the same kernel from `nvdisasm --print-line-info`.
Usage: python scripts/ncu_lines.py <x.ncu-rep> <kernel mangled-name substring> [top N]
The library must be the build the capture was taken with."""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("VO_SO", os.path.join(root, "voroffset_b200", "libvoroffset_b200.so"))     # the build the capture was taken with
srcdir = os.environ.get("VO_SRC", os.path.join(root, "voroffset_b200", "csrc"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout
# instruction -> (file, line) for the wanted function
lines, cur, on = [], ("?", 0), False
for l in dis.splitlines():
    if l.startswith("//--------------------- .text."):
        on = pat in l
        continue
    if not on: continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append((cur, l.split("*/", 1)[1].strip()))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
# one section per profiled launch ("Kernel Name" row, header row, instructions); NCU_KERNEL = substring of the
# demangled name picks the launch (default: the first section)
want = os.environ.get("NCU_KERNEL", "")
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
pick = next((i for i in starts if want in rows[i][1]), starts[0])
if os.environ.get("NCU_INDEX"):          # ... or the n-th profiled launch of the capture (template arguments are not in the name)
    pick = starts[int(os.environ["NCU_INDEX"])]
nxt = next((i for i in starts if i > pick), len(rows))
rows = rows[pick:nxt]
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
if len(body) != len(lines):
    print(f"warning: {len(body)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = [0, 0, 0]
for (loc, _), r in zip(lines, body):
    v = [int(r[ci["# Samples"]] or 0), int(r[ci["Instructions Executed"]] or 0), int(r[ci["Thread Instructions Executed"]] or 0)]
    for k in range(3):
        agg[loc][k] += v[k]; tot[k] += v[k]
src = {}
print(f"total: samples {tot[0]}  warp-inst {tot[1]}  thread-inst {tot[2]}")
print(f"{'file:line':28s} {'samples%':>8s} {'inst%':>7s} {'thr/warp':>8s}  source")
for loc, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f = os.path.join(srcdir, loc[0])
    if loc[0] not in src:
        try: src[loc[0]] = open(f).read().splitlines()
        except Exception: src[loc[0]] = []
    text = src[loc[0]][loc[1] - 1].strip()[:110] if 0 < loc[1] <= len(src[loc[0]]) else ""
    print(f"{loc[0] + ':' + str(loc[1]):28s} {100 * v[0] / max(tot[0], 1):8.2f} {100 * v[1] / max(tot[1], 1):7.2f} {v[2] / max(v[1], 1):8.1f}  {text}")
