"""Attribute executed instructions / stall samples of one kernel to CUDA source lines, offline:
SASS rows of `ncu --page source --csv` are zipped (by position) with the line markers of `nvdisasm -g -c` on the
cubin extracted from the SAME libpvb.so the profile was taken with.
usage: python scripts/ncu_lines.py report.ncu-rep libpvb.so kernel_substring [top_n]"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, so, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin") and "bvh_build" not in f][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout.splitlines()
# locate the function section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
lines = []
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//---------------------"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True, stderr=subprocess.DEVNULL).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
print(f"sass rows in profile {len(data)}, instructions in disassembly {len(lines)}")
n = min(len(data), len(lines))
inst = defaultdict(float); samp = defaultdict(float); thr = defaultdict(float)
for r, ln in zip(data[:n], lines[:n]):
    inst[ln] += float(r[ix["Instructions Executed"]] or 0)
    thr[ln] += float(r[ix["Thread Instructions Executed"]] or 0)
    samp[ln] += float(r[ix["# Samples"]] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
src_cache = {}
def src(ln):
    f, no = ln
    for base in ("pytorch_volumetric_b200/csrc", "."):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), base, f)
        if os.path.exists(p):
            if p not in src_cache: src_cache[p] = open(p).read().splitlines()
            return src_cache[p][no - 1].strip()[:90] if 0 < no <= len(src_cache[p]) else ""
    return ""
print("inst%  samp%  lanes  file:line  source")
for ln, v in sorted(inst.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{v / ti * 100:5.2f}  {samp[ln] / ts * 100:5.2f}  {thr[ln] / max(v, 1):5.1f}  {ln[0]}:{ln[1]}  {src(ln)}")
