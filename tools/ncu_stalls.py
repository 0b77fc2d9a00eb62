"""Per source line: which stall reasons the samples of a kernel fall on (ncu --page source of a -lineinfo build)."""
import collections, csv, re, subprocess, sys, tempfile
from pathlib import Path
rep, sym = sys.argv[1], sys.argv[2]
lib = sys.argv[4] if len(sys.argv) > 4 else "/root/repo/mjpl_b200/lib/libmjpl_b200.so"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
shdr, data = srows[1], srows[2:]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
    cubin = next(Path(td).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(".text." + sym)][0]
end = [i for i, l in enumerate(dis[start + 1:], start + 1) if l.startswith(".text.") or l.startswith(".section")][0]
insts, block, prev = [], [], False
for l in dis[start:end]:
    mm = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if mm:
        if not prev: block = []
        block.append((mm.group(1).split("/")[-1], int(mm.group(2)))); prev = True; continue
    m2 = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", l)
    if m2:
        insts.append(((block[0] if block else ("?", 0)), m2.group(1)))
    prev = False
assert len(insts) == len(data), (len(insts), len(data))
st = [i for i, n in enumerate(shdr) if n.startswith("stall_") and "Not Issued" not in n]
ism = shdr.index("# Samples")
tots = sum(int(r[ism]) for r in data)
rows = []
for r, (loc, txt) in zip(data, insts):
    d = {shdr[i][6:]: int(r[i] or 0) for i in st if int(r[i] or 0)}
    rows.append((int(r[ism]), loc, txt, d))
rows.sort(key=lambda x: -x[0])
for n, loc, txt, d in rows[:top]:
    dd = ", ".join(f"{k} {v}" for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:3])
    print(f"{n/tots*100:5.1f}%  {loc[0]}:{loc[1]:<5d} {txt[:60]:60s} {dd}")
