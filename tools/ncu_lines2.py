import collections, csv, re, subprocess, sys, tempfile
from pathlib import Path
rep, sym, lib = sys.argv[1], sys.argv[2], "/root/repo/mjpl_b200/lib/libmjpl_b200.so"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
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
        block.append((mm.group(1).split("/")[-1], int(mm.group(2)), (mm.group(3) or "").split("/")[-1], int(mm.group(4) or 0)))
        prev = True; continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        inner = (block[0][0], block[0][1]) if block else ("?", 0)
        last = block[-1] if block else None
        outer = (last[2], last[3]) if last and last[2] else ((last[0], last[1]) if last else ("?", 0))
        insts.append((inner, outer))
    prev = False
print(len(insts), len(data))
ia, it, ism = shdr.index("Instructions Executed"), shdr.index("Thread Instructions Executed"), shdr.index("# Samples")
ino, iw = shdr.index("stall_no_inst"), shdr.index("stall_wait")
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[ism]) for r in data); totthr = sum(int(r[it]) for r in data)
print(f"total warp-inst {tot} avg thr {totthr/tot:.2f} samples {tots} no_inst {sum(int(r[ino]) for r in data)/tots*100:.1f}% wait {sum(int(r[iw]) for r in data)/tots*100:.1f}%")
files = {n: (Path("/root/repo/mjpl_b200/csrc")/n).read_text().split("\n") for n in ("vk_kernels.cuh","vk_core.cuh","vk_split.cuh","vk_pipe.cuh")}
for title, idx in (("outermost", 1), ("innermost", 0)):
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    for r, ii in zip(data, insts):
        k = ii[idx]; agg[k][0] += int(r[ia]); agg[k][1] += int(r[it]); agg[k][2] += int(r[ism]); agg[k][3] += 1
    print("== by", title)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        txt = files.get(k[0], [""])[k[1]-1].strip()[:70] if k[0] in files and 0 < k[1] <= len(files[k[0]]) else ""
        print(f"{k[0]}:{k[1]:4d} sass {v[3]:4d} inst {v[0]/tot*100:5.1f}% thr {v[1]/max(v[0],1):5.1f} samples {v[2]/tots*100:5.1f}%  {txt}")
