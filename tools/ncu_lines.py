"""Attribute an ncu --set full capture of validity_kernel to source lines.

usage: python tools/ncu_lines.py gpurun_out/prof.ncu-rep [topN]
Joins `ncu --page source --csv` (per-SASS-instruction counters) with `nvdisasm -gi` of the
library's cubin (inline-aware line info) by instruction order.
"""
import collections, csv, re, subprocess, sys, tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, val = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want or ("issue_stalled" in h and "per_warp_active" in h and float(val[i] or 0) > 3):
        print(f"{h:95s} {val[i]:>16s} {units[i]}")
kname = val[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""
m = re.search(r"validity_kernel<\(?i?n?t?\)?(\d+)>", kname)
tile = m.group(1) if m else "512"
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(src.splitlines()))
shdr, data = srows[1], srows[2:]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(ROOT / "mjpl_b200/lib/libmjpl_b200.so")], cwd=td, capture_output=True)
    cubin = next(Path(td).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True).stdout.split("\n")
start = [i for i, l in enumerate(dis) if l.startswith(f".text._ZN2vk15validity_kernelILi{tile}")][0]
end = [i for i, l in enumerate(dis[start + 1:], start + 1) if l.startswith(".text.") or l.startswith(".section")][0]
insts, block, prev = [], [], False
for l in dis[start:end]:
    mm = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if mm:
        if not prev:
            block = []
        block.append((mm.group(1).split("/")[-1], int(mm.group(2)), (mm.group(3) or "").split("/")[-1], int(mm.group(4) or 0)))
        prev = True
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        inner = (block[0][0], block[0][1]) if block else ("?", 0)
        last = block[-1] if block else None
        outer = (last[2], last[3]) if last and last[2] else ((last[0], last[1]) if last else ("?", 0))
        insts.append((inner, outer))
    prev = False
if len(insts) != len(data):
    print(f"WARNING: {len(insts)} disassembled instructions vs {len(data)} profiled (library rebuilt since the capture?)")
ia, it, ism = shdr.index("Instructions Executed"), shdr.index("Thread Instructions Executed"), shdr.index("# Samples")
tot = sum(int(r[ia]) for r in data); tots = sum(int(r[ism]) for r in data); totthr = sum(int(r[it]) for r in data)
print(f"total warp-inst {tot}  avg active threads {totthr / tot:.2f}  samples {tots}")
files = {"vk_kernels.cuh": (ROOT / "mjpl_b200/csrc/vk_kernels.cuh").read_text().split("\n"),
         "vk_core.cuh": (ROOT / "mjpl_b200/csrc/vk_core.cuh").read_text().split("\n")}
for title, idx in (("outermost (kernel) line", 1), ("innermost line", 0)):
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for r, ii in zip(data, insts):
        k = ii[idx]
        agg[k][0] += int(r[ia]); agg[k][1] += int(r[it]); agg[k][2] += int(r[ism])
    print(f"== by {title} ==")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        txt = files.get(k[0], [""])[k[1] - 1].strip()[:64] if k[0] in files and 0 < k[1] <= len(files[k[0]]) else ""
        print(f"{k[0]}:{k[1]:4d} inst {v[0] / tot * 100:5.1f}% thr {v[1] / max(v[0], 1):5.1f} samples {v[2] / tots * 100:5.1f}%  {txt}")
