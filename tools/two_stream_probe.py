"""Does running two halves of a batch on two handles / two streams fill the kernels' tails?  (experiment)"""
import sys, os
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
from mjpl_b200.engine import ValidityEngine
from bench import make_rows, MODEL, ALLOWED
model = models.load(MODEL)
N = int(os.environ.get("AB_ROWS", "1000000"))
q = torch.from_numpy(make_rows(model, N)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for parts in (1, 2, 3, 4):
    engs = [ValidityEngine(model, ALLOWED) for _ in range(parts)]
    streams = [torch.cuda.Stream() for _ in range(parts)]
    cuts = [N * i // parts for i in range(parts + 1)]
    def step():
        cur = torch.cuda.current_stream()
        outs = []
        for e, s, a, b in zip(engs, streams, cuts[:-1], cuts[1:]):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                outs.append(e.valid_configs(q[a:b]))
        for s in streams: cur.wait_stream(s)
        return outs
    for _ in range(5): step()
    torch.cuda.synchronize()
    ts = []
    for _ in range(30):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); o = step(); e1.record(); ts.append((e0, e1))
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ts)
    print(f"parts {parts}: step median {ms[len(ms)//2]:.3f} ms min {ms[0]:.3f}  valid {torch.cat(o).float().mean().item():.5f}")
