"""Per-kernel times of the bench workload for whichever library MJPL_B200_LIB points at (same-box A/B)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
from bench import make_rows, MODEL, ALLOWED
model = models.load(MODEL); eng = mj.get_engine(model, ALLOWED)
import os
NROWS = int(os.environ.get('AB_ROWS', '1000000'))
q = torch.from_numpy(make_rows(model, NROWS)).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(5): eng.valid_configs(q)
torch.cuda.synchronize()
KT = not os.environ.get('AB_NOTIMING')
if KT: eng.kernel_timing(True)
ts = []
for _ in range(40):
    flush.fill_(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); v = eng.valid_configs(q); e1.record(); ts.append((e0, e1))
torch.cuda.synchronize()
k = eng.kernel_timing(False, read=True) if KT else {'launches': 1, 'first_ms': 0, 'mid_ms': 0, 'narrow_ms': 0, 'fp64_ms': 0}
ms = sorted(a.elapsed_time(b) for a, b in ts)
n = k["launches"]
print(f"step median {ms[len(ms)//2]:.3f} ms min {ms[0]:.3f} | first {k['first_ms']/n:.3f} mid {k['mid_ms']/n:.3f} narrow {k['narrow_ms']/n:.3f} fp64 {k['fp64_ms']/n:.3f} | valid {v.float().mean().item():.5f}")
