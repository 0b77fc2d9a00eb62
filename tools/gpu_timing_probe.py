"""Why does the per-step event time differ from ncu's kernel time? Probe flush / sync variants."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
from bench import make_rows, MODEL, ALLOWED

model = models.load(MODEL); eng = mj.get_engine(model, ALLOWED)
q = torch.from_numpy(make_rows(model, 1_000_000)).cuda()
out = torch.empty(len(q), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
import ctypes as C
from mjpl_b200 import _abi
L = _abi.lib()
def raw():
    _abi.check(L.mjb_check_configs(eng._h, q.data_ptr(), len(q), 9, out.data_ptr(), 3, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
for _ in range(5): raw()
torch.cuda.synchronize()
def timed(label, pre=None, sync=False, n=10, fn=raw):
    ts = []
    for _ in range(n):
        if pre: pre()
        if sync: torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        ts.append((e0, e1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ts]
    print(f"{label:40s} mean {np.mean(ms):.3f} ms  min {min(ms):.3f} max {max(ms):.3f}")
timed("raw C call, no flush")
timed("raw C call, no flush, sync before", sync=True)
timed("raw C call, flush", pre=lambda: flush.fill_(1))
timed("raw C call, flush + sync", pre=lambda: flush.fill_(1), sync=True)
timed("raw C call, flush zero_", pre=lambda: flush.zero_())
timed("engine.valid_configs, no flush", fn=lambda: eng.valid_configs(q, 3))
timed("engine.valid_configs, flush", pre=lambda: flush.fill_(1), fn=lambda: eng.valid_configs(q, 3))
small = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
timed("raw C call, flush 64MB", pre=lambda: small.fill_(1))
