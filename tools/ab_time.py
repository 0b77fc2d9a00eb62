"""Time the raw validity call of whichever library MJPL_B200_LIB points at (median of many)."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models, _abi
from bench import make_rows, MODEL, ALLOWED
model = models.load(MODEL); eng = mj.get_engine(model, ALLOWED)
q = torch.from_numpy(make_rows(model, 1_000_000)).cuda()
out = torch.empty(len(q), dtype=torch.uint8, device="cuda")
L = _abi.lib()
FLAGS = int(sys.argv[1]) if len(sys.argv) > 1 else 3   # 11 = skip the fp64 re-evaluation kernel
def raw():
    _abi.check(L.mjb_check_configs(eng._h, q.data_ptr(), len(q), 9, out.data_ptr(), FLAGS, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
for _ in range(10): raw()
torch.cuda.synchronize()
ts = []
for _ in range(60):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); raw(); e1.record(); ts.append((e0, e1))
torch.cuda.synchronize()
ms = sorted(a.elapsed_time(b) for a, b in ts)
print(f"median {ms[len(ms)//2]:.3f} ms  min {ms[0]:.3f}  valid {out.float().mean().item():.4f}")
