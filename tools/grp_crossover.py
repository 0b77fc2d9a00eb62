"""Single validity kernel, one lane per item vs GRP_SMALL lanes per item (MJB_GRP_ROWS), raw call at several batch sizes."""
import os, sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models, _abi
from bench import make_rows, MODEL, ALLOWED
model = models.load(MODEL); rows = make_rows(model, 300_000); L = _abi.lib()
os.environ["MJB_SPLIT"] = "0"
for n in (1024, 4096, 8192, 16384, 32768, 65536, 131072, 262144):
    res = {}
    for mode in ("0", "100000000"):
        os.environ["MJB_GRP_ROWS"] = mode
        eng = mj.ValidityEngine(model, ALLOWED)
        q = torch.from_numpy(rows[:n]).cuda(); out = torch.empty(n, dtype=torch.uint8, device="cuda")
        def raw():
            _abi.check(L.mjb_check_configs(eng._h, q.data_ptr(), n, 9, out.data_ptr(), 3, C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        for _ in range(10): raw()
        torch.cuda.synchronize(); ts = []
        for _ in range(40):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); raw(); e1.record(); ts.append((e0, e1))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ts); res[mode] = (ms[len(ms) // 2], out.float().mean().item()); eng.close()
    print(f"rows {n:7d}: 1 lane/item {res['0'][0]*1e3:8.1f} us   8 lanes/item {res['100000000'][0]*1e3:8.1f} us   valid {res['0'][1]:.4f} {res['100000000'][1]:.4f}")
