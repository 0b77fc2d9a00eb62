"""Raw validity call, throughput single kernel (MJB_ROWK_ROWS=0) vs the one-warp-per-row kernel, at small batch sizes."""
import os, sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models, _abi
from bench import make_rows, MODEL, ALLOWED
model = models.load(MODEL); rows = make_rows(model, 300_000); L = _abi.lib()
os.environ["MJB_SPLIT"] = "0"
for n in (1, 64, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
    res = {}
    for mode in ("0", "100000000"):
        os.environ["MJB_ROWK_ROWS"] = mode
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
        ms = sorted(a.elapsed_time(b) for a, b in ts); res[mode] = (ms[len(ms) // 2], out.cpu().numpy().copy()); eng.close()
    same = np.array_equal(res['0'][1], res['100000000'][1])
    print(f"rows {n:7d}: lane-per-row {res['0'][0]*1e3:8.1f} us   warp-per-row {res['100000000'][0]*1e3:8.1f} us   same mask {same}  valid {res['0'][1].mean():.4f}")
