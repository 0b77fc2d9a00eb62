import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
from bench import make_rows, MODEL, ALLOWED
model = models.load(MODEL); eng = mj.get_engine(model, ALLOWED)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
q = torch.from_numpy(make_rows(model, n)).pin_memory()
for _ in range(3): m = eng.valid_configs(q)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): m = eng.valid_configs(q)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print(f"rows {n}: e2e {dt*1e3:.3f} ms  {n/dt:.3e} cfg/s")
