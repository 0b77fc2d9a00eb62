import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
from bench import make_rows, MODEL, ALLOWED
model = models.load(MODEL); eng = mj.get_engine(model, ALLOWED)
q = torch.from_numpy(make_rows(model, 1_000_000)).pin_memory()
for _ in range(6): m = eng.valid_configs(q)
