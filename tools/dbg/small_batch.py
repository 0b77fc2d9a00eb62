"""Per-kernel durations of small launches (pipeline forced on / off): where does the fixed cost sit?"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
from bench import make_rows, MODEL, ALLOWED
model = models.load(MODEL); eng = mj.ValidityEngine(model, ALLOWED)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
q = torch.from_numpy(make_rows(model, n)).cuda()
for _ in range(8): eng.valid_configs(q)
torch.cuda.synchronize()
