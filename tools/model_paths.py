"""Pipeline vs single kernel per model (dense rows): which path should large batches of a model take?"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
from mjpl_b200.engine import ValidityEngine
N = int(os.environ.get("AB_ROWS", "1000000"))
for name, allowed in (("ur5e_scene", []), ("franka_scene", []), ("franka_scene_with_obstacles", [("left_finger", "right_finger")])):
    model = models.load(name)
    rng = np.random.default_rng(0)
    lo, hi = model.jnt_range[:, 0], model.jnt_range[:, 1]
    q = torch.from_numpy(rng.uniform(lo, hi, size=(N, model.nq)).astype(np.float32)).cuda()
    for split in ("auto", "0"):
        if split == "auto": os.environ.pop("MJB_SPLIT", None)
        else: os.environ["MJB_SPLIT"] = split
        eng = ValidityEngine(model, allowed)
        for _ in range(3): v = eng.valid_configs(q)
        torch.cuda.synchronize(); ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); v = eng.valid_configs(q); e1.record(); ts.append((e0, e1))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ts)
        print(f"{name:30s} split {split:5s}: {ms[len(ms)//2]:.3f} ms per {N} rows  valid {v.float().mean().item():.4f}")
        eng.close()
