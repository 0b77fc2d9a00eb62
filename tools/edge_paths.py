"""Edges: pipeline vs single kernel per model (which path should large edge batches take?)"""
import os, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
from mjpl_b200.engine import ValidityEngine
NE = int(os.environ.get("AB_EDGES", "100000"))
for name, allowed in (("ur5e_scene", []), ("franka_scene_with_obstacles", [("left_finger", "right_finger")])):
    model = models.load(name)
    rng = np.random.default_rng(0)
    lo, hi = model.jnt_range[:, 0], model.jnt_range[:, 1]
    q0 = torch.from_numpy(rng.uniform(lo, hi, size=(NE, model.nq)).astype(np.float32)).cuda()
    q1 = torch.from_numpy(rng.uniform(lo, hi, size=(NE, model.nq)).astype(np.float32)).cuda()
    res = {}
    for split in ("auto", "0"):
        if split == "auto": os.environ.pop("MJB_SPLIT", None)
        else: os.environ["MJB_SPLIT"] = split
        eng = ValidityEngine(model, allowed)
        for _ in range(3): v, fb = eng.valid_edges(q0, q1, 0.05, want_first_bad=True)
        torch.cuda.synchronize(); eng.reset_stats(); ts = []
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); v, fb = eng.valid_edges(q0, q1, 0.05, want_first_bad=True); e1.record(); ts.append((e0, e1))
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ts)
        res[split] = (v.cpu(), fb.cpu())
        print(f"{name:30s} split {split:5s}: {ms[len(ms)//2]:.3f} ms per {NE} edges, {eng.stats()['rows']//6} waypoints, valid {v.float().mean().item():.4f}")
        eng.close()
    print("   same answers:", bool(torch.equal(res['auto'][0], res['0'][0]) and torch.equal(res['auto'][1], res['0'][1])))
