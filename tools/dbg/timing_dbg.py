"""exercise mjb_kernel_timing + close (debug aid for a crash in mjb_model_destroy)"""
import sys, faulthandler
sys.path.insert(0, ".")
faulthandler.enable()
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
step = sys.argv[1] if len(sys.argv) > 1 else "all"
m = models.load("franka_scene_with_obstacles")
e = mj.ValidityEngine(m, [("left_finger", "right_finger")])
Q = torch.from_numpy(np.random.default_rng(0).uniform(m.jnt_range[:, 0], m.jnt_range[:, 1], size=(200000, m.nq)).astype(np.float32)).cuda()
if step in ("all", "timing", "timing_noread", "all_numpy", "timing_small_host"):
    e.kernel_timing(True)
for _ in range(3):
    e.valid_configs(Q)
torch.cuda.synchronize()
if step in ("all", "timing"):
    print(e.kernel_timing(False, read=True))
if step in ("all", "host"):
    print(e.valid_configs(Q.cpu().pin_memory()).float().mean())
if step in ("all_numpy",):
    print(e.kernel_timing(False, read=True))
    print(e.valid_configs(Q.cpu().numpy()).mean())
if step in ("timing_small_host",):
    print(e.kernel_timing(False, read=True))
    print(e.valid_configs(Q[:1000].cpu().numpy()).mean())
print("closing"); sys.stdout.flush()
e.close()
print("closed ok", step)
