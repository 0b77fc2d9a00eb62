import sys; sys.path.insert(0,'.')
import numpy as np, torch, mjpl_b200 as mj
m = mj.models.load("franka_scene_with_obstacles")
eng = mj.get_engine(m, [("left_finger","right_finger")])
rng=np.random.default_rng(0)
Q=torch.from_numpy(rng.uniform(m.jnt_range[:,0],m.jnt_range[:,1],size=(1000000,m.nq)).astype(np.float32)).cuda()
eng.kernel_timing(True)
for _ in range(3): v=eng.valid_configs(Q)
print(eng.kernel_timing(False, read=True), float(v.float().mean()), eng.stats())
