#!/bin/bash
# Full GPU session: all gpu tests, smoke, bench (both arms), sanitizer on a small case.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -16 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import mjpl_b200 as mj
from mjpl_b200 import models
for name, al in (("franka_scene_with_obstacles", [("left_finger", "right_finger")]), ("ur5e_scene", [])):
    m = models.load(name); e = mj.get_engine(m, al)
    rng = np.random.default_rng(0)
    Q = rng.uniform(m.jnt_range[:, 0], m.jnt_range[:, 1], size=(3000, m.nq)).astype(np.float32)
    print(name, e.valid_configs(Q).mean(), e.valid_edges(Q[:200], Q[200:400], 0.05).mean(), e.sweep(1, 0, 2000).float().mean().item())
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py 2>&1 | tail -5 | tee gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san.py 2>&1 | tail -5 | tee gpurun_out/sanitizer_racecheck.log
