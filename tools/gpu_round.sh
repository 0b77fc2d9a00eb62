#!/bin/bash
# Full GPU session: all gpu tests, smoke, bench (both arms), launch list + ncu captures, workload benches, sanitizers.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -45 > gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err
MJB_SPLIT=0 timeout 400 python bench.py --no-cpu-baseline > gpurun_out/bench_single_kernel.json 2>> gpurun_out/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 48 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:broad_kernel -s 3 -c 1 -o gpurun_out/prof_broad python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:narrow_kernel -s 3 -c 1 -o gpurun_out/prof_narrow python bench.py --steps 2 --warmup 3 --no-cpu-baseline >> gpurun_out/ncu_full.log 2>&1
timeout 900 python tools/bench_extra.py edges plans poses --queries 4096 2>&1 | grep '^{' > gpurun_out/bench_extra.jsonl
cut -c1-300 gpurun_out/bench_extra.jsonl
PYTHONPATH=. timeout 600 python examples/benchmark.py 2>&1 | tail -1 > gpurun_out/bench_config1.json
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
m = models.load("ur5e_scene")
s = mj.DLSIKSolver(m, mj.all_joints(m), iterations=50)
p = mj.site_pose(m, m.keyframe("home").qpos, "attachment_site")
print("ik", s.solve_rows(np.tile(p.translation(), (64, 1)), np.tile(p.rotation().wxyz, (64, 1)), np.zeros((64, 6)), "attachment_site")[1].mean())
PY
MJB_SPLIT=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_memcheck.log
MJB_SPLIT=1 timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san.py 2>&1 | tail -6 | tee gpurun_out/sanitizer_racecheck.log
