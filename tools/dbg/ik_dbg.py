"""device vs host IK: one iteration at a time from identical inputs (debug aid)"""
import sys
sys.path.insert(0, ".")
import numpy as np
import oracle
import mjpl_b200 as mj
from mjpl_b200 import models
from tests.hostsim import HostSim

np.set_printoptions(precision=6, linewidth=200)
for mname, site in (("ur5e_scene", "attachment_site"),):
    model = models.load(mname)
    po = oracle.PoseOracle(model, site, [0, 0, 0], [1, 0, 0, 0], [(-np.inf, np.inf)] * 6)
    rng = np.random.default_rng(11)
    lo, hi = model.jnt_range.T
    n = 256
    poses = [po.site_pose(q) for q in rng.uniform(lo, hi, size=(n, model.nq))]
    tp = np.array([p for p, _ in poses]); tq = np.array([r for _, r in poses])
    q = rng.uniform(lo, hi, size=(n, model.nq))
    s = mj.DLSIKSolver(model, mj.all_joints(model), iterations=1)
    hs = HostSim(model)
    for step in range(4):
        Q, ok, it, er = s.solve_rows(tp, tq, q, site)
        Qh, okh, ith, erh = hs.ik(s._spec(site), tp, tq, q)
        d = np.abs(Q - Qh).max(axis=1)
        bad = np.flatnonzero(d > 1e-9)
        print("step", step, "rows differing", len(bad), "max", d.max())
        for i in bad[:6]:
            atl = (q[i] <= lo) | (q[i] >= hi)
            print(" row", i, "in   ", q[i], "at limit", atl.astype(int))
            print("        dev  ", Q[i] - q[i])
            print("        host ", Qh[i] - q[i])
        q = Qh
