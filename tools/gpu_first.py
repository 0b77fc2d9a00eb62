"""First GPU shake-down: parity vs oracle on all bundled models + rough throughput."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import mjpl_b200 as mj, oracle

def run(name, allowed, n=200000):
    model = mj.models.load(name)
    eng = mj.get_engine(model, allowed)
    orc = oracle.Oracle(model, allowed); oracle.Oracle.set_threads(16)
    rng = np.random.default_rng(0)
    Q = rng.uniform(model.jnt_range[:, 0], model.jnt_range[:, 1], size=(n, model.nq)).astype(np.float32)
    Qd = torch.from_numpy(Q).cuda()
    for flags in (3, 3 | 4):
        got = eng.valid_configs(Qd, flags).cpu().numpy()
        want, dist, pair = orc.check(Q.astype(np.float64), 3, want_dist=True)
        bad = np.flatnonzero(got != want)
        print(f"{name} flags={flags}: rows={n} valid={got.mean():.4f} mismatches={len(bad)} outside-band={(np.abs(dist[bad])>=1e-5).sum()} stats={eng.stats()}")
    xp, xq = eng.fk(Qd[:20000]); op, oq = orc.fk(Q[:20000].astype(np.float64))
    xq = xq.cpu().numpy(); dq = np.minimum(np.abs(xq-oq).max(-1), np.abs(xq+oq).max(-1))
    print(f"  fk pos err {np.abs(xp.cpu().numpy()-op).max():.2e} quat err {dq.max():.2e}")
    # timing
    big = Qd.repeat((max(1, 1000000 // n), 1)).contiguous()
    out = eng.valid_configs(big); torch.cuda.synchronize()
    for tile in ("",):
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(5): eng.valid_configs(big)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / 5
        print(f"  {len(big)} rows: {ms:.3f} ms -> {len(big)/ms*1e3:.3e} configs/s")

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    run("two_dof_ball", [], 2000 if quick else 20000)
    if quick:
        run("franka_scene", [], 3000)
        sys.exit(0)
    run("ur5e_scene", [])
    run("franka_scene", [])
    run("franka_scene_with_obstacles", [("left_finger", "right_finger")])
    import __graft_entry__ as g
    g.smoke()
