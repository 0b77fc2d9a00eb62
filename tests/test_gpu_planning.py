"""Planners on the real CUDA constraints; every returned path must replay as valid under the
fp64 oracle (the "reference checker" stand-in)."""

import numpy as np
import pytest

import oracle
import mjpl_b200 as mj
from mjpl_b200 import models
from mjpl_b200.planning.tree import Node, Tree
from mjpl_b200.planning.utils import _constrained_extend, _valid_collision_interval

pytestmark = pytest.mark.gpu


def replay_valid(model, allowed, path, eps):
    orc = oracle.Oracle(model, allowed)
    P = np.array(path)
    assert orc.check(P, 3).all()
    assert (np.linalg.norm(np.diff(P, axis=0), axis=1) <= eps + 1e-9).all()


def test_reference_planning_utils_cases():
    # reference test/test_planning_utils.py:207-344 on the CUDA constraints
    model = models.load("one_dof_ball")
    c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model)]
    tree = Tree(Node(np.array([-0.1])))
    q_goal = np.array([0.15])
    np.testing.assert_equal(_constrained_extend(q_goal, tree, 0.1, c), q_goal)
    path = [n.q for n in tree.get_path(tree.nearest_neighbor(q_goal))]
    for p, e in zip(path, [[0.15], [0.1], [0.0], [-0.1]]):
        np.testing.assert_allclose(p, e, atol=1e-9)
    tree = Tree(Node(np.array([0.0])))
    r = _constrained_extend(np.array([1.0]), tree, 0.1, c)
    assert 0.0 < r[0] < 0.85
    tree = Tree(Node(np.array([0.8])))
    np.testing.assert_equal(_constrained_extend(np.array([1.8]), tree, np.inf, c, (0.3, c[1])), [1.8])
    tree = Tree(Node(np.array([0.8])))
    np.testing.assert_equal(_constrained_extend(np.array([1.8]), tree, np.inf, c, (0.1, c[1])), [0.8])
    assert not _valid_collision_interval(np.array([0.8]), np.array([1.5]), 0.1, c[1])
    assert _valid_collision_interval(np.array([0.8]), np.array([1.5]), 0.2, c[1])
    assert _valid_collision_interval(np.array([0.0]), np.array([0.2]), 0.01, c[1])
    with pytest.raises(ValueError, match="step_dist"):
        _valid_collision_interval(np.array([0.0]), np.array([0.2]), 0.0, c[1])


def test_smooth_path_two_dof():
    model = models.load("two_dof_ball")
    c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model)]
    wps = [np.array(p) for p in [[0.0, 0.0], [0.25, 0.0], [0.25, 1.5], [0.5, 1.5], [1.0, 1.5], [1.0, 0.0], [1.0, 0.0]]]
    out = mj.smooth_path(wps, c, eps=0.1, seed=5)
    assert mj.path_length(out) < mj.path_length(wps) and len(out) > 2
    replay_valid(model, [], out, 0.1)


def test_rrt_franka_benchmark_query():
    """BASELINE config 1 (examples/benchmark.py:28-48), goal given as the configuration
    random_config(seed=42) produces (the reference converts it to a pose and runs IK)."""
    model = models.load("franka_scene")
    joints = [f"joint{i}" for i in range(1, 8)]
    c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model)]
    q_init = model.keyframe("home").qpos.copy()
    q_goal = mj.random_config(model, q_init, joints, 42, c)
    assert mj.obeys_constraints(q_goal, c)
    np.testing.assert_equal(q_goal[7:], q_init[7:])
    planner = mj.RRT(model, joints, c, max_planning_time=10, epsilon=0.05, seed=42, goal_biasing_probability=0.1)
    path = planner.plan_to_config(q_init, q_goal)
    assert path, "planner timed out"
    np.testing.assert_equal(path[0], q_init)
    np.testing.assert_equal(path[-1], q_goal)
    replay_valid(model, [], path, 0.05)
    short = mj.smooth_path(path, c, eps=0.05, seed=42)
    assert mj.path_length(short) <= mj.path_length(path)
    replay_valid(model, [], short, 0.05)


def test_smooth_path_ur5e():
    # reference test_planning_utils.py:154-182 on the local ur5e.xml
    model = models.load("ur5e_scene")
    c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model)]
    wps = [model.keyframe("home").qpos.copy()]
    for i in range(5):
        wps.append(mj.random_config(model, np.zeros(model.nq), mj.all_joints(model), 42 + i, c))
    assert len({tuple(w) for w in wps}) == 6
    out = mj.smooth_path(wps, c, seed=42)
    assert mj.path_length(out) <= mj.path_length(wps)
    np.testing.assert_equal(out[0], wps[0])
    np.testing.assert_equal(out[-1], wps[-1])
    assert oracle.Oracle(model).check(np.array(out), 3).all()


def test_batched_rrt_franka_queries():
    model = models.load("franka_scene_with_obstacles")
    allowed = [("left_finger", "right_finger")]
    joints = [f"joint{i}" for i in range(1, 8)]
    c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, allowed)]
    q_init = model.keyframe("home").qpos.copy()
    B = 32
    goals = np.array([mj.random_config(model, q_init, joints, s, c) for s in range(B)])
    planner = mj.BatchedRRT(model, joints, c, max_planning_time=60, epsilon=0.05, seed=0, goal_biasing_probability=0.1)
    paths = planner.plan(np.tile(q_init, (B, 1)), goals)
    solved = [p for p in paths if p]
    print("batched rrt:", planner.stats)
    assert len(solved) >= B - 2
    for p, g in zip(paths, goals):
        if p:
            np.testing.assert_equal(p[0], q_init)
            np.testing.assert_equal(p[-1], g)
            replay_valid(model, allowed, p, 0.05)


def test_nearest_batch_kernel_matches_numpy():
    import ctypes as C

    import torch

    from mjpl_b200 import _abi

    L = _abi.lib()
    rng = np.random.default_rng(0)
    B, cap, nq = 37, 300, 7
    nodes = rng.normal(size=(B, cap, nq))
    count = rng.integers(1, cap, size=B)
    nodes[3, 0] = np.inf  # a sink root must never win
    count[3] = max(count[3], 5)
    targets = rng.normal(size=(B, nq))
    dn, dc, dt = torch.from_numpy(nodes).cuda(), torch.from_numpy(count).cuda(), torch.from_numpy(targets).cuda()
    out = torch.empty(B, dtype=torch.int64, device="cuda")
    _abi.check(L.mjb_nearest_batch(dn.data_ptr(), cap, nq, dc.data_ptr(), None, dt.data_ptr(), B, out.data_ptr(),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    want = []
    for b in range(B):
        with np.errstate(invalid="ignore"):
            d2 = ((nodes[b, : count[b]] - targets[b]) ** 2).sum(1)
        want.append(int(np.argmin(np.where(np.isfinite(d2), d2, np.inf))))
    np.testing.assert_array_equal(out.cpu().numpy(), want)
    # with an explicit row selection
    rows = torch.tensor([5, 5, 0, 36], device="cuda")
    out2 = torch.empty(4, dtype=torch.int64, device="cuda")
    t2 = dt[[1, 2, 3, 4]].contiguous()
    _abi.check(L.mjb_nearest_batch(dn.data_ptr(), cap, nq, dc.data_ptr(), rows.data_ptr(), t2.data_ptr(), 4, out2.data_ptr(),
                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    for k, (b, ti) in enumerate(zip([5, 5, 0, 36], [1, 2, 3, 4])):
        d2 = ((nodes[b, : count[b]] - targets[ti]) ** 2).sum(1)
        assert int(out2[k]) == int(np.argmin(d2))


def test_batched_rrt_device_driver_many_queries():
    model = models.load("franka_scene_with_obstacles")
    allowed = [("left_finger", "right_finger")]
    joints = [f"joint{i}" for i in range(1, 8)]
    c = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, allowed)]
    q_init = model.keyframe("home").qpos.copy()
    rows = c[1].engine.sweep_rows(3, 0, 4096).double().cpu().numpy()
    rows[:, 7:] = q_init[7:]
    goals = rows[np.asarray(mj.obeys_constraints_batch(rows, c))][:256]
    planner = mj.BatchedRRT(model, joints, c, max_planning_time=60, epsilon=0.05, seed=1, goal_biasing_probability=0.1)
    paths = planner.plan(np.tile(q_init, (len(goals), 1)), goals)
    print("batched rrt (device):", planner.stats)
    assert planner.stats["driver"] == "device"
    assert planner.stats["solved"] >= 0.95 * len(goals)
    orc = oracle.Oracle(model, allowed)
    for p, g in list(zip(paths, goals))[:64]:
        if p:
            np.testing.assert_equal(p[0], q_init)
            np.testing.assert_equal(p[-1], g)
            replay_valid(model, allowed, p, 0.05)


def test_trajectory_revalidation_matches_oracle():
    """trajectory/utils.py:40-41 as one block: the first invalid sample of a dense trajectory is the
    one the reference's per-sample loop would stop at (checked with the CPU oracle)."""
    import oracle
    from mjpl_b200.trajectory import Trajectory, first_invalid_position

    model = models.load("franka_scene_with_obstacles")
    allowed = [("left_finger", "right_finger")]
    cons = [mj.JointLimitConstraint(model), mj.CollisionConstraint(model, allowed)]
    orc = oracle.Oracle(model, allowed)
    q0 = model.keyframe("home").qpos.copy()
    rng = np.random.default_rng(4)
    hits = 0
    for _ in range(6):
        q1 = q0.copy()
        q1[:7] = rng.uniform(model.jnt_range[:7, 0], model.jnt_range[:7, 1])
        pos = q0[None, :] + np.linspace(0, 1, 2001)[1:, None] * (q1 - q0)[None, :]
        traj = Trajectory(dt=0.002, q_init=q0, positions=[p for p in pos], velocities=[], accelerations=[])
        got = first_invalid_position(traj, cons)
        ok, dist, _ = orc.check(pos.astype(np.float32).astype(np.float64), 3, want_dist=True)
        bad = np.flatnonzero(~ok)
        want = int(bad[0]) if len(bad) else -1
        if got != want:  # only a sample inside the 1e-5 band may differ
            lo, hi = sorted((got if got >= 0 else len(pos) - 1, want if want >= 0 else len(pos) - 1))
            assert np.abs(dist[lo:hi + 1]).min() < 1e-5
        hits += want >= 0
    assert hits >= 2


def test_rrt_extend_matches_the_reference_extend_call_by_call():
    """``mjb_rrt_extend`` (nearest node + chain + validity + stop rules + append, all on the device)
    against ``_constrained_extend`` restated on the oracle (reference planning/utils.py:135-164),
    one call at a time: same nearest node, same number of appended nodes, same appended
    configurations, same reached configuration -- for every (tree, target) pair whose chain has no
    row inside the contact band."""
    import ctypes as C
    import os

    import torch

    from mjpl_b200 import _abi

    model = models.load("franka_scene_with_obstacles")
    allowed = [("left_finger", "right_finger")]
    eng = mj.get_engine(model, allowed)
    orc = oracle.Oracle(model, allowed)
    oracle.Oracle.set_threads(len(os.sched_getaffinity(0)))
    rng = np.random.default_rng(77)
    nq, B, cap, kcap, eps = model.nq, 10_000, 96, 48, 0.05
    lo, hi = model.jnt_range[:, 0], model.jnt_range[:, 1]
    # trees: 1..24 valid nodes each (random valid configurations; parents form a chain)
    pool = rng.uniform(lo, hi, size=(400_000, nq))
    pool = pool[orc.check(pool.astype(np.float32).astype(np.float64), 3)]
    count = rng.integers(1, 25, size=B)
    nodes = np.full((B, cap, nq), np.inf)
    parent = np.full((B, cap), -1, dtype=np.int64)
    take = rng.integers(0, len(pool), size=(B, 24))
    for k in range(24):
        has = count > k
        nodes[has, k] = pool[take[has, k]]
        parent[has, k] = k - 1
    # targets: random configurations, some close to a node of the tree, some exactly a node, some out of limits
    targets = rng.uniform(lo, hi, size=(B, nq))
    close = rng.random(B) < 0.3
    targets[close] = nodes[close, 0] + rng.normal(0, 0.04, size=(int(close.sum()), nq))
    far = rng.random(B) < 0.05
    targets[far, 0] = hi[0] + 0.2
    same = rng.random(B) < 0.02
    targets[same] = nodes[same, 0]
    dev = eng.torch_device
    d_nodes, d_parent = torch.from_numpy(nodes).to(dev), torch.from_numpy(parent).to(dev)
    d_count, d_targets = torch.from_numpy(count.astype(np.int64)).to(dev), torch.from_numpy(targets).to(dev)
    d_slots = torch.arange(B, dtype=torch.int64, device=dev)
    reached = torch.empty((B, nq), dtype=torch.float64, device=dev)
    last = torch.empty(B, dtype=torch.int64, device=dev)
    _abi.check(_abi.lib().mjb_rrt_extend(eng._h, d_nodes.data_ptr(), d_parent.data_ptr(), d_count.data_ptr(), cap,
                                         d_slots.data_ptr(), d_targets.data_ptr(), B, eps, kcap, 3, reached.data_ptr(),
                                         last.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    g_nodes, g_parent, g_count = d_nodes.cpu().numpy(), d_parent.cpu().numpy(), d_count.cpu().numpy()
    g_reached, g_last = reached.cpu().numpy(), last.cpu().numpy()
    # ---- the reference's extend on the oracle -----------------------------------------------------
    d2 = ((nodes[:, :24] - targets[:, None, :]) ** 2).sum(-1)
    d2[np.arange(24)[None, :] >= count[:, None]] = np.inf
    nn = np.argmin(d2, axis=1)
    near = nodes[np.arange(B), nn]
    delta = targets - near
    dist = np.linalg.norm(delta, axis=1)
    K = np.minimum(np.where(dist > 0, np.ceil(dist / eps), 0).astype(np.int64), kcap)
    ks = np.arange(kcap)
    real = ks[None, :] < K[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        s = np.minimum((ks[None, :] + 1) * eps / dist[:, None], 1.0)
    chain = near[:, None, :] + s[:, :, None] * delta[:, None, :]
    lands = ((ks[None, :] + 1) * eps >= dist[:, None]) & real
    chain[lands] = np.broadcast_to(targets[:, None, :], chain.shape)[lands]
    flat = chain[real]
    lim_ok = ((flat >= lo) & (flat <= hi)).all(axis=1)                   # limits: the fp64 chain point
    col_ok, cdist, _ = orc.check(flat.astype(np.float32).astype(np.float64), 2, want_dist=True)   # collision: its fp32 rounding
    ok = np.zeros(real.shape, dtype=bool)
    ok[real] = lim_ok & col_ok
    band = np.zeros(real.shape, dtype=bool)
    band[real] = np.abs(cdist) < 1e-5
    prev = np.concatenate([near[:, None, :], chain[:, :-1]], axis=1)
    moved = np.linalg.norm(chain - prev, axis=2) >= 1e-8                 # stop rule of planning/utils.py:153
    good = ok & moved & real
    n_ok = np.where(good.all(axis=1), kcap, np.argmin(good, axis=1))
    n_ok = np.minimum(n_ok, K)
    in_band = (band & (ks[None, :] <= n_ok[:, None])).any(axis=1)
    # ---- compare ----------------------------------------------------------------------------------
    clean = ~in_band
    appended = g_count - count
    assert (appended[clean] == n_ok[clean]).all(), int((appended[clean] != n_ok[clean]).sum())
    want_last = np.where(n_ok > 0, count + n_ok - 1, nn)
    assert (g_last[clean] == want_last[clean]).all()
    want_reached = np.where((n_ok > 0)[:, None], chain[np.arange(B), np.maximum(n_ok - 1, 0)], near)
    np.testing.assert_allclose(g_reached[clean], want_reached[clean], rtol=0, atol=1e-12)
    for b in np.flatnonzero(clean & (n_ok > 0))[:2000]:
        k = int(n_ok[b])
        np.testing.assert_allclose(g_nodes[b, count[b]:count[b] + k], chain[b, :k], rtol=0, atol=1e-12)
        assert g_parent[b, count[b]] == nn[b]
        assert (g_parent[b, count[b] + 1:count[b] + k] == np.arange(count[b], count[b] + k - 1)).all()
    np.testing.assert_array_equal(g_nodes[:, :24][np.arange(24)[None, :] < count[:, None]],
                                  nodes[:, :24][np.arange(24)[None, :] < count[:, None]])   # old nodes untouched
    print(f"rrt_extend: {B} calls, mean appended {appended.mean():.2f}, reached the target {int((n_ok == K).sum())}, "
          f"chains with a row in the band {int(in_band.sum())}")
    assert (same <= (appended == 0)).all() and appended.mean() > 1.0


@pytest.mark.gpu
def test_tree_paths_follow_the_parent_links():
    """``mjb_tree_paths`` (Tree.get_path, reference tree.py:68-81, for many trees at once) against a host walk:
    random forests, chains from random nodes to the root, tree selection through ``d_rows``, padding, and a chain
    that does not fit ``max_depth`` reported as such."""
    import torch

    from mjpl_b200 import _abi

    L = _abi.lib()
    rng = np.random.default_rng(5)
    ntrees, cap = 37, 300
    parent = np.full((ntrees, cap), -1, dtype=np.int64)
    counts = rng.integers(1, cap, size=ntrees)
    for t in range(ntrees):
        for i in range(1, counts[t]):
            parent[t, i] = rng.integers(max(0, i - 4), i)       # deep, thin trees
    rows = rng.integers(0, ntrees, size=64).astype(np.int64)
    first = np.array([rng.integers(0, counts[t]) for t in rows], dtype=np.int64)
    want = []
    for t, f in zip(rows, first):
        chain, i = [], int(f)
        while i >= 0:
            chain.append(i)
            i = int(parent[t, i])
        want.append(chain)
    depth = max(len(c) for c in want)
    dev = torch.device("cuda")
    P, R, F = (torch.from_numpy(a).to(dev) for a in (parent, rows, first))
    for max_depth in (depth, depth + 7):
        steps = torch.empty((len(rows), max_depth), dtype=torch.int64, device=dev)
        length = torch.empty(len(rows), dtype=torch.int64, device=dev)
        _abi.check(L.mjb_tree_paths(P.data_ptr(), cap, R.data_ptr(), F.data_ptr(), len(rows), max_depth,
                                    steps.data_ptr(), length.data_ptr(), None))
        steps, length = steps.cpu().numpy(), length.cpu().numpy()
        for i, chain in enumerate(want):
            assert length[i] == len(chain)
            assert steps[i, :len(chain)].tolist() == chain and (steps[i, len(chain):] == -1).all()
    short = depth - 1
    steps = torch.empty((len(rows), short), dtype=torch.int64, device=dev)
    length = torch.empty(len(rows), dtype=torch.int64, device=dev)
    _abi.check(L.mjb_tree_paths(P.data_ptr(), cap, R.data_ptr(), F.data_ptr(), len(rows), short, steps.data_ptr(), length.data_ptr(), None))
    length = length.cpu().numpy()
    assert all((length[i] == -1) == (len(c) > short) for i, c in enumerate(want))
