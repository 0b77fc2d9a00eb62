"""Host-side planner logic on CPU, with oracle-backed constraint doubles.

These restate the reference's planning tests (test/test_planning_utils.py, test/test_rrt.py,
test/test_tree.py, test/test_utils.py::test_random_config) on the bundled ball models.
"""

import numpy as np
import pytest

import mjpl_b200 as mj
from mjpl_b200 import models
from mjpl_b200.planning.batched_rrt import BatchedRRT
from mjpl_b200.planning.rrt import RRT
from mjpl_b200.planning.tree import Node, Tree
from mjpl_b200.planning.utils import (
    _chain,
    _combine_paths,
    _constrained_extend,
    _step,
    _valid_collision_interval,
    path_length,
    smooth_path,
)

from .doubles import OracleCollisionConstraint, OracleJointLimitConstraint


def cons(model, allowed=()):
    return [OracleJointLimitConstraint(model), OracleCollisionConstraint(model, allowed)]


# ------------------------------------------------------------------ tree (reference test/test_tree.py)
def test_node_equality_and_hash():
    a = Node(np.array([0, 1, 2]), None)
    b = Node(np.array([0, 1, 2]), a)
    c = Node(np.array([3, 4, 5]), a)
    assert a == b and a != c and c == Node(np.array([3, 4, 5]), b) and a != 5
    assert hash(a) == hash(b)


def test_tree_invariants():
    root = Node(np.array([0.0, 0.0]))
    n1, n2 = Node(np.array([1.0, 0.0]), root), Node(np.array([0.0, 1.0]), root)
    n3 = Node(np.array([2.0, 0.0]), n1)
    with pytest.raises(ValueError, match="root node should have no parent"):
        Tree(Node(np.array([1.0]), parent=Node(np.array([0.0]))))
    t = Tree(root)
    for n in (n1, n2, n3):
        t.add_node(n)
    assert len(t.nodes) == 4 and root in t
    with pytest.raises(ValueError, match="already exists"):
        t.add_node(Node(n2.q, n3))
    with pytest.raises(ValueError, match="does not have a parent"):
        t.add_node(Node(np.array([5.0, 5.0])))
    with pytest.raises(ValueError, match="parent is not in the tree"):
        t.add_node(Node(np.array([5.0, 5.0]), Node(np.array([9.0, 9.0]), root)))
    assert t.nearest_neighbor(np.array([1.9, 0.1])) == n3
    assert t.nearest_neighbor(np.array([0.1, 0.8])) == n2
    assert [n.q.tolist() for n in t.get_path(n3)] == [[2, 0], [1, 0], [0, 0]]
    with pytest.raises(ValueError, match="not in the tree"):
        t.get_path(Node(np.array([7.0, 7.0]), root))
    # +inf sink never wins (reference rrt.py:184-188)
    sink = Node(np.ones(2) * np.inf)
    g = Tree(sink)
    g.add_node(Node(np.array([3.0, 3.0]), sink))
    assert g.nearest_neighbor(np.zeros(2)).q.tolist() == [3.0, 3.0]


# ------------------------------------------------------------------ step / chain
def test_step():
    start, target = np.array([0.0, 0.0]), np.array([0.5, 0.0])
    np.testing.assert_equal(_step(start, target, 5.0), target)
    np.testing.assert_allclose(_step(start, target, 0.1), [0.1, 0.0], atol=1e-8)
    np.testing.assert_equal(_step(target, target, np.inf), target)
    with pytest.raises(ValueError, match="`max_step_dist` must be > 0.0"):
        _step(start, target, 0.0)


def test_chain_matches_repeated_step():
    rng = np.random.default_rng(0)
    for _ in range(20):
        a, b = rng.uniform(-3, 3, 6), rng.uniform(-3, 3, 6)
        eps = float(rng.uniform(0.03, 0.5))
        ch = _chain(a, b, eps)
        q, seq = a, []
        while not np.array_equal(q, b):
            q = _step(q, b, eps)
            seq.append(q)
        assert len(seq) == len(ch)
        np.testing.assert_allclose(np.array(seq), ch, atol=1e-12)
        np.testing.assert_equal(ch[-1], b)
        assert np.all(np.linalg.norm(np.diff(np.vstack([a, ch]), axis=0), axis=1) <= eps + 1e-12)
    assert _chain(a, a, 0.1).shape == (0, 6)


# ------------------------------------------------------------------ extend / interval (reference test_planning_utils.py:207-344)
def test_constrained_extend_reaches_target():
    model = models.load("one_dof_ball")
    tree = Tree(Node(np.array([-0.1])))
    q_goal = np.array([0.15])
    q_reached = _constrained_extend(q_goal, tree, 0.1, cons(model))
    np.testing.assert_equal(q_reached, q_goal)
    path = [n.q for n in tree.get_path(tree.nearest_neighbor(q_goal))]
    expected = [q_goal, np.array([0.1]), np.array([0.0]), np.array([-0.1])]
    assert len(path) == len(expected)
    for p, e in zip(path, expected):
        np.testing.assert_allclose(p, e, rtol=0, atol=1e-9)


def test_constrained_extend_stops_before_obstacle():
    model = models.load("one_dof_ball")
    obstacle = model.geom("wall_obstacle")
    min_x = obstacle.pos[0] - obstacle.size[0]
    tree = Tree(Node(np.array([0.0])))
    q_reached = _constrained_extend(np.array([1.0]), tree, 0.1, cons(model))
    assert 0.0 < q_reached[0] < min_x


def test_constrained_extend_with_interval_check():
    model = models.load("one_dof_ball")
    c = cons(model)
    q_init, q_goal = np.array([0.8]), np.array([1.8])
    assert mj.obeys_constraints(q_init, c) and mj.obeys_constraints(q_goal, c)
    tree = Tree(Node(q_init))
    np.testing.assert_equal(_constrained_extend(q_goal, tree, np.inf, c, (0.3, c[1])), q_goal)
    tree = Tree(Node(q_init))
    np.testing.assert_equal(_constrained_extend(q_goal, tree, np.inf, c, (0.1, c[1])), q_init)


def test_constrained_extend_towards_existing_config():
    model = models.load("one_dof_ball")
    root = Node(np.array([0.0]))
    tree = Tree(root)
    np.testing.assert_equal(_constrained_extend(root.q, tree, 0.1, cons(model)), root.q)
    assert tree.nodes == {root}


def test_valid_collision_interval():
    model = models.load("one_dof_ball")
    c = OracleCollisionConstraint(model)
    assert not _valid_collision_interval(np.array([0.8]), np.array([1.5]), 0.1, c)
    assert _valid_collision_interval(np.array([0.8]), np.array([1.5]), 0.2, c)
    assert _valid_collision_interval(np.array([0.0]), np.array([0.2]), 0.01, c)
    with pytest.raises(ValueError, match="step_dist"):
        _valid_collision_interval(np.array([0.0]), np.array([0.2]), 0.0, c)


def test_combine_paths():
    rs, rg = Node(np.array([0.0])), Node(np.array([0.3]))
    cs, cg = Node(np.array([0.1]), rs), Node(np.array([0.2]), rg)
    st, gt = Tree(rs), Tree(rg)
    st.add_node(cs)
    gt.add_node(cg)
    assert [p.tolist() for p in _combine_paths(st, cs, gt, cg)] == [[0.0], [0.1], [0.2], [0.3]]
    q_new = np.array([0.15])
    gs, gg = Node(q_new, cs), Node(q_new, cg)
    st.add_node(gs)
    gt.add_node(gg)
    assert [p.tolist() for p in _combine_paths(st, gs, gt, gg)] == [[0.0], [0.1], [0.15], [0.2], [0.3]]


def test_path_length():
    wps = [np.array([0.0, 0, 0]), np.array([1.0, 0, 0]), np.array([1.0, 1, 0]), np.array([1.0, 1, 1])]
    assert path_length(wps) == pytest.approx(3.0)


# ------------------------------------------------------------------ smoothing (reference test_planning_utils.py:72-196)
def _check_smoothed(smoothed, original, c, eps):
    assert path_length(smoothed) < path_length(original)
    np.testing.assert_equal(smoothed[0], original[0])
    np.testing.assert_equal(smoothed[-1], original[-1])
    for a, b in zip(smoothed[:-1], smoothed[1:]):
        assert np.linalg.norm(b - a) <= eps + 1e-8
    for wp in smoothed:
        assert mj.obeys_constraints(wp, c)


def test_smooth_path_directly_connectable():
    model = models.load("two_dof_ball")
    c = cons(model)
    wps = [np.array(p) for p in [[0.0, 0.0], [0.25, 0.0], [0.25, 0.75], [0.5, 0.75], [1.0, 0.75], [1.0, -0.75], [0.5, -0.75]]]
    out = smooth_path(wps, c, eps=0.1, seed=5)
    assert len(out) > 2
    _check_smoothed(out, wps, c, 0.1)
    sparse = smooth_path(wps, c, eps=0.1, seed=5, sparse=True)
    assert len(sparse) == 2 and sparse[0] is wps[0] and sparse[1] is wps[-1]


def test_smooth_path_around_obstacle():
    model = models.load("two_dof_ball")
    c = cons(model)
    wps = [np.array(p) for p in [[0.0, 0.0], [0.25, 0.0], [0.25, 1.5], [0.5, 1.5], [1.0, 1.5], [1.0, 0.0], [1.0, 0.0]]]
    out = smooth_path(wps, c, eps=0.1, seed=5)
    assert len(out) > 2
    _check_smoothed(out, wps, c, 0.1)
    sparse = smooth_path(wps, c, eps=0.1, seed=5, sparse=True)
    assert len(sparse) > 2 and path_length(sparse) < path_length(wps)
    for wp in sparse:
        assert mj.obeys_constraints(wp, c)


def _smooth_path_one_try_at_a_time(waypoints, constraints, interval, eps, num_tries, seed, sparse):
    """The reference's loop (src/mjpl/planning/utils.py:53-85) as it is written there: the checker for the
    speculative implementation."""
    from mjpl_b200.planning.tree import Node, Tree
    from mjpl_b200.planning.utils import _constrained_extend

    smoothed = waypoints
    rng = np.random.default_rng(seed=seed)
    for _ in range(num_tries):
        start = rng.integers(0, len(smoothed) - 1)
        end = rng.integers(start + 1, len(smoothed))
        tree = Tree(Node(smoothed[start]))
        q_reached = _constrained_extend(smoothed[end], tree, eps, constraints, interval)
        if not np.array_equal(q_reached, smoothed[end]):
            continue
        segment = [n.q for n in tree.get_path(tree.nearest_neighbor(q_reached))]
        if path_length(segment) < path_length(smoothed[start:end + 1]):
            if sparse:
                smoothed = smoothed[:start + 1] + smoothed[end:]
            else:
                segment.reverse()
                smoothed = smoothed[:start] + segment[:-1] + smoothed[end:]
    return smoothed


def test_speculative_smooth_path_equals_the_reference_loop():
    """All remaining tries are evaluated against the current path in one batched validity call and the
    first one the reference would accept is applied: same waypoints as the try-by-try loop for every
    seed, in far fewer validity rounds than tries."""
    from mjpl_b200.planning.rrt import RRT

    model = models.load("two_dof_ball")
    c = cons(model)
    wps = [np.array(p) for p in [[0.0, 0.0], [0.25, 0.0], [0.25, 1.5], [0.5, 1.5], [1.0, 1.5], [1.0, 0.0], [1.0, 0.0]]]
    cases = [(wps, c, None, 0.1)]
    ur = models.load("ur5e_scene")
    cu = cons(ur)
    q0 = ur.keyframe("home").qpos.copy()
    q1 = mj.random_config(ur, q0, mj.all_joints(ur), 3, cu)
    path = RRT(ur, mj.all_joints(ur), cu, max_planning_time=30, epsilon=0.1, seed=3).plan_to_config(q0, q1)
    assert path
    cases.append((path, cu, None, 0.1))
    cases.append((path, cu, (0.05, cu[1]), 0.1))          # with interval checks
    rounds = []
    for way, constraints, interval, eps in cases:
        for seed in (0, 5, 42):
            for sparse in (False, True):
                want = _smooth_path_one_try_at_a_time(list(way), constraints, interval, eps, 60, seed, sparse)
                got = smooth_path(list(way), constraints, interval, eps=eps, num_tries=60, seed=seed, sparse=sparse)
                assert len(got) == len(want)
                for a, b in zip(got, want):
                    np.testing.assert_array_equal(a, b)
                rounds.append(smooth_path.last_launches)
    print('validity rounds for 60 tries:', rounds)
    assert max(rounds) <= 50 and np.mean(rounds) <= 30, rounds     # 60 tries each; every accepted shortcut ends a round


def test_smooth_path_invalid_args():
    with pytest.raises(ValueError, match="waypoints"):
        smooth_path([], [])
    wps = [np.zeros(6), np.ones(6)]
    with pytest.raises(ValueError, match="eps"):
        smooth_path(wps, [], eps=0.0)
    with pytest.raises(ValueError, match="num_tries"):
        smooth_path(wps, [], num_tries=0)


# ------------------------------------------------------------------ RRT (reference test/test_rrt.py)
def _check_plan(wps, q_init, q_goal, eps, c):
    assert len(wps) >= 2
    np.testing.assert_equal(wps[0], q_init)
    np.testing.assert_equal(wps[-1], q_goal)
    for a, b in zip(wps[:-1], wps[1:]):
        assert np.linalg.norm(b - a) <= eps + 1e-12
    for wp in wps:
        assert mj.obeys_constraints(wp, c)


def test_run_rrt():
    model = models.load("one_dof_ball")
    c = cons(model)
    planner = RRT(model, mj.all_joints(model), c, max_planning_time=5.0, epsilon=0.1, seed=42)
    wps = planner.plan_to_config(np.array([-0.2]), np.array([0.35]))
    assert len(wps) > 2
    _check_plan(wps, np.array([-0.2]), np.array([0.35]), 0.1, c)


def test_run_rrt_subset_joints_and_trivial():
    model = models.load("two_dof_ball")
    c = cons(model)
    planner = RRT(model, ["ball_slide_x"], c, max_planning_time=5.0, epsilon=0.1, seed=42)
    wps = planner.plan_to_config(np.array([0.0, 0.0]), np.array([0.3, 0.0]))
    assert len(wps) > 2
    _check_plan(wps, np.array([0.0, 0.0]), np.array([0.3, 0.0]), 0.1, c)
    wps = planner.plan_to_config(np.array([0.0, 0.0]), np.array([0.05, 0.0]))
    assert len(wps) == 2


def test_rrt_around_wall_2dof():
    model = models.load("two_dof_ball")
    c = cons(model)
    planner = RRT(model, mj.all_joints(model), c, max_planning_time=20.0, epsilon=0.1, seed=3)
    q0, q1 = np.array([0.0, 0.0]), np.array([1.2, 0.0])
    wps = planner.plan_to_config(q0, q1)
    assert wps, "planner timed out"
    _check_plan(wps, q0, q1, 0.1, c)
    assert path_length(wps) > 1.2  # had to go around the wall


def test_rrt_invalid_args():
    model = models.load("one_dof_ball")
    joints = mj.all_joints(model)
    with pytest.raises(ValueError, match="max_planning_time"):
        RRT(model, joints, [], max_planning_time=0.0)
    with pytest.raises(ValueError, match="epsilon"):
        RRT(model, joints, [], epsilon=0.0)
    with pytest.raises(ValueError, match="goal_biasing_probability"):
        RRT(model, joints, [], goal_biasing_probability=2.0)
    with pytest.raises(ValueError, match="planning_joints"):
        RRT(model, [], [])
    model = models.load("two_dof_ball")
    planner = RRT(model, ["ball_slide_y"], [], max_planning_time=5.0, epsilon=0.1, seed=42)
    with pytest.raises(ValueError, match="values for joints outside of the planner's planning joints"):
        planner.plan_to_config(np.array([0.0, 0.0]), np.array([0.1, 0.0]))
    c = cons(model)
    planner = RRT(model, mj.all_joints(model), c, seed=1)
    with pytest.raises(ValueError, match="q_init is not a valid configuration"):
        planner.plan_to_config(np.array([0.6, 0.0]), np.array([0.0, 0.0]))
    with pytest.raises(ValueError, match="goal config is not a valid configuration"):
        planner.plan_to_config(np.array([0.0, 0.0]), np.array([0.6, 0.0]))


# ------------------------------------------------------------------ random_config (reference test/test_utils.py:104-129)
def test_random_config():
    model = models.load("two_dof_ball")
    c = cons(model)
    joints = mj.all_joints(model)
    q_init = np.zeros(model.nq)
    a = mj.random_config(model, q_init, joints, 42, c)
    b = mj.random_config(model, q_init, joints, 42, c)
    np.testing.assert_equal(a, b)
    assert mj.obeys_constraints(a, c)
    # same candidate sequence as the reference's one-at-a-time loop
    rng = np.random.default_rng(42)
    while True:
        cand = rng.uniform(*model.jnt_range.T)
        if mj.obeys_constraints(cand, c):
            break
    np.testing.assert_equal(a, cand)
    q = mj.random_config(model, q_init, ["ball_slide_y"], 42, c)
    assert q[mj.qpos_idx(model, ["ball_slide_x"])[0]] == 0.0 and mj.obeys_constraints(q, c)


# ------------------------------------------------------------------ batched lock-step RRT
def test_batched_rrt_many_queries():
    model = models.load("two_dof_ball")
    c = cons(model)
    rng = np.random.default_rng(0)
    qi, qg = [], []
    while len(qi) < 24:
        a, b = rng.uniform(-1.5, 1.5, 2), rng.uniform(-1.5, 1.5, 2)
        if mj.obeys_constraints(a, c) and mj.obeys_constraints(b, c):
            qi.append(a)
            qg.append(b)
    qi[0], qg[0] = np.array([0.0, 0.0]), np.array([1.2, 0.0])  # wall in between
    qi[1], qg[1] = np.array([0.1, 0.1]), np.array([0.12, 0.1])  # trivially connectable
    planner = BatchedRRT(model, mj.all_joints(model), c, max_planning_time=60.0, epsilon=0.1, seed=11)
    paths = planner.plan(np.array(qi), np.array(qg))
    assert planner.stats["solved"] == 24
    assert len(paths[1]) == 2
    for p, a, b in zip(paths, qi, qg):
        _check_plan(p, a, b, 0.1, c)
    assert path_length(paths[0]) > 1.2
    # far fewer validity launches than configurations checked: chains are evaluated as blocks
    assert planner.stats["launches"] * 10 < planner.stats["configs_checked"]
    with pytest.raises(ValueError, match="q_init is not a valid configuration"):
        planner.plan(np.array([[0.6, 0.0]]), np.array([[0.0, 0.0]]))


def test_obeys_constraints_batch_generic_and_fused_paths():
    model = models.load("two_dof_ball")
    c = cons(model)
    Q = np.array([[0.0, 0.0], [0.6, 0.0], [2.5, 0.0], [1.0, 1.9]])
    np.testing.assert_array_equal(mj.obeys_constraints_batch(Q, c), [True, False, False, True])
    np.testing.assert_array_equal(mj.obeys_constraints_batch(Q, []), [True] * 4)
    assert [mj.obeys_constraints(q, c) for q in Q] == [True, False, False, True]
    q0 = Q[0].copy()
    assert mj.apply_constraints(q0, q0, c) is q0  # the same array object comes back when valid
    assert mj.apply_constraints(q0, Q[1].copy(), c) is None


def test_collision_ruleset_semantics():
    # reference test/test_collision_constraint.py:36-192 on the bundled Franka tables
    from mjpl_b200.constraint.collision_constraint import CollisionRuleset

    m = models.load("franka_scene")
    g = {n: m.body_geomadr[m.body(n).id] for n in ("link1", "link2", "link3", "link4", "link5", "link6", "link7", "left_finger", "right_finger")}
    cr = CollisionRuleset(m, [("link1", "link2")])
    assert cr.obeys_ruleset(np.empty((0, 2)))
    assert cr.obeys_ruleset(np.array([[g["link1"], g["link2"]]]))
    assert cr.obeys_ruleset(np.array([[g["link2"], g["link1"]], [g["link1"], g["link2"]]]))
    assert not cr.obeys_ruleset(np.array([[g["left_finger"], g["right_finger"]]]))
    assert not cr.obeys_ruleset(np.array([[g["link1"], g["link2"]], [g["link1"], g["link5"]]]))
    cr = CollisionRuleset(m, [("link1", "link2"), ("link6", "link7"), ("left_finger", "right_finger")])
    assert cr.obeys_ruleset(np.array([[g["left_finger"], g["right_finger"]], [g["link7"], g["link6"]], [g["link2"], g["link1"]]]))
    assert not cr.obeys_ruleset(np.array([[g["link2"], g["link3"]]]))
    cr = CollisionRuleset(m)
    assert cr.obeys_ruleset(np.empty((0, 2))) and not cr.obeys_ruleset(np.array([[g["link3"], g["link2"]]]))
    with pytest.raises(ValueError, match="nx2"):
        cr.obeys_ruleset(np.zeros((1, 3)))
    with pytest.raises(ValueError, match="nx2"):
        cr.obeys_ruleset(np.zeros((1, 2, 1)))
    with pytest.raises(KeyError):
        CollisionRuleset(m, [("link1", "nope")])


def test_pipeline_bins_from_calibration():
    """Host side of the multi-kernel pipeline: every pair has a bin (closed-form pairs -- plane against
    a box / capsule, capsule against capsule -- share bin 0 with the smallest hull scans), and the
    calibrated per-bin item rates that size the bins add up to the calibrated total."""
    from mjpl_b200 import models
    from tests.hostsim import HostSim

    m = models.load("franka_scene_with_obstacles")
    hs = HostSim(m, [("left_finger", "right_finger")])
    be, total, pb = hs.bins()
    pairs = hs.pairs()
    assert len(pb) == len(pairs) == 324
    assert ((pb >= 0) & (pb < 8)).all()
    assert (be >= 0).all() and 0.5 < be.sum() and abs(be.sum() - total) < 1e-9
    mesh = m.geom_type == 7
    both_mesh = mesh[pairs[:, 0]] & mesh[pairs[:, 1]]
    assert both_mesh.any() and (pb[both_mesh] >= 5).all()     # hull against hull: the three largest bins
    one_mesh = mesh[pairs[:, 0]] ^ mesh[pairs[:, 1]]
    assert ((pb[one_mesh] >= 0) & (pb[one_mesh] <= 4)).all()  # hull against a small core (or the plane)
    u = models.load("ur5e_scene")
    be_u, total_u, pb_u = HostSim(u).bins()
    assert (pb_u == 0).all() and be_u[1:].sum() == 0.0        # UR5e: capsules, a cylinder, a plane


def test_bounding_capsules_and_cull_groups_are_conservative():
    """Level 0 (group spheres / world-fixed capsules), the bounding-capsule cull and the OBB cull of the
    pipeline only ever remove pairs that cannot be in contact: every vertex lies inside its shape's
    capsule and its group's sphere, the fp32 segment-segment distance never overestimates by more than
    rounding, and the culled evaluation gives the oracle's answer on every row outside the band."""
    import oracle
    from mjpl_b200 import mjcf, models
    from tests import hostsim, toy_models as toys
    from tests.hostsim import HostSim

    assert hostsim.segseg_check(30000) < 1e-6
    # level 0's squared limits carry 8 eps32 |e|_max^2 for the expanded square; measured: about 2 eps32 |e|_max^2
    assert hostsim.point_segment_check(200000) < 4 * 5.97e-8
    zoo = mjcf.from_xml_string(toys.PRIMITIVE_ARM)
    zoo.geom_margin = np.where(np.arange(zoo.ngeom) % 2 == 0, 0.015, 0.004)   # margins widen the culls too
    cases = [(models.load("franka_scene_with_obstacles"), [("left_finger", "right_finger")], 6000),
             (models.load("franka_scene"), [], 4000), (models.load("ur5e_scene"), [], 4000),
             (mjcf.from_xml_string(toys.PRIMITIVE_ARM), [], 6000), (zoo, [], 6000)]
    for m, allowed, n in cases:
        hs = HostSim(m, allowed)
        assert hs.bounds_check() <= 0.0
        assert 1e-9 < hs.l0_sq_err() < 1e-4      # Franka obstacle scene: reach 2.6 m + segments to 1.5 m -> 8e-6 m^2
        rng = np.random.default_rng(8)
        Q = rng.uniform(m.jnt_range[:, 0], m.jnt_range[:, 1], size=(n, m.nq)).astype(np.float32)
        got, st = hs.check_pipe(Q)
        assert st["level0"] >= st["capsule"] >= st["items"] >= st["contacts"] > 0
        assert st["expanded"] >= st["capsule"]
        want, dist, _ = oracle.Oracle(m, allowed).check(Q.astype(np.float64), 3, want_dist=True)
        bad = got.astype(bool) != want
        assert not (bad & (np.abs(dist) >= 1e-5)).any()
        np.testing.assert_array_equal(got, hs.check(Q)[0])   # same answer as the per-pair sphere + mid-phase order


def test_inner_shapes_only_certify_real_contacts():
    """The certain-contact shortcut (inner capsule per shape, inner ball / tube per cull group): every inner
    capsule lies inside its shape (surface samples against the shape's support function), the shortcut
    fires on a good share of the colliding rows, and NEVER on a row the full evaluation finds free."""
    from mjpl_b200 import mjcf, models
    from tests import toy_models as toys
    from tests.hostsim import HostSim

    zoo = mjcf.from_xml_string(toys.PRIMITIVE_ARM)
    zoo.geom_margin = np.where(np.arange(zoo.ngeom) % 2 == 0, 0.015, 0.004)
    cases = [(models.load("franka_scene_with_obstacles"), [("left_finger", "right_finger")], 20000, 0.6),
             (models.load("franka_scene"), [], 8000, 0.4), (models.load("ur5e_scene"), [], 8000, 0.4),
             (mjcf.from_xml_string(toys.PRIMITIVE_ARM), [], 8000, 0.5), (zoo, [], 8000, 0.1)]
    for m, allowed, n, share in cases:
        hs = HostSim(m, allowed)
        nshape, worst = hs.inner_check(600)
        assert nshape > 0 and worst <= 1e-7, (nshape, worst)
        rng = np.random.default_rng(11)
        Q = rng.uniform(m.jnt_range[:, 0], m.jnt_range[:, 1], size=(n, m.nq)).astype(np.float32)
        inner = hs.check_pipe(Q)[1]["inner"]
        assert inner["false_positives"] == 0
        assert inner["caught_any"] >= inner["caught_level0"] > 0
        assert inner["caught_any"] >= share * inner["invalid_rows"], inner


def test_inner_shapes_on_rows_at_the_contact_boundary():
    """Where a too generous inner shape would show: rows bisected between a free and a colliding configuration
    until the two are 1e-4 rad apart (distances to contact down to 1e-8 m).  The shortcut must not fire on any free
    one, and the culled evaluation must still agree with the oracle outside the 1e-5 band."""
    import oracle
    from mjpl_b200 import models
    from tests.hostsim import HostSim

    for name, allowed in (("franka_scene_with_obstacles", [("left_finger", "right_finger")]), ("ur5e_scene", [])):
        m = models.load(name)
        hs, orc = HostSim(m, allowed), oracle.Oracle(m, allowed)
        rng = np.random.default_rng(3)
        Q = rng.uniform(m.jnt_range[:, 0], m.jnt_range[:, 1], size=(4000, m.nq))
        ok = orc.check(Q, 2)
        n = min(int(ok.sum()), int((~ok).sum()), 600)
        A, B = Q[ok][:n].copy(), Q[~ok][:n].copy()
        mids = []
        for _ in range(14):
            M = 0.5 * (A + B)
            v = orc.check(M, 2)
            A[v], B[~v] = M[v], M[~v]
            mids.append(M.copy())
        R = np.concatenate([A, B] + mids[-4:]).astype(np.float32)
        got, st = hs.check_pipe(R)
        want, dist, _ = orc.check(R.astype(np.float64), 2, want_dist=True)
        assert np.abs(dist).min() < 1e-6 and 0.3 < want.mean() < 0.7
        assert st["inner"]["false_positives"] == 0
        assert not ((got.astype(bool) != want) & (np.abs(dist) >= 1e-5)).any()


def test_support_maps_return_the_full_scan_maximum():
    """The cube-map support tables of the larger hulls (narrow_kernel) list, per cell, a rigorous superset of
    the vertices that can be a support for a direction of the cell: on random, near-axis and cell-border
    directions the mapped support value equals the full scan's, with a handful of candidates per query."""
    from mjpl_b200 import models
    from tests.hostsim import HostSim

    hs = HostSim(models.load("franka_scene_with_obstacles"), [("left_finger", "right_finger")])
    nmapped, worst, avg = hs.smap_check(40000, seed=9)
    assert nmapped == 13 and worst == 0.0 and 2.0 < avg < 9.0
    assert HostSim(models.load("ur5e_scene")).smap_check(100)[0] == 0      # no hulls: nothing to map


def test_min_distance_core_matches_oracle():
    """mjb_min_distance's fp64 routine (GJK to convergence, expanding-polytope depth, closed forms, capsule
    pruning) against the oracle's signed distance: equal to 1e-9 wherever the oracle's own value is exact
    (it only evaluates pairs within 1 mm of its bounding spheres), same sign everywhere."""
    import oracle
    from mjpl_b200 import mjcf, models
    from tests import toy_models as toys
    from tests.hostsim import HostSim

    zoo = mjcf.from_xml_string(toys.PRIMITIVE_ARM)
    zoo_m = mjcf.from_xml_string(toys.PRIMITIVE_ARM)
    zoo_m.geom_margin = np.where(np.arange(zoo_m.ngeom) % 2 == 0, 0.015, 0.004)
    for m, allowed, n in ((models.load("franka_scene_with_obstacles"), [("left_finger", "right_finger")], 1500),
                          (models.load("ur5e_scene"), [], 4000), (zoo, [], 4000), (zoo_m, [], 4000)):
        rng = np.random.default_rng(5)
        Q = rng.uniform(m.jnt_range[:, 0], m.jnt_range[:, 1], size=(n, m.nq)).astype(np.float32)
        d, p = HostSim(m, allowed).min_distance(Q, 0.01)
        _, od, _ = oracle.Oracle(m, allowed).check(Q.astype(np.float64), 2, want_dist=True)
        assert ((d <= 0) == (od <= 0)).all()
        exact = np.minimum(d, od) < 1e-3
        assert exact.sum() > 100 and np.abs(d - od)[exact].max() < 1e-9
        assert (d <= 0.01).all() and (d >= -1e-3).all() and ((p >= 0) == (d < 0.01)).all()


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): rank 0 prints one JSON
    line with the contract's keys, other ranks print nothing; no GPU involved."""
    import json
    import os
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    cmd = [sys.executable, str(root / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, "RANK": "0"})
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "collision-checked configs/sec" and d["unit"] == "configs/s"
    assert d["higher_is_better"] is True and d["value"] > 1e4 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    other = subprocess.run(cmd, capture_output=True, text=True, timeout=60, env={**os.environ, "RANK": "1"})
    assert other.returncode == 0 and other.stdout.strip() == ""
