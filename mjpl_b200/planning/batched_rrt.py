"""Many independent bi-directional RRT queries advanced in lock-step.

This is the batched planning entry point the north star asks for: query ``i`` runs the same
CBiRRT as :class:`mjpl_b200.planning.rrt.RRT` (same sampling distribution, extend / connect /
swap structure of the reference's ``plan_to_configs``, ``src/mjpl/planning/rrt.py:195-235``,
and the stop rules of ``_constrained_extend``, ``planning/utils.py:139-163``), but every
iteration gathers the extend chains of ALL unsolved queries into one block of configurations
and validates it with a single fused kernel launch.  Queries are independent (own trees, own
random stream seeded ``seed + i``), so they shard across GPUs without communication.

Trees are stored as padded ``(nqueries, capacity, nq)`` arrays; nearest neighbours are one
vectorised reduction per iteration for all queries.

Two drivers share the algorithm:

* the **device driver** (used when the constraints are this package's CUDA constraints on one
  model): forests live in HBM as fp64 tensors, nearest neighbours come from
  ``mjb_nearest_batch`` (one warp per tree), chains are built, validated
  (``mjb_check_configs``) and appended without leaving the GPU; the host only learns, once per
  iteration, which queries have met.  Sampling uses one vectorised random stream for all queries.
* the **host driver** (any other ``Constraint``, e.g. the oracle-backed doubles of the CPU test
  suite): the same steps in numpy with one ``valid_configs`` call per extend phase and one random
  stream per query (``seed + i``).
"""

from __future__ import annotations

import os
import time

import numpy as np

from ..constraint.constraint_interface import Constraint
from ..constraint.utils import apply_constraints_batch, _fusable, obeys_constraints_batch
from ..utils import qpos_idx


def _row_norms(D: np.ndarray) -> np.ndarray:
    """Euclidean norm of every row.  Small blocks are computed the way ``np.linalg.norm`` does for
    one vector (``sqrt(x.dot(x))``), so that a batched step is bit-identical to the reference's
    scalar step (the tests compare whole paths with the sequential planner); large blocks use one
    vectorised reduction, which can differ from that in the last bit."""
    if len(D) <= 64:
        return np.sqrt(np.array([x.dot(x) for x in D], dtype=np.float64)) if len(D) else np.zeros(0)
    return np.sqrt(np.einsum("ij,ij->i", D, D))


class _Forest:
    """``B`` trees in padded arrays: q (B,cap,nq), parent (B,cap), count (B,)."""

    def __init__(self, roots: np.ndarray, cap: int = 256):
        B, nq = roots.shape
        self.q = np.full((B, cap, nq), np.inf)
        self.parent = np.full((B, cap), -1, dtype=np.int64)
        self.count = np.ones(B, dtype=np.int64)
        self.q[:, 0] = roots

    def _grow(self, need: int):
        cap = self.q.shape[1]
        if need <= cap:
            return
        new = max(need, 2 * cap)
        B, _, nq = self.q.shape
        q = np.full((B, new, nq), np.inf)
        q[:, :cap] = self.q
        p = np.full((B, new), -1, dtype=np.int64)
        p[:, :cap] = self.parent
        self.q, self.parent = q, p

    def nearest(self, rows: np.ndarray, targets: np.ndarray) -> np.ndarray:
        """index of the nearest node of tree rows[i] to targets[i]"""
        n = int(self.count[rows].max())
        d = self.q[rows, :n] - targets[:, None, :]
        d2 = np.einsum("bij,bij->bi", d, d)  # unused slots hold +inf
        return np.argmin(np.nan_to_num(d2, nan=np.inf), axis=1)

    def append_chains(self, rows, parents, chains, lengths):
        """chains: (len(rows), Kmax, nq) padded; appends chains[i][:lengths[i]] under parents[i];
        returns the index of the last node of each chain (or the parent when length 0)."""
        self._grow(int((self.count[rows] + lengths).max()))
        last = parents.copy()
        for i in np.flatnonzero(lengths > 0):
            b, k, c = rows[i], int(lengths[i]), int(self.count[rows[i]])
            self.q[b, c : c + k] = chains[i, :k]
            self.parent[b, c] = parents[i]
            if k > 1:
                self.parent[b, c + 1 : c + k] = np.arange(c, c + k - 1)
            self.count[b] = c + k
            last[i] = c + k - 1
        return last

    def append_one(self, rows, parents, q):
        """one node per tree (``rows`` are distinct trees); returns the new node indices"""
        c = self.count[rows].copy()
        self._grow(int(c.max()) + 1)
        self.q[rows, c] = q
        self.parent[rows, c] = parents
        self.count[rows] = c + 1
        return c

    def path_to_root(self, b: int, idx: int) -> list[np.ndarray]:
        out = []
        while idx >= 0:
            out.append(self.q[b, idx].copy())
            idx = int(self.parent[b, idx])
        return out


def _paths_of_solved(L, stream, solved, nodes_s, parent_s, first_s, nodes_g, parent_g, first_g, max_depth):
    """Tree.get_path for every solved slot: one ``mjb_tree_paths`` launch per tree, the rows gathered on the device
    in query-major order, one D2H copy per tree.  -> per solved slot the list of waypoints root of the start tree
    .. connecting node .. root of the goal tree (the connecting configuration once if both trees hold it)."""
    import torch

    from .. import _abi

    n = len(solved)
    dev = solved.device
    cap = parent_s.shape[1]

    def chains(nodes, parent, first):
        steps = torch.empty((n, max_depth), dtype=torch.int64, device=dev)
        length = torch.empty(n, dtype=torch.int64, device=dev)
        _abi.check(L.mjb_tree_paths(parent.data_ptr(), cap, solved.data_ptr(), first.data_ptr(), n, max_depth,
                                    steps.data_ptr(), length.data_ptr(), stream()))
        counts = length.cpu().numpy()
        if (counts < 0).any():
            raise RuntimeError("a parent chain is longer than its tree")
        steps = steps[:, :max(int(counts.max()), 1)]
        valid = steps >= 0
        sl = solved[:, None].expand_as(steps)[valid]
        return nodes[sl, steps[valid]].cpu().numpy(), counts          # query by query, first node .. root

    assert parent_s.is_contiguous() and parent_g.is_contiguous() and parent_g.shape[1] == cap
    rs, cs = chains(nodes_s, parent_s, first_s.contiguous())
    rg, cg = chains(nodes_g, parent_g, first_g.contiguous())
    os_, og = np.concatenate(([0], np.cumsum(cs))), np.concatenate(([0], np.cumsum(cg)))
    same = (rs[os_[:-1]] == rg[og[:-1]]).all(axis=1)                   # the two extends met in one configuration
    out = []
    for i in range(n):
        a = rs[os_[i] + (1 if same[i] else 0):os_[i + 1]][::-1]        # root of the start tree -> connecting node
        out.append(list(a) + list(rg[og[i]:og[i + 1]]))               # connecting node -> root of the goal tree
    return out


class BatchedRRT:
    """Lock-step bi-RRT over many (q_init, q_goal) queries.

    With non-projecting constraints whole extend chains are validated at once (device driver when
    the constraints fuse into the engine).  With a projecting constraint (``PoseConstraint``) the
    reference's step-by-step extend (``planning/utils.py:139-164``) is kept, but every step is
    taken by all queries together: one projection block and one validity block per step."""

    def __init__(self, model, planning_joints: list[str], constraints: list[Constraint],
                 max_planning_time: float = 10.0, epsilon: float = 0.05, seed: int | None = None,
                 goal_biasing_probability: float = 0.05, max_iterations: int = 1000000,
                 max_chain: int = 512, max_active: int = 4096, max_iterations_per_query: int = 3000,
                 sync_every: int = 8, use_cuda_graph: bool = True, device_projection: bool = True) -> None:
        if not planning_joints:
            raise ValueError("`planning_joints` cannot be empty.")
        if max_planning_time <= 0.0:
            raise ValueError("`max_planning_time` must be > 0.0")
        if epsilon <= 0.0:
            raise ValueError("`epsilon` must be > 0.0")
        if goal_biasing_probability < 0.0 or goal_biasing_probability > 1.0:
            raise ValueError("`goal_biasing_probability` must be within [0.0, 1.0].")
        self.model = model
        self.planning_joints = planning_joints
        self.constraints = constraints
        self.max_planning_time = max_planning_time
        self.epsilon = epsilon
        self.seed = seed
        self.goal_biasing_probability = goal_biasing_probability
        self.max_iterations = max_iterations
        self.max_chain = max_chain
        self.max_active = max_active                              # slots of the device driver
        self.max_iterations_per_query = max_iterations_per_query  # a query that exceeds it returns []
        self.sync_every = sync_every            # device driver: iterations enqueued between two looks of the host
        self.use_cuda_graph = use_cuda_graph    # device driver: replay iteration pairs as a CUDA graph
        self.device_projection = device_projection  # projecting constraint: device ticks (False: lock-step host driver,
        #                                             one random stream per query as in the sequential planner)
        self.stats: dict = {}

    # one extend for a set of queries: returns reached configs and node indices
    def _extend(self, forest: _Forest, rows: np.ndarray, targets: np.ndarray):
        eps = self.epsilon
        near_idx = forest.nearest(rows, targets)
        near = forest.q[rows, near_idx]
        d = targets - near
        dist = np.linalg.norm(d, axis=1)
        k = np.where(dist > 0, np.ceil(dist / eps), 0).astype(np.int64)
        k = np.minimum(k, self.max_chain)
        kmax = int(k.max()) if len(k) else 0
        nq = near.shape[1]
        if kmax == 0:
            return near, near_idx
        steps = np.arange(1, kmax + 1, dtype=np.float64)[None, :, None]  # (1,K,1)
        with np.errstate(divide="ignore", invalid="ignore"):
            frac = np.where(dist > 0, eps / dist, 0.0)[:, None, None] * steps
        chains = near[:, None, :] + np.minimum(frac, 1.0) * d[:, None, :]
        # the step that covers the remaining distance lands on the target itself
        full = (k * eps >= dist) & (k > 0)
        chains[np.flatnonzero(full), k[full] - 1] = targets[full]
        mask = np.arange(kmax)[None, :] < k[:, None]  # (n,K) real rows
        flat = chains[mask]
        ok_flat = np.asarray(obeys_constraints_batch(flat, self.constraints)).astype(bool)
        self.stats["configs_checked"] = self.stats.get("configs_checked", 0) + int(len(flat))
        self.stats["launches"] = self.stats.get("launches", 0) + 1
        ok = np.zeros(mask.shape, dtype=bool)
        ok[mask] = ok_flat
        # stop rule: a step shorter than 1e-8 ends the chain (reference planning/utils.py:153)
        prev = np.concatenate([near[:, None, :], chains[:, :-1]], axis=1)
        ok &= np.linalg.norm(chains - prev, axis=2) >= 1e-8
        good = np.where(ok.all(axis=1), k, np.argmin(ok, axis=1))
        good = np.minimum(good, k)
        last = forest.append_chains(rows, near_idx, chains, good)
        return forest.q[rows, last], last

    # one extend with a projecting constraint: all queries take each step together
    def _extend_projected(self, forest: _Forest, rows: np.ndarray, targets: np.ndarray):
        eps = self.epsilon
        last = forest.nearest(rows, targets)
        q_old = forest.q[rows, last].copy()
        alive = np.ones(len(rows), dtype=bool)
        for _ in range(self.max_chain):
            alive &= ~np.all(q_old == targets, axis=1)      # reference: `if np.array_equal(q_target, q): return q`
            idx = np.flatnonzero(alive)
            if not len(idx):
                break
            cur, tgt = q_old[idx], targets[idx]
            d = tgt - cur
            dist = _row_norms(d)
            with np.errstate(divide="ignore", invalid="ignore"):
                q = np.where((dist <= eps)[:, None], tgt, cur + d * (eps / dist)[:, None])   # _step
            qn, ok = apply_constraints_batch(cur, q, self.constraints)
            self.stats["configs_checked"] = self.stats.get("configs_checked", 0) + len(idx)
            self.stats["launches"] = self.stats.get("launches", 0) + 1
            # the reference's stop rules (:151-160): no progress, or further from the target than before
            ok &= _row_norms(qn - cur) >= 1e-8
            ok &= _row_norms(tgt - qn) <= dist
            good = idx[ok]
            if len(good):
                last[good] = forest.append_one(rows[good], last[good], qn[ok])
                q_old[good] = qn[ok]
            alive[idx[~ok]] = False
        return q_old, last

    def plan(self, q_inits: np.ndarray, q_goals: np.ndarray) -> list[list[np.ndarray]]:
        """``q_inits``, ``q_goals``: (B, nq).  Returns one waypoint list per query (empty on failure)."""
        q_inits = np.ascontiguousarray(q_inits, dtype=np.float64)
        q_goals = np.ascontiguousarray(q_goals, dtype=np.float64)
        if q_inits.shape != q_goals.shape or q_inits.ndim != 2:
            raise ValueError("q_inits and q_goals must both be (B, nq)")
        fused = _fusable(self.constraints)
        if fused is not None:
            return self._plan_device(q_inits, q_goals, *fused)
        proj = self._projected_on_device() if self.device_projection else None
        if proj is not None:
            return self._plan_device_projected(q_inits, q_goals, *proj)
        return self._plan_host(q_inits, q_goals)

    def _projected_on_device(self):
        """``[JointLimitConstraint?, PoseConstraint, CollisionConstraint?]`` on one model, the collision
        constraint (if any) AFTER the pose constraint, so that it sees the projected configuration just as
        the final re-validation does: that list runs as device ticks (``mjb_cbirrt_tick``).  Returns
        ``(engine, pose_constraint, flags, limits_before)`` or None (any other list: the host driver)."""
        from ..constraint.collision_constraint import CollisionConstraint
        from ..constraint.joint_limit_constraint import JointLimitConstraint
        from ..constraint.pose_constraint import PoseConstraint

        pose, coll, limits_before, seen_pose = None, None, False, False
        for c in self.constraints:
            if type(c) is PoseConstraint and pose is None:
                pose, seen_pose = c, True
            elif type(c) is JointLimitConstraint:
                limits_before = limits_before or not seen_pose
            elif type(c) is CollisionConstraint and coll is None and seen_pose:
                coll = c
            else:
                return None
            if c.model is not self.model:
                return None
        if pose is None:
            return None
        eng = coll.engine if coll is not None else pose.engine
        return eng, pose, (2 if coll is not None else 0), limits_before

    def plan_to_poses(self, q_inits: np.ndarray, poses, site: str, solver=None) -> list[list[np.ndarray]]:
        """One pose goal per query (the reference's ``RRT.plan_to_pose``, rrt.py:69-139, for a block of
        queries): batched IK from each query's ``q_init``, then ``plan`` on the solved ones.  Queries
        whose pose has no IK solution return ``[]``."""
        q_inits = np.ascontiguousarray(q_inits, dtype=np.float64)
        if q_inits.ndim != 2 or len(poses) != len(q_inits):
            raise ValueError("q_inits must be (B, nq) with one pose per query")
        if solver is None:
            from ..inverse_kinematics import DLSIKSolver

            solver = DLSIKSolver(model=self.model, joints=self.planning_joints, constraints=self.constraints,
                                 seed=self.seed, max_attempts=5)
        goals, solved = solver.solve_ik_batch(list(poses), site, q_inits)
        out: list[list[np.ndarray]] = [[] for _ in range(len(q_inits))]
        idx = np.flatnonzero(solved)
        if len(idx):
            for i, path in zip(idx, self.plan(q_inits[idx], goals[idx])):
                out[i] = path
        self.stats["ik_solved"] = int(solved.sum())
        return out

    # ------------------------------------------------------------------ device driver
    def _plan_device(self, q_inits, q_goals, eng, flags):
        """All planner state lives in HBM and whole iterations are enqueued without a host round trip:
        sampling (``mjb_rrt_sample``), the two extends (``mjb_rrt_extend_masked``: nearest node, chain,
        validity, stop rules, append) and the connection test (``mjb_rrt_meet``) only read and write
        device memory, and an even + odd iteration pair is captured ONCE into a CUDA graph that is then
        replayed; the host looks at a four-word status every ``sync_every`` iterations (how many queries
        are still active, how full the trees are).  Queries are processed in waves of ``max_active``."""
        import torch

        dev = eng.torch_device
        B, nq = q_inits.shape
        with torch.cuda.device(eng.device):
            QI = torch.from_numpy(q_inits).to(dev)
            QG = torch.from_numpy(q_goals).to(dev)
            ok = eng.valid_configs(torch.cat([QI, QG]), flags)
            if not bool(ok[:B].all()):
                raise ValueError("q_init is not a valid configuration")
            if not bool(ok[B:].all()):
                bad = q_goals[int((~ok[B:]).nonzero()[0])]
                raise ValueError(f"The following goal config is not a valid configuration: {bad}")
            q_idx = qpos_idx(self.model, self.planning_joints)
            fixed = [i for i in range(nq) if i not in set(q_idx)]
            if fixed and not np.allclose(q_inits[:, fixed], q_goals[:, fixed], rtol=0, atol=1e-12):
                raise ValueError("goal configs have values for joints outside of the planner's planning joints "
                                 "that don't match q_init")
            direct = np.linalg.norm(q_goals - q_inits, axis=1) <= self.epsilon
            paths: list[list[np.ndarray]] = [[] for _ in range(B)]
            for b in np.flatnonzero(direct):
                paths[b] = [q_inits[b].copy(), q_goals[b].copy()]
            pending = np.flatnonzero(~direct)
            S = int(min(len(pending), int(self.max_active)))
            self.stats = {"iterations": 0, "configs_checked": 0, "launches": 0, "driver": "device", "slots": S, "gave_up": 0,
                          "graph_replays": 0, "host_syncs": 0, "waves": 0, "chains_clipped_at_capacity": 0}
            eng.reset_stats()
            t0 = time.time()
            for w0 in range(0, len(pending), max(S, 1)):
                ids = pending[w0:w0 + S]
                left = self.max_planning_time - (time.time() - t0)
                if left <= 0:
                    break
                for qid, path in zip(ids, self._run_wave(eng, flags, QI[ids], QG[ids], left, wave=w0)):
                    paths[int(qid)] = path
                self.stats["waves"] += 1
            self.stats["solved"] = int(sum(1 for p in paths if p))
            self.stats["seconds"] = time.time() - t0
            est = eng.stats()
            self.stats["configs_checked"] = est["rows"]
            self.stats["chains_clipped_at_capacity"] = est["queue_overflow"]
            return paths

    def _run_wave(self, eng, flags, QI, QG, time_left, wave=0):
        """One wave of queries, one slot each (see ``_plan_device``) -> list of paths (``[]`` = unsolved)."""
        import ctypes as C

        import torch

        from .. import _abi

        L = _abi.lib()
        dev, f64, i64 = eng.torch_device, torch.float64, torch.int64
        S, nq = QI.shape
        t_wave0 = time.perf_counter()
        eps, kcap = float(self.epsilon), int(self.max_chain)
        q_idx = qpos_idx(self.model, self.planning_joints)
        plan_mask = torch.zeros(nq, dtype=torch.uint8, device=dev)
        plan_mask[q_idx] = 1
        lo = torch.from_numpy(np.ascontiguousarray(self.model.jnt_range[:, 0], dtype=np.float64)).to(dev)
        hi = torch.from_numpy(np.ascontiguousarray(self.model.jnt_range[:, 1], dtype=np.float64)).to(dev)
        # the longest chain an extend can build: the diameter of the sampled box over epsilon
        span = float(np.linalg.norm((self.model.jnt_range[:, 1] - self.model.jnt_range[:, 0])[q_idx]))
        chain_max = int(min(kcap, np.ceil(span / eps) + 1))
        pairs_per_sync = max(1, int(self.sync_every) // 2)
        headroom = 2 * pairs_per_sync * chain_max + 1      # nodes one tree can gain between two looks of the host
        seed = int(self.seed if self.seed is not None else np.random.SeedSequence().entropy % (1 << 62)) + 7919 * int(wave)

        class Forest:
            def __init__(self, roots, cap):
                self.cap = cap
                self.q = torch.empty((S, cap, nq), dtype=f64, device=dev)
                self.parent = torch.full((S, cap), -1, dtype=i64, device=dev)
                self.count = torch.ones(S, dtype=i64, device=dev)
                self.q[:, 0] = roots

            def grow(self, cap):
                q = torch.empty((S, cap, nq), dtype=f64, device=dev)
                q[:, : self.cap] = self.q
                p = torch.full((S, cap), -1, dtype=i64, device=dev)
                p[:, : self.cap] = self.parent
                self.q, self.parent, self.cap = q, p, cap

        cap0 = 1 << int(np.ceil(np.log2(max(2 * headroom, 1024))))
        start, goal = Forest(QI, cap0), Forest(QG, cap0)
        QI, QG = QI.contiguous(), QG.contiguous()
        active = torch.ones(S, dtype=torch.uint8, device=dev)
        age = torch.zeros(S, dtype=i64, device=dev)
        res_s = torch.full((S,), -1, dtype=i64, device=dev)
        res_g = torch.full((S,), -1, dtype=i64, device=dev)
        counters = torch.zeros(8, dtype=i64, device=dev)
        counters[3] = S
        slots = torch.arange(S, dtype=i64, device=dev)
        targets = torch.empty((S, nq), dtype=f64, device=dev)
        qa, qb = torch.empty_like(targets), torch.empty_like(targets)
        ia, ib = torch.empty(S, dtype=i64, device=dev), torch.empty(S, dtype=i64, device=dev)
        status = torch.empty(4, dtype=i64, device=dev)

        def stream():
            return C.c_void_p(torch.cuda.current_stream().cuda_stream)

        def extend(F, tgt, reached, last):
            _abi.check(L.mjb_rrt_extend_masked(eng._h, F.q.data_ptr(), F.parent.data_ptr(), F.count.data_ptr(), F.cap,
                                               slots.data_ptr(), tgt.data_ptr(), active.data_ptr(), S, eps, kcap, flags,
                                               reached.data_ptr(), last.data_ptr(), stream()))

        def half(fa, fb):
            _abi.check(L.mjb_rrt_sample(seed, counters.data_ptr(), S, nq, QI.data_ptr(), QG.data_ptr(), plan_mask.data_ptr(),
                                        lo.data_ptr(), hi.data_ptr(), float(self.goal_biasing_probability), active.data_ptr(),
                                        targets.data_ptr(), stream()))
            extend(fa, targets, qa, ia)
            extend(fb, qa, qb, ib)
            _abi.check(L.mjb_rrt_meet(S, nq, qa.data_ptr(), qb.data_ptr(), ia.data_ptr(), ib.data_ptr(),
                                      int(self.max_iterations_per_query), active.data_ptr(), age.data_ptr(), res_s.data_ptr(),
                                      res_g.data_ptr(), counters.data_ptr(), stream()))

        def pair():            # an even and an odd iteration: the trees swap roles (rrt.py:231-235)
            half(start, goal)
            half(goal, start)

        def look():            # the host's only read-back: iteration, active slots, fullest trees
            status[0], status[1] = counters[0], counters[3]
            status[2], status[3] = start.count.max(), goal.count.max()
            self.stats["host_syncs"] += 1
            return [int(x) for x in status.cpu()]

        side = torch.cuda.Stream(device=dev)
        graph = None

        def capture():
            # capture_begin / capture_end directly: the torch.cuda.graph context manager first runs gc.collect() and
            # empty_cache() (40 ms per capture on this workload, more with the previous batch's paths still alive)
            g = torch.cuda.CUDAGraph()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                g.capture_begin()
                try:
                    pair()
                finally:
                    g.capture_end()
            torch.cuda.current_stream().wait_stream(side)
            return g

        with eng._call_lock:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                pair()             # eager once: the library grows its scratch here, never inside a capture
            side.synchronize()
            def try_capture():     # capture not possible (e.g. a profiler that forbids it): enqueue eagerly
                try:
                    return capture()
                except Exception:
                    return None

            graph = try_capture() if self.use_cuda_graph else None
            t0 = time.time()
            small_tail = False
            _abi.check(L.mjb_set_chain_hint(eng._h, 0))
            trace = [] if os.environ.get("MJPL_PLAN_TRACE") else None
            while True:
                if trace is not None:
                    trace.append(time.perf_counter())
                it, n_active, hs, hg = look()
                if trace is not None:
                    trace.append(time.perf_counter())
                if n_active == 0 or it >= self.max_iterations or time.time() - t0 >= time_left:
                    break
                need = max(hs, hg) + headroom
                if need > start.cap:
                    new_cap = 1 << int(np.ceil(np.log2(need + headroom)))
                    start.grow(new_cap); goal.grow(new_cap)
                    graph = try_capture() if (graph is not None) else None
                # the tail: a few hard queries are left and an extend checks a few hundred rows.  Tell the library
                # (launch choices only: the one-warp-per-row kernel alone, a small fp64 grid) and capture again.
                if not small_tail and n_active * 24 <= 4096:
                    small_tail = True
                    _abi.check(L.mjb_set_chain_hint(eng._h, max(1, n_active * 24)))
                    graph = try_capture() if (graph is not None) else None
                if graph is not None:
                    for _ in range(pairs_per_sync):
                        graph.replay()
                    self.stats["graph_replays"] += pairs_per_sync
                else:
                    with torch.cuda.stream(side):
                        for _ in range(pairs_per_sync):
                            pair()
                    side.synchronize()
            t_loop1 = time.perf_counter()
            if trace is not None and len(trace) > 4:
                print(f"[plan trace] setup {1e3 * (trace[0] - t_wave0):.1f} ms, loop {1e3 * (t_loop1 - trace[0]):.1f} ms")
                tr = np.array(trace)
                look_ms = (tr[1::2] - tr[0::2]) * 1e3          # enqueue of the status reads + wait for the GPU to get there
                enq_ms = (tr[2::2] - tr[1:-1:2]) * 1e3         # growth / re-capture + enqueue of the replays
                print(f"[plan trace] syncs {len(look_ms)}: waits sum {look_ms.sum():.1f} ms max {look_ms.max():.1f}; "
                      f"enqueue sum {enq_ms.sum():.1f} ms max {enq_ms.max():.1f}, over 5 ms: {np.round(enq_ms[enq_ms > 5], 1).tolist()}")
            torch.cuda.current_stream().wait_stream(side)
            it, n_active, hs, hg = look()
            if small_tail:
                _abi.check(L.mjb_set_chain_hint(eng._h, 0))
            self.stats["iterations"] = max(self.stats["iterations"], it)
            self.stats["gave_up"] += int(counters[2])
            # ---- paths of the solved slots: walk the parent links of both trees on the device ----------
            out: list[list[np.ndarray]] = [[] for _ in range(S)]
            solved = (res_s >= 0).nonzero(as_tuple=True)[0]
            if len(solved):
                tp = [time.perf_counter()]
                found = _paths_of_solved(L, stream, solved, start.q, start.parent, res_s[solved], goal.q, goal.parent, res_g[solved],
                                         max(int(hs), int(hg), 1))     # a chain cannot be longer than its tree has nodes
                for k, path in zip(solved.cpu().numpy(), found):
                    out[int(k)] = path
            if trace is not None:
                tp.append(time.perf_counter())
                print(f"[plan trace] paths to the host {1e3 * (time.perf_counter() - t_loop1):.1f} ms, of which extraction {1e3 * (tp[-1] - tp[0]):.1f} ms")
            return out

    # ------------------------------------------------------------------ device driver, projecting constraint
    def _plan_device_projected(self, q_inits, q_goals, eng, pose, flags, limits_before):
        """BASELINE configs[3] on the device: every query is a small state machine in HBM and one
        ``mjb_cbirrt_tick`` advances ALL of them by one projected extend step (or one setup); ticks are
        captured into a CUDA graph and replayed, the host reads a status word every ``sync_every`` ticks.
        Same algorithm per query as the sequential planner (sampling distribution, step-by-step extend
        with projection, stop rules, connection test, swap); the random stream is counter based."""
        import ctypes as C

        import torch

        from .. import _abi
        from ..constraint.utils import obeys_constraints_batch as _obeys

        L = _abi.lib()
        dev, f64, i64 = eng.torch_device, torch.float64, torch.int64
        B, nq = q_inits.shape
        ok = np.asarray(_obeys(np.concatenate([q_inits, q_goals]), self.constraints)).astype(bool)
        if not ok[:B].all():
            raise ValueError("q_init is not a valid configuration")
        if not ok[B:].all():
            bad = q_goals[np.flatnonzero(~ok[B:])[0]]
            raise ValueError(f"The following goal config is not a valid configuration: {bad}")
        q_idx = qpos_idx(self.model, self.planning_joints)
        fixed = [i for i in range(nq) if i not in set(q_idx)]
        if fixed and not np.allclose(q_inits[:, fixed], q_goals[:, fixed], rtol=0, atol=1e-12):
            raise ValueError("goal configs have values for joints outside of the planner's planning joints "
                             "that don't match q_init")
        paths: list[list[np.ndarray]] = [[] for _ in range(B)]
        direct = np.linalg.norm(q_goals - q_inits, axis=1) <= self.epsilon
        for b in np.flatnonzero(direct):
            paths[b] = [q_inits[b].copy(), q_goals[b].copy()]
        ids = np.flatnonzero(~direct)
        S = len(ids)
        self.stats = {"iterations": 0, "ticks": 0, "configs_checked": 0, "launches": 0, "driver": "device ticks", "slots": S,
                      "gave_up": 0, "host_syncs": 0, "appends_refused_at_capacity": 0}
        t0 = time.time()
        if S == 0:
            self.stats.update(solved=int(direct.sum()), seconds=0.0)
            return paths
        eng.reset_stats()
        with torch.cuda.device(eng.device), eng._call_lock:
            QI = torch.from_numpy(np.ascontiguousarray(q_inits[ids])).to(dev)
            QG = torch.from_numpy(np.ascontiguousarray(q_goals[ids])).to(dev)
            cap = 2048
            plan_mask = torch.zeros(nq, dtype=torch.uint8, device=dev)
            plan_mask[q_idx] = 1
            lo = torch.from_numpy(np.ascontiguousarray(self.model.jnt_range[:, 0], dtype=np.float64)).to(dev)
            hi = torch.from_numpy(np.ascontiguousarray(self.model.jnt_range[:, 1], dtype=np.float64)).to(dev)
            T = {}

            def alloc(cap):
                old = dict(T)
                for k, roots in ((0, QI), (1, QG)):
                    T[f"nodes{k}"] = torch.empty((S, cap, nq), dtype=f64, device=dev)
                    T[f"parent{k}"] = torch.full((S, cap), -1, dtype=i64, device=dev)
                    if old:
                        oc = old[f"nodes{k}"].shape[1]
                        T[f"nodes{k}"][:, :oc] = old[f"nodes{k}"]
                        T[f"parent{k}"][:, :oc] = old[f"parent{k}"]
                    else:
                        T[f"nodes{k}"][:, 0] = roots
                        T[f"count{k}"] = torch.ones(S, dtype=i64, device=dev)

            alloc(cap)
            z = lambda *shape, dt=f64: torch.zeros(shape, dtype=dt, device=dev)
            bufs = {"phase": z(S, dt=torch.int32), "swapped": z(S, dt=torch.uint8), "age": z(S, dt=i64),
                    "target": z(S, nq), "tip": z(S, nq), "qa": z(S, nq), "last": z(S, dt=i64), "ia": z(S, dt=i64),
                    "cand": z(S, nq), "cand32": z(S, nq, dt=torch.float32), "proj": z(S, nq),
                    "proj_ok": z(S, dt=torch.uint8), "valid": z(S, dt=torch.uint8), "stepping": z(S, dt=torch.uint8),
                    "res_start": torch.full((S,), -1, dtype=i64, device=dev), "res_goal": torch.full((S,), -1, dtype=i64, device=dev),
                    "counters": z(8, dt=i64)}
            bufs["counters"][3] = S
            spec = pose._spec()

            def make_state(cap):
                st = _abi.CbirrtState()
                st.nslots, st.cap, st.nq, st.check_limits_before = S, cap, nq, int(limits_before)
                st.eps, st.goal_bias = float(self.epsilon), float(self.goal_biasing_probability)
                st.seed = int(self.seed if self.seed is not None else np.random.SeedSequence().entropy % (1 << 62))
                st.max_age = int(self.max_iterations_per_query)
                st.q_init, st.q_goal, st.plan_mask, st.lo, st.hi = (x.data_ptr() for x in (QI, QG, plan_mask, lo, hi))
                for k in (0, 1):
                    st.nodes[k], st.parent[k], st.count[k] = (T[f"{n}{k}"].data_ptr() for n in ("nodes", "parent", "count"))
                for name, tns in bufs.items():
                    setattr(st, name, tns.data_ptr())
                return st

            state = make_state(cap)

            def tick():
                _abi.check(L.mjb_cbirrt_tick(eng._h, C.byref(state), C.byref(spec), int(pose.max_iterations), flags,
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))

            ticks_per_sync = max(1, int(self.sync_every))
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                tick()             # eager once: the library grows its scratch here, never inside a capture
            side.synchronize()

            def capture():      # (capture_begin / capture_end directly, see _run_wave)
                g = torch.cuda.CUDAGraph()
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    g.capture_begin()
                    try:
                        for _ in range(ticks_per_sync):
                            tick()
                    finally:
                        g.capture_end()
                torch.cuda.current_stream().wait_stream(side)
                return g

            def try_capture():
                try:
                    return capture()
                except Exception:
                    return None

            graph = try_capture() if self.use_cuda_graph else None
            status = torch.empty(4, dtype=i64, device=dev)
            while True:
                status[0], status[1] = bufs["counters"][0], bufs["counters"][3]
                status[2], status[3] = T["count0"].max(), T["count1"].max()
                nt, n_alive, c0, c1 = (int(x) for x in status.cpu())
                self.stats["host_syncs"] += 1
                if n_alive == 0 or time.time() - t0 >= self.max_planning_time:
                    break
                if max(c0, c1) + ticks_per_sync + 1 > cap:   # a tree can gain one node per tick
                    cap *= 2
                    alloc(cap)
                    state = make_state(cap)
                    graph = try_capture() if graph is not None else None
                if graph is not None:
                    graph.replay()
                else:
                    with torch.cuda.stream(side):
                        for _ in range(ticks_per_sync):
                            tick()
                    side.synchronize()
            torch.cuda.current_stream().wait_stream(side)
            cnt = bufs["counters"].cpu().numpy()
            self.stats.update(ticks=int(cnt[0]), gave_up=int(cnt[2]), appends_refused_at_capacity=int(cnt[5]),
                              iterations=int(bufs["age"].max()))
            res_s, res_g = bufs["res_start"], bufs["res_goal"]
            solved = (res_s >= 0).nonzero(as_tuple=True)[0]
            if len(solved):
                depth = max(int(T["count0"].max()), int(T["count1"].max()), 1)
                found = _paths_of_solved(L, lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream), solved,
                                         T["nodes0"], T["parent0"], res_s[solved], T["nodes1"], T["parent1"], res_g[solved], depth)
                for s_, path in zip(solved.cpu().numpy(), found):
                    paths[int(ids[s_])] = path
        self.stats["solved"] = int(sum(1 for p in paths if p))
        self.stats["seconds"] = time.time() - t0
        est = eng.stats()
        self.stats["configs_checked"] = est["rows"]
        return paths

    # ------------------------------------------------------------------ host driver
    def _plan_host(self, q_inits, q_goals):
        B, nq = q_inits.shape
        ok = np.asarray(obeys_constraints_batch(np.concatenate([q_inits, q_goals]), self.constraints)).astype(bool)
        if not ok[:B].all():
            raise ValueError("q_init is not a valid configuration")
        if not ok[B:].all():
            bad = q_goals[np.flatnonzero(~ok[B:])[0]]
            raise ValueError(f"The following goal config is not a valid configuration: {bad}")
        q_idx = np.array(qpos_idx(self.model, self.planning_joints))
        fixed = np.array([i for i in range(nq) if i not in set(q_idx.tolist())], dtype=np.int64)
        if len(fixed) and not np.allclose(q_inits[:, fixed], q_goals[:, fixed], rtol=0, atol=1e-12):
            raise ValueError("goal configs have values for joints outside of the planner's planning joints "
                             "that don't match q_init")
        projecting = any(getattr(c, "projects", True) for c in self.constraints)
        paths: list[list[np.ndarray]] = [[] for _ in range(B)]
        direct = np.linalg.norm(q_goals - q_inits, axis=1) <= self.epsilon
        for b in np.flatnonzero(direct):
            paths[b] = [q_inits[b].copy(), q_goals[b].copy()]
        start, goal = _Forest(q_inits), _Forest(q_goals)
        active = np.flatnonzero(~direct)
        base = 0 if self.seed is None else int(self.seed)
        rngs = [np.random.default_rng(None if self.seed is None else base + b) for b in range(B)]
        lo, hi = self.model.jnt_range.T
        swapped = False
        self.stats = {"iterations": 0, "configs_checked": 0, "launches": 0}
        t0 = time.time()
        it = 0
        while len(active) and it < self.max_iterations and time.time() - t0 < self.max_planning_time:
            it += 1
            fa, fb = (goal, start) if swapped else (start, goal)
            # ---- sample (one stream per query, the reference's draw order) -------------------
            targets = np.empty((len(active), nq))
            for i, b in enumerate(active):
                r = rngs[b]
                if r.random() <= self.goal_biasing_probability:
                    if swapped:
                        targets[i] = q_inits[b]
                    else:
                        r.integers(0, 1)  # the reference draws a goal index even with one goal
                        targets[i] = q_goals[b]
                else:
                    t = q_inits[b].copy()
                    t[q_idx] = r.uniform(lo, hi)[q_idx]
                    targets[i] = t
            # ---- extend A towards the samples, then B towards what A reached (connect) ------
            extend = self._extend_projected if projecting else self._extend
            qa, ia = extend(fa, active, targets)
            qb, ib = extend(fb, active, qa)
            met = np.all(qa == qb, axis=1)
            for i in np.flatnonzero(met):
                b = int(active[i])
                s_idx, g_idx = (ib[i], ia[i]) if swapped else (ia[i], ib[i])
                ps = start.path_to_root(b, int(s_idx))[::-1]
                pg = goal.path_to_root(b, int(g_idx))
                if np.array_equal(ps[-1], pg[0]):
                    ps.pop()
                paths[b] = ps + pg
            active = active[~met]
            swapped = not swapped
        self.stats["iterations"] = it
        self.stats["solved"] = int(sum(1 for p in paths if p))
        self.stats["seconds"] = time.time() - t0
        return paths
