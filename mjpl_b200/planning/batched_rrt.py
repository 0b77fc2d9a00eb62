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


class BatchedRRT:
    """Lock-step bi-RRT over many (q_init, q_goal) queries.

    With non-projecting constraints whole extend chains are validated at once (device driver when
    the constraints fuse into the engine).  With a projecting constraint (``PoseConstraint``) the
    reference's step-by-step extend (``planning/utils.py:139-164``) is kept, but every step is
    taken by all queries together: one projection block and one validity block per step."""

    def __init__(self, model, planning_joints: list[str], constraints: list[Constraint],
                 max_planning_time: float = 10.0, epsilon: float = 0.05, seed: int | None = None,
                 goal_biasing_probability: float = 0.05, max_iterations: int = 1000000,
                 max_chain: int = 512, max_active: int = 4096, max_iterations_per_query: int = 3000) -> None:
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
        return self._plan_host(q_inits, q_goals)

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
        import ctypes as C

        import torch

        from .. import _abi

        dev, f64 = eng.torch_device, torch.float64
        L = _abi.lib()
        B, nq = q_inits.shape
        eps = float(self.epsilon)
        with torch.cuda.device(eng.device):
            QI = torch.from_numpy(q_inits).to(dev)
            QG = torch.from_numpy(q_goals).to(dev)
            ok = eng.valid_configs(torch.cat([QI, QG]).float(), flags)
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
            plan_mask = torch.zeros(nq, dtype=torch.bool, device=dev)
            plan_mask[q_idx] = True
            lo = torch.from_numpy(np.ascontiguousarray(self.model.jnt_range[:, 0])).to(dev)
            hi = torch.from_numpy(np.ascontiguousarray(self.model.jnt_range[:, 1])).to(dev)
            gen = torch.Generator(device=dev)
            if self.seed is not None:
                gen.manual_seed(int(self.seed))

            class Forest:
                def __init__(self, roots, cap=1024):
                    self.cap = cap
                    self.q = torch.full((B, cap, nq), float("inf"), dtype=f64, device=dev)
                    self.parent = torch.full((B, cap), -1, dtype=torch.int64, device=dev)
                    self.count = torch.ones(B, dtype=torch.int64, device=dev)
                    self.q[:, 0] = roots
                    self.hi = 1  # host-side upper bound of count.max()

                def reserve(self, extra):
                    if self.hi + extra <= self.cap:
                        return
                    new = max(self.hi + extra, 2 * self.cap)
                    q = torch.full((B, new, nq), float("inf"), dtype=f64, device=dev)
                    q[:, : self.cap] = self.q
                    p = torch.full((B, new), -1, dtype=torch.int64, device=dev)
                    p[:, : self.cap] = self.parent
                    self.q, self.parent, self.cap = q, p, new

                def nearest(self, act, targets):
                    out = torch.empty(len(act), dtype=torch.int64, device=dev)
                    t = targets.contiguous()
                    _abi.check(L.mjb_nearest_batch(self.q.data_ptr(), self.cap, nq, self.count.data_ptr(), act.data_ptr(),
                                                   t.data_ptr(), len(act), out.data_ptr(),
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
                    return out

            kcap = int(self.max_chain)
            eng.reset_stats()

            def extend(F, act, targets):
                """One ``_constrained_extend`` per active slot, entirely on the device
                (``mjb_rrt_extend``: nearest node, chain, validity, stop rules, append)."""
                n = len(act)
                F.reserve(kcap)
                reached = torch.empty((n, nq), dtype=f64, device=dev)
                last = torch.empty(n, dtype=torch.int64, device=dev)
                t = targets.contiguous()
                _abi.check(L.mjb_rrt_extend(eng._h, F.q.data_ptr(), F.parent.data_ptr(), F.count.data_ptr(), F.cap,
                                            act.data_ptr(), t.data_ptr(), n, eps, kcap, flags, reached.data_ptr(),
                                            last.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
                F.hi += kcap  # upper bound until the next exact refresh
                self.stats["launches"] += 1
                return reached, last

            direct = torch.linalg.vector_norm(QG - QI, dim=1) <= eps
            paths: list[list[np.ndarray]] = [[] for _ in range(B)]
            for b in direct.nonzero(as_tuple=True)[0].tolist():
                paths[b] = [q_inits[b].copy(), q_goals[b].copy()]
            pending = (~direct).nonzero(as_tuple=True)[0].cpu().numpy().tolist()[::-1]  # pop() takes the lowest id
            # ---- continuous batching: S slots, each holding one query's two trees; a slot whose
            # query is solved (or out of budget) is handed to the next pending query at once, so
            # the long tail of hard queries never idles the rest of the batch.
            S = min(len(pending), int(self.max_active)) if pending else 0
            self.stats = {"iterations": 0, "configs_checked": 0, "launches": 0, "driver": "device", "slots": S,
                          "gave_up": 0}
            t0 = time.time()
            it = 0
            if S:
                slot_query = torch.full((S,), -1, dtype=torch.int64, device=dev)
                slot_age = torch.zeros(S, dtype=torch.int64, device=dev)
                first = [pending.pop() for _ in range(S)]
                slot_query[:] = torch.tensor(first, device=dev)
                SQI, SQG = QI[slot_query].clone(), QG[slot_query].clone()   # per-slot endpoints
                B_forest = S

                class SlotForest(Forest):
                    def __init__(self, roots):
                        self.cap = 1024
                        self.q = torch.full((B_forest, self.cap, nq), float("inf"), dtype=f64, device=dev)
                        self.parent = torch.full((B_forest, self.cap), -1, dtype=torch.int64, device=dev)
                        self.count = torch.ones(B_forest, dtype=torch.int64, device=dev)
                        self.q[:, 0] = roots
                        self.hi = 1

                    def reserve(self, extra):
                        if self.hi + extra <= self.cap:
                            return
                        new = max(self.hi + extra, 2 * self.cap)
                        q = torch.full((B_forest, new, nq), float("inf"), dtype=f64, device=dev)
                        q[:, : self.cap] = self.q
                        p = torch.full((B_forest, new), -1, dtype=torch.int64, device=dev)
                        p[:, : self.cap] = self.parent
                        self.q, self.parent, self.cap = q, p, new

                start, goal = SlotForest(SQI), SlotForest(SQG)
                live = torch.arange(S, device=dev)       # slots that hold a query
                swapped = False
                while len(live) and it < self.max_iterations and time.time() - t0 < self.max_planning_time:
                    it += 1
                    n = len(live)
                    fa, fb = (goal, start) if swapped else (start, goal)
                    u = torch.rand(n, generator=gen, device=dev, dtype=f64)
                    rnd = lo[None, :] + (hi - lo)[None, :] * torch.rand((n, nq), generator=gen, device=dev, dtype=f64)
                    qi_a = SQI[live]
                    targets = torch.where(plan_mask[None, :], rnd, qi_a)
                    bias = (u <= self.goal_biasing_probability)[:, None]
                    targets = torch.where(bias, qi_a if swapped else SQG[live], targets)
                    qa, ia = extend(fa, live, targets)
                    qb, ib = extend(fb, live, qa)
                    slot_age[live] += 1
                    met = (qa == qb).all(dim=1)
                    done = met | (slot_age[live] >= self.max_iterations_per_query)
                    # the only host read-back of the iteration: any slot finished? + exact tree sizes
                    info = torch.stack([done.any().to(torch.int64), start.count.max(), goal.count.max()]).cpu()
                    start.hi, goal.hi = int(info[1]), int(info[2])
                    if bool(info[0]):
                        di = done.nonzero(as_tuple=True)[0]
                        dslots = live[di]
                        dmet = met[di].cpu().numpy()
                        dq = slot_query[dslots].cpu().numpy()
                        s_idx, g_idx = (ib, ia) if swapped else (ia, ib)
                        ds, dg = s_idx[di].cpu().numpy(), g_idx[di].cpu().numpy()
                        hi_s = int(start.count[dslots].max())
                        hi_g = int(goal.count[dslots].max())
                        sp = start.parent[dslots, :hi_s].cpu().numpy()
                        gp = goal.parent[dslots, :hi_g].cpu().numpy()
                        want_s, want_g, who = [], [], []
                        for k in range(len(di)):
                            if not dmet[k]:
                                self.stats["gave_up"] += 1
                                continue
                            a, si = [], int(ds[k])
                            while si >= 0:
                                a.append(si)
                                si = int(sp[k, si])
                            g, gi = [], int(dg[k])
                            while gi >= 0:
                                g.append(gi)
                                gi = int(gp[k, gi])
                            want_s.append((int(dslots[k]), a[::-1]))
                            want_g.append((int(dslots[k]), g))
                            who.append(int(dq[k]))

                        def gather(F, lists):
                            if not lists:
                                return []
                            bb = np.concatenate([np.full(len(ix), sl) for sl, ix in lists])
                            ii = np.concatenate([np.asarray(ix) for _, ix in lists])
                            rows = F.q[torch.from_numpy(bb).to(dev), torch.from_numpy(ii).to(dev)].cpu().numpy()
                            out, o = [], 0
                            for _, ix in lists:
                                out.append(rows[o : o + len(ix)])
                                o += len(ix)
                            return out

                        for qid, ps, pg in zip(who, gather(start, want_s), gather(goal, want_g)):
                            ps, pg = list(ps), list(pg)
                            if np.array_equal(ps[-1], pg[0]):
                                ps.pop()
                            paths[qid] = ps + pg
                        # hand the freed slots to pending queries (or retire them)
                        refill = [pending.pop() for _ in range(min(len(pending), len(dslots)))]
                        nr = len(refill)
                        if nr:
                            rs = dslots[:nr]
                            rq = torch.tensor(refill, device=dev)
                            slot_query[rs] = rq
                            SQI[rs], SQG[rs] = QI[rq], QG[rq]
                            for F, roots in ((start, QI[rq]), (goal, QG[rq])):
                                F.q[rs, 0] = roots
                                F.count[rs] = 1
                                F.parent[rs, 0] = -1
                            slot_age[rs] = 0
                        if nr < len(dslots):
                            keep = torch.ones(n, dtype=torch.bool, device=dev)
                            keep[di[nr:]] = False
                            live = live[keep]
                    swapped = not swapped
            self.stats["iterations"] = it
            self.stats["solved"] = int(sum(1 for p in paths if p))
            self.stats["seconds"] = time.time() - t0
            est = eng.stats()
            self.stats["configs_checked"] = est["rows"]
            self.stats["chains_clipped_at_capacity"] = est["queue_overflow"]
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
