"""CBiRRT with block-evaluated extends.

Reference: ``src/mjpl/planning/rrt.py`` (constructor checks :51-58, ``plan_to_configs``
:141-237).  The control flow, argument validation, error messages and random stream
(``rng.random()``, then ``rng.integers`` or ``rng.uniform(*jnt_range.T)`` over all joints) are
the reference's; each ``_constrained_extend`` call validates its whole chain in one kernel
launch (see ``planning/utils.py``).  ``plan_to_pose(s)`` needs an IK solver object
(``IKSolver.solve_ik``); the reference's default (mink + daqp QP) is not available here, so the
solver argument is mandatory -- SURVEY.md section 8(f) lists a native IK as a later row.
"""

from __future__ import annotations

import time

import numpy as np

from ..constraint.collision_constraint import CollisionConstraint
from ..constraint.constraint_interface import Constraint
from ..constraint.utils import obeys_constraints, obeys_constraints_batch
from ..utils import qpos_idx
from .tree import Node, Tree
from .utils import _combine_paths, _constrained_extend


class RRT:
    """CBiRRT: bi-directional RRT with support for constraints."""

    def __init__(self, model, planning_joints: list[str], constraints: list[Constraint],
                 collision_interval_check: tuple[float, CollisionConstraint] | None = None,
                 max_planning_time: float = 10.0, epsilon: float = 0.05, seed: int | None = None,
                 goal_biasing_probability: float = 0.05) -> None:
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
        self.collision_interval_check = collision_interval_check
        self.max_planning_time = max_planning_time
        self.epsilon = epsilon
        self.seed = seed
        self.goal_biasing_probability = goal_biasing_probability

    # ---- pose goals: IK, then plan to the configurations (reference rrt.py:69-139) ----------------
    def plan_to_pose(self, q_init, pose, site: str, solver=None) -> list[np.ndarray]:
        return self.plan_to_poses(q_init, [pose], site, solver)

    def plan_to_poses(self, q_init, poses, site: str, solver=None) -> list[np.ndarray]:
        if solver is None:
            from ..inverse_kinematics import DLSIKSolver

            solver = DLSIKSolver(model=self.model, joints=self.planning_joints, constraints=self.constraints,
                                 seed=self.seed, max_attempts=5)
        if hasattr(solver, "solve_ik_batch"):
            Q, solved = solver.solve_ik_batch(list(poses), site, np.asarray(q_init, dtype=np.float64))
            cands = [q for q, k in zip(Q, solved) if k]
        else:
            cands = [q for p in poses for q in solver.solve_ik(p, site, q_init_guess=q_init)]
        if not cands:
            return []
        ok = np.asarray(obeys_constraints_batch(np.asarray(cands, dtype=np.float64), self.constraints))
        configs = [q for q, k in zip(cands, ok) if k]
        return [] if not configs else self.plan_to_configs(q_init, configs)

    def plan_to_config(self, q_init: np.ndarray, q_goal: np.ndarray) -> list[np.ndarray]:
        return self.plan_to_configs(q_init, [q_goal])

    # ---- validation shared with the batched planner ---------------------------------------------
    def _validate(self, q_init, q_goals):
        if not obeys_constraints(q_init, self.constraints):
            raise ValueError("q_init is not a valid configuration")
        for q in q_goals:
            if not obeys_constraints(q, self.constraints):
                raise ValueError(f"The following goal config is not a valid configuration: {q}")
        q_idx = qpos_idx(self.model, self.planning_joints)
        fixed = [i for i in range(self.model.nq) if i not in q_idx]
        for q in q_goals:
            if not np.allclose(q_init[fixed], q[fixed], rtol=0, atol=1e-12):
                raise ValueError(
                    f"The following goal config has values for joints outside of "
                    f"the planner's planning joints that don't match q_init: {q}. "
                    f"q_init is {q_init}, and the planning joints are {self.planning_joints}")
        return q_idx

    def plan_to_configs(self, q_init: np.ndarray, q_goals: list[np.ndarray]) -> list[np.ndarray]:
        """Path from ``q_init`` to one of ``q_goals`` (empty list on timeout)."""
        q_idx = self._validate(q_init, q_goals)
        for q in q_goals:  # direct connection?
            if np.linalg.norm(q - q_init) <= self.epsilon:
                return [q_init, q]

        start_tree = Tree(Node(q_init))
        # goal tree: a sink root at +inf (never the nearest neighbour) with the goals as children
        sink = Node(np.ones_like(q_init) * np.inf)
        goal_nodes = [Node(q, sink) for q in q_goals]
        goal_tree = Tree(sink)
        for n in goal_nodes:
            goal_tree.add_node(n)

        rng = np.random.default_rng(seed=self.seed)
        lo, hi = self.model.jnt_range.T
        tree_a, tree_b = start_tree, goal_tree
        swapped = False
        t0 = time.time()
        while time.time() - t0 < self.max_planning_time:
            if rng.random() <= self.goal_biasing_probability:
                q_rand = q_init if swapped else goal_nodes[rng.integers(0, len(goal_nodes))].q
            else:
                q_rand = q_init.copy()
                q_rand[q_idx] = rng.uniform(lo, hi)[q_idx]
            q_a = _constrained_extend(q_rand, tree_a, self.epsilon, self.constraints, self.collision_interval_check)
            q_b = _constrained_extend(q_a, tree_b, self.epsilon, self.constraints, self.collision_interval_check)
            if np.array_equal(q_a, q_b):
                waypoints = _combine_paths(start_tree, start_tree.nearest_neighbor(q_a),
                                           goal_tree, goal_tree.nearest_neighbor(q_a))
                return waypoints[:-1]  # drop the sink
            tree_a, tree_b = tree_b, tree_a
            swapped = not swapped
        return []
