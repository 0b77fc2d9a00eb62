"""CBiRRT building blocks re-expressed to feed the validity engine whole blocks of q.

Reference: ``src/mjpl/planning/utils.py`` -- ``smooth_path`` :9-87, ``path_length`` :90-102,
``_constrained_extend`` :105-164, ``_step`` :167-185, ``_valid_collision_interval`` :188-216,
``_combine_paths`` :219-249.

What changes and what does not
------------------------------
* With non-projecting constraints (``apply`` returns ``q`` or ``None``) the extend chain
  ``q_k`` is a deterministic function of (nearest node, target, eps).  The reference evaluates
  it one configuration at a time, running FK + collision twice per step (``apply`` and the
  re-validation in ``apply_constraints``); here the whole chain is generated up front,
  validated in ONE fused kernel launch and truncated at the first failing index.  The four stop
  rules of the reference (:151-160) are applied to the same quantities.
* Deliberate deviation: when the remaining distance is <= eps the step lands
  on ``target`` itself instead of ``start + unit * magnitude``, which in the reference can be one
  ulp off and then fails the exact-equality connection tests (``rrt.py:223``,
  ``planning/utils.py:64``).  Every other waypoint differs from the reference's accumulated
  sum by rounding only.
* Projecting constraints (``Constraint.projects``) keep the reference's sequential algorithm.
"""

from __future__ import annotations

import numpy as np

from ..constraint.collision_constraint import CollisionConstraint
from ..constraint.constraint_interface import Constraint
from ..constraint.utils import apply_constraints, obeys_constraints_batch
from .tree import Node, Tree


def path_length(waypoints: list[np.ndarray]) -> float:
    """Length of a waypoint list in configuration space."""
    path = np.asarray(waypoints, dtype=np.float64)
    if len(path) < 2:
        return 0.0
    return float(np.sum(np.linalg.norm(np.diff(path, axis=0), axis=1)))


def _step(start: np.ndarray, target: np.ndarray, max_step_dist: float) -> np.ndarray:
    """One step of at most ``max_step_dist`` from ``start`` towards ``target``."""
    if max_step_dist <= 0.0:
        raise ValueError("`max_step_dist` must be > 0.0")
    if np.array_equal(start, target):
        return start.copy()
    direction = target - start
    magnitude = np.linalg.norm(direction)
    if magnitude <= max_step_dist:
        return np.array(target, dtype=np.float64, copy=True)
    return start + direction * (max_step_dist / magnitude)


def _chain(start: np.ndarray, target: np.ndarray, eps: float) -> np.ndarray:
    """All configurations ``_step`` would visit from ``start`` to ``target`` (excluding start,
    including target), as one ``(K, nq)`` block."""
    start = np.asarray(start, dtype=np.float64)
    target = np.asarray(target, dtype=np.float64)
    d = target - start
    dist = float(np.linalg.norm(d))
    if dist == 0.0:
        return np.empty((0, len(start)))
    if not np.isfinite(eps) or dist <= eps:
        return target[None, :].copy()
    k = int(np.ceil(dist / eps))
    t = (np.arange(1, k, dtype=np.float64) * (eps / dist))[:, None]
    return np.concatenate([start[None, :] + t * d[None, :], target[None, :]], axis=0)


def _valid_collision_interval(start: np.ndarray, end: np.ndarray, step_dist: float,
                              constraint: CollisionConstraint) -> bool:
    """Do the configurations strictly between ``start`` and ``end`` (every ``step_dist``) obey
    ``constraint``?  One edge through the batched edge kernel."""
    if step_dist <= 0.0:
        raise ValueError("`step_dist` must be > 0")
    if hasattr(constraint, "valid_edges"):
        ok = constraint.valid_edges(np.asarray(start, dtype=np.float64)[None, :],
                                    np.asarray(end, dtype=np.float64)[None, :], step_dist)
        return bool(np.asarray(ok)[0])
    wps = _chain(start, end, step_dist)[:-1]
    return bool(np.all(constraint.valid_configs(wps))) if len(wps) else True


def _valid_intervals(starts: np.ndarray, ends: np.ndarray, step_dist: float,
                     constraint: CollisionConstraint) -> np.ndarray:
    if step_dist <= 0.0:
        raise ValueError("`step_dist` must be > 0")
    if hasattr(constraint, "valid_edges"):
        return np.asarray(constraint.valid_edges(starts, ends, step_dist))
    return np.array([_valid_collision_interval(a, b, step_dist, constraint) for a, b in zip(starts, ends)])


def _constrained_extend_sequential(q_target, tree, eps, constraints, collision_interval_check, equality_threshold):
    """The reference's step-by-step algorithm (needed when a constraint projects)."""
    closest_node = tree.nearest_neighbor(q_target)
    q = closest_node.q
    q_old = closest_node.q
    while True:
        if np.array_equal(q_target, q):
            return q
        q = _step(q, q_target, eps)
        q = apply_constraints(q_old, q, constraints)
        if (
            q is None
            or np.linalg.norm(q - q_old) < equality_threshold
            or np.linalg.norm(q_target - q) > np.linalg.norm(q_target - q_old)
            or (collision_interval_check is not None
                and not _valid_collision_interval(q_old, q, *collision_interval_check))
        ):
            return q_old
        closest_node = Node(q, closest_node)
        tree.add_node(closest_node)
        q_old = q


def _extend_block(q_near: np.ndarray, q_target: np.ndarray, eps: float, constraints: list[Constraint],
                  collision_interval_check, equality_threshold: float) -> np.ndarray:
    """Rows of the extend chain from ``q_near`` that survive all stop rules, ``(M, nq)``."""
    chain = _chain(q_near, q_target, eps)
    if len(chain) == 0:
        return chain
    ok = np.asarray(obeys_constraints_batch(chain, constraints)).astype(bool)
    prev = np.concatenate([np.asarray(q_near, dtype=np.float64)[None, :], chain[:-1]], axis=0)
    # stop rule 2: the step did not move (only the final, possibly tiny, step can trip it)
    ok &= np.linalg.norm(chain - prev, axis=1) >= equality_threshold
    # stop rule 3 (moved away from the target) cannot fire on a straight chain
    n_ok = len(ok) if ok.all() else int(np.argmin(ok))
    if n_ok and collision_interval_check is not None:
        step_dist, cc = collision_interval_check
        iv = _valid_intervals(prev[:n_ok], chain[:n_ok], step_dist, cc)
        if not iv.all():
            n_ok = int(np.argmin(iv))
    return chain[:n_ok]


def _constrained_extend(q_target: np.ndarray, tree: Tree, eps: float, constraints: list[Constraint],
                        collision_interval_check: tuple[float, CollisionConstraint] | None = None,
                        equality_threshold: float = 1e-8) -> np.ndarray:
    """Extend ``tree`` towards ``q_target`` subject to ``constraints``; returns the
    configuration that was reached (CBiRRT algorithm 2)."""
    if eps <= 0.0:
        raise ValueError("`max_step_dist` must be > 0.0")
    if any(getattr(c, "projects", True) for c in constraints):
        return _constrained_extend_sequential(q_target, tree, eps, constraints, collision_interval_check,
                                              equality_threshold)
    closest = tree.nearest_neighbor(q_target)
    if np.array_equal(q_target, closest.q):
        return closest.q
    rows = _extend_block(closest.q, np.asarray(q_target, dtype=np.float64), eps, constraints,
                         collision_interval_check, equality_threshold)
    if len(rows) == 0:
        return closest.q
    last = tree.add_chain(closest, rows)
    return last.q


def smooth_path(waypoints: list[np.ndarray], constraints: list[Constraint],
                collision_interval_check: tuple[float, CollisionConstraint] | None = None,
                eps: float = 0.05, num_tries: int = 100, seed: int | None = None,
                sparse: bool = False) -> list[np.ndarray]:
    """Shortcut smoothing (CBiRRT algorithm 3): ``num_tries`` times pick two waypoints and
    replace the sub-path by a direct constrained connection when that is shorter.  The random
    stream (two ``rng.integers`` per try) is the reference's."""
    if not waypoints:
        raise ValueError("`waypoints` cannot be empty.")
    if eps <= 0.0:
        raise ValueError("`eps` must be > 0.")
    if num_tries <= 0:
        raise ValueError("`num_tries` must be > 0.")

    smoothed = waypoints
    rng = np.random.default_rng(seed=seed)
    for _ in range(num_tries):
        start = rng.integers(0, len(smoothed) - 1)
        end = rng.integers(start + 1, len(smoothed))
        tree = Tree(Node(smoothed[start]))
        q_reached = _constrained_extend(smoothed[end], tree, eps, constraints, collision_interval_check)
        if not np.array_equal(q_reached, smoothed[end]):
            continue
        end_node = tree.nearest_neighbor(q_reached)
        segment = [n.q for n in tree.get_path(end_node)]
        if path_length(segment) < path_length(smoothed[start : end + 1]):
            if sparse:
                smoothed = smoothed[: start + 1] + smoothed[end:]
            else:
                segment.reverse()  # get_path runs node -> root
                smoothed = smoothed[:start] + segment[:-1] + smoothed[end:]
    return smoothed


def _combine_paths(start_tree: Tree, start_tree_node: Node, goal_tree: Tree, goal_tree_node: Node) -> list[np.ndarray]:
    """Root of ``start_tree`` -> ``start_tree_node`` -> ``goal_tree_node`` -> root of ``goal_tree``."""
    path_start = [n.q for n in start_tree.get_path(start_tree_node)]
    path_start.reverse()
    path_end = [n.q for n in goal_tree.get_path(goal_tree_node)]
    if np.array_equal(path_start[-1], path_end[0]):
        path_start.pop()
    return path_start + path_end
