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


def _extend_blocks(starts: list[np.ndarray], targets: list[np.ndarray], eps: float, constraints: list[Constraint],
                   collision_interval_check, equality_threshold: float = 1e-8) -> list[np.ndarray]:
    """``_extend_block`` for many independent (start, target) pairs at once: every chain is generated
    up front, ALL rows of ALL chains are validated by one ``obeys_constraints_batch`` call (one fused
    kernel launch), and the stop rules of the reference (``planning/utils.py:151-160``) cut each chain
    at its first failing row.  Interval checks, when asked for, are one more batched call over the
    edges of the surviving prefixes."""
    chains = [_chain(a, b, eps) for a, b in zip(starts, targets)]
    sizes = [len(c) for c in chains]
    nq = len(np.asarray(starts[0]))
    flat = np.concatenate(chains, axis=0) if sum(sizes) else np.empty((0, nq))
    ok_flat = np.asarray(obeys_constraints_batch(flat, constraints)).astype(bool) if len(flat) else np.zeros(0, dtype=bool)
    out, prevs, keep, o = [], [], [], 0
    for a, c, k in zip(starts, chains, sizes):
        ok = ok_flat[o:o + k].copy()
        o += k
        prev = np.concatenate([np.asarray(a, dtype=np.float64)[None, :], c[:-1]], axis=0) if k else c
        if k:
            ok &= np.linalg.norm(c - prev, axis=1) >= equality_threshold
        n_ok = k if ok.all() else int(np.argmin(ok))
        prevs.append(prev)
        keep.append(n_ok)
    if collision_interval_check is not None and sum(keep):
        step_dist, cc = collision_interval_check
        e0 = np.concatenate([p[:n] for p, n in zip(prevs, keep)], axis=0)
        e1 = np.concatenate([c[:n] for c, n in zip(chains, keep)], axis=0)
        iv = _valid_intervals(e0, e1, step_dist, cc)
        o = 0
        for i, n in enumerate(keep):
            seg = iv[o:o + n]
            o += n
            if n and not seg.all():
                keep[i] = int(np.argmin(seg))
    for c, n in zip(chains, keep):
        out.append(c[:n])
    return out


def smooth_path(waypoints: list[np.ndarray], constraints: list[Constraint],
                collision_interval_check: tuple[float, CollisionConstraint] | None = None,
                eps: float = 0.05, num_tries: int = 100, seed: int | None = None,
                sparse: bool = False) -> list[np.ndarray]:
    """Shortcut smoothing (CBiRRT algorithm 3; reference ``src/mjpl/planning/utils.py:9-87``).

    The reference runs ``num_tries`` tries one after the other: two ``rng.integers`` draws pick a
    sub-path, a constrained extend tries to connect its ends directly, and the sub-path is replaced
    when the connection exists and is shorter.  A try changes nothing unless it is accepted, so here
    ALL remaining tries are evaluated speculatively against the current path -- their draws are taken
    from the reference's random stream in order, and every candidate connection goes into ONE batched
    validity call -- and the first try that the reference would accept is applied; the random stream
    is rewound to just after that try and the rest is speculated again on the new path.  The result is
    the reference's, waypoint for waypoint, for the same seed.  Every accepted shortcut ends a round
    (the draws after it depend on the new path), so the number of validity launches is the number of
    accepted shortcuts plus the rounds without one; the speculation depth adapts (4 tries, doubling
    after a round without an accepted shortcut, halving after one with).  A projecting constraint makes a
    connection depend on its own intermediate results, so those are evaluated one try at a time."""
    if not waypoints:
        raise ValueError("`waypoints` cannot be empty.")
    if eps <= 0.0:
        raise ValueError("`eps` must be > 0.")
    if num_tries <= 0:
        raise ValueError("`num_tries` must be > 0.")

    projecting = any(getattr(c, "projects", True) for c in constraints)
    smoothed = list(waypoints)
    rng = np.random.default_rng(seed=seed)
    done = 0
    smooth_path.last_launches = 0    # introspection for tests / benches: batched validity rounds of the last call
    depth = 4                        # tries speculated per round: doubles after a round without an accepted
    while done < num_tries:          # shortcut, halves after one with (rows past an accepted try are wasted work)
        # speculate: the draws of the next tries, assuming none before them is accepted
        lookahead = 1 if projecting else min(depth, num_tries - done)
        draws, states = [], []
        for _ in range(lookahead):
            start = int(rng.integers(0, len(smoothed) - 1))
            end = int(rng.integers(start + 1, len(smoothed)))
            draws.append((start, end))
            states.append(rng.bit_generator.state)
        if projecting:
            start, end = draws[0]
            tree = Tree(Node(smoothed[start]))
            reached = _constrained_extend(smoothed[end], tree, eps, constraints, collision_interval_check)
            segs = [None]
            if np.array_equal(reached, smoothed[end]):
                seg = [n.q for n in tree.get_path(tree.nearest_neighbor(reached))]
                seg.reverse()               # get_path runs node -> root
                segs = [seg]
        else:
            blocks = _extend_blocks([smoothed[s] for s, _ in draws], [smoothed[e] for _, e in draws], eps, constraints,
                                    collision_interval_check)
            segs = []
            for (s, e), rows in zip(draws, blocks):
                tgt = np.asarray(smoothed[e], dtype=np.float64)
                if np.array_equal(smoothed[s], tgt):
                    segs.append([smoothed[s]])            # the extend returns at once: already there
                elif len(rows) and np.array_equal(rows[-1], tgt):
                    segs.append([smoothed[s]] + list(rows))
                else:
                    segs.append(None)
        smooth_path.last_launches += 1
        accepted = None
        for k, ((start, end), seg) in enumerate(zip(draws, segs)):
            # (the reference measures the new segment in tree order, end -> start: planning/utils.py:70-73;
            #  summed in that order so that a tie in exact arithmetic rounds the way it does there)
            if seg is not None and path_length(seg[::-1]) < path_length(smoothed[start:end + 1]):
                accepted = k
                break
        if accepted is None:
            done += lookahead
            depth = min(2 * depth, 64)
            continue
        depth = max(depth // 2, 2)
        start, end = draws[accepted]
        seg = segs[accepted]
        if sparse:
            smoothed = smoothed[:start + 1] + smoothed[end:]
        else:
            smoothed = smoothed[:start] + seg[:-1] + smoothed[end:]
        rng.bit_generator.state = states[accepted]     # the stream continues right after the accepted try
        done += accepted + 1
    return smoothed


def _combine_paths(start_tree: Tree, start_tree_node: Node, goal_tree: Tree, goal_tree_node: Node) -> list[np.ndarray]:
    """Root of ``start_tree`` -> ``start_tree_node`` -> ``goal_tree_node`` -> root of ``goal_tree``."""
    path_start = [n.q for n in start_tree.get_path(start_tree_node)]
    path_start.reverse()
    path_end = [n.q for n in goal_tree.get_path(goal_tree_node)]
    if np.array_equal(path_start[-1], path_end[0]):
        path_start.pop()
    return path_start + path_end
