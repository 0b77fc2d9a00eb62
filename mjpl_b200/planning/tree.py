"""Search tree with array-backed storage.

Same observable behaviour as the reference's ``Node`` / ``Tree``
(``src/mjpl/planning/tree.py``): nodes are equal iff their ``q`` bytes are equal (:15-23),
``add_node`` rejects duplicates and orphans (:37-50), ``nearest_neighbor`` minimises the
Euclidean distance (:57-66), ``get_path`` walks to the root (:68-85).  The difference is the
representation: configurations live in one growing ``(capacity, nq)`` matrix, so the
nearest-neighbour query is one vectorised reduction instead of a Python loop with one
``np.linalg.norm`` per node, and whole extend chains are appended in one call.
"""

from __future__ import annotations

import numpy as np


class Node:
    """Tree node: a configuration and a parent link.  Hash / equality use ``q`` only."""

    __slots__ = ("q", "parent", "_key")

    def __init__(self, q: np.ndarray, parent: "Node | None" = None):
        self.q = q
        self.parent = parent
        self._key = np.ascontiguousarray(q).tobytes()

    def __hash__(self):
        return hash(self._key)

    def __eq__(self, other):
        if not isinstance(other, Node):
            return False
        return np.array_equal(self.q, other.q)

    def __repr__(self):
        return f"Node(q={self.q!r})"


class Tree:
    """Tree of nodes."""

    def __init__(self, root: Node):
        if root.parent:
            raise ValueError("The root node should have no parent.")
        nq = int(np.asarray(root.q).shape[0])
        self._q = np.empty((64, nq), dtype=np.float64)
        self._nodes: list[Node] = []
        self._index: dict[bytes, int] = {}
        self._append(root)

    # -- storage ------------------------------------------------------------------------------
    def _append(self, node: Node) -> None:
        n = len(self._nodes)
        if n == len(self._q):
            self._q = np.concatenate([self._q, np.empty_like(self._q)], axis=0)
        self._q[n] = node.q
        self._nodes.append(node)
        self._index[node._key] = n

    @property
    def nodes(self) -> set:
        """The nodes as a set (reference attribute)."""
        return set(self._nodes)

    def __len__(self) -> int:
        return len(self._nodes)

    def __contains__(self, node: Node) -> bool:
        return node._key in self._index

    # -- reference API --------------------------------------------------------------------------
    def add_node(self, node: Node) -> None:
        if not node.parent:
            raise ValueError("Node does not have a parent.")
        if node in self:
            raise ValueError(f"A node with q={node.q} already exists in the tree.")
        if node.parent not in self:
            raise ValueError("Node's parent is not in the tree.")
        self._append(node)

    def add_chain(self, parent: Node, Q: np.ndarray) -> Node:
        """Append rows of ``Q`` as a chain hanging off ``parent``; returns the last node.
        Rows that already exist in the tree end the chain there (the reference would raise
        from ``add_node`` on the same input)."""
        last = parent
        for q in Q:
            node = Node(q, last)
            if node in self:
                break
            self._append(node)
            last = node
        return last

    def nearest_neighbor(self, q: np.ndarray) -> Node:
        n = len(self._nodes)
        with np.errstate(invalid="ignore", over="ignore"):
            d = self._q[:n] - np.asarray(q, dtype=np.float64)
            d2 = np.einsum("ij,ij->i", d, d)
        d2 = np.where(np.isnan(d2), np.inf, d2)  # the +inf sink root (rrt.py:184-188 in the reference)
        return self._nodes[int(np.argmin(d2))]

    def get_path(self, node: Node) -> list[Node]:
        if node not in self:
            raise ValueError("Node is not in the tree.")
        path = []
        cur = node
        while cur is not None:
            path.append(cur)
            cur = cur.parent
        return path
