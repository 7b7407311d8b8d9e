"""Drop-in for the reference's ``dup.cluster`` (src/dup/cluster.py:12-70): connected components
over the matches flagged ``is_duplicate``.  Host code; the input is a handful of verified pairs."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable

from .refine import RefinedMatch


@dataclass
class Cluster:
    representative: int
    members: list[int]
    matches: list[RefinedMatch]


class ClusterBuilder:
    """Union-find where the smaller root wins; representative = smallest member; clusters sorted by it."""

    def build(self, matches: Iterable[RefinedMatch]) -> list[Cluster]:
        kept = [m for m in matches if m.is_duplicate]
        if not kept:
            return []
        parent: dict[int, int] = {}

        def root_of(x: int) -> int:
            parent.setdefault(x, x)
            r = x
            while parent[r] != r:
                r = parent[r]
            while parent[x] != r:
                parent[x], x = r, parent[x]
            return r

        for m in kept:
            ra, rb = root_of(m.file_id_a), root_of(m.file_id_b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)
        members: dict[int, list[int]] = {}
        for node in parent:
            members.setdefault(root_of(node), []).append(node)
        edges: dict[int, list[RefinedMatch]] = {}
        for m in kept:
            edges.setdefault(root_of(m.file_id_a), []).append(m)
        clusters = [Cluster(representative=min(ms), members=sorted(ms), matches=edges.get(r, []))
                    for r, ms in members.items()]
        clusters.sort(key=lambda c: c.representative)
        return clusters


__all__ = ["Cluster", "ClusterBuilder"]
