"""Utterance partitioning.  Requests are independent, so multi-GPU is a static partition of whole
utterances per rank and there is no collective on the data path."""
from __future__ import annotations

from typing import List, Sequence


def lpt_partition(costs: Sequence[float], n_parts: int) -> List[List[int]]:
    """Longest-processing-time-first: indices of ``costs`` dealt to ``n_parts`` bins, heaviest first into the
    lightest bin.  Deterministic (ties by index)."""
    if n_parts <= 0:
        raise ValueError("n_parts must be positive")
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * n_parts
    parts: List[List[int]] = [[] for _ in range(n_parts)]
    for i in order:
        j = min(range(n_parts), key=lambda p: (loads[p], p))
        parts[j].append(i)
        loads[j] += float(costs[i])
    for p in parts:
        p.sort()
    return parts


def micro_batches(costs: Sequence[float], budget: float, max_items: int = 1 << 30) -> List[List[int]]:
    """Greedy grouping of consecutive-by-cost utterances into micro-batches whose summed cost stays within ``budget``
    (a frame budget bounds the engine's workspace).  Items are sorted by cost so a micro-batch is length-homogeneous.
    An item larger than the budget gets a micro-batch of its own."""
    order = sorted(range(len(costs)), key=lambda i: (float(costs[i]), i))
    out: List[List[int]] = []
    cur: List[int] = []
    load = 0.0
    for i in order:
        c = float(costs[i])
        if cur and (load + c > budget or len(cur) >= max_items):
            out.append(cur)
            cur, load = [], 0.0
        cur.append(i)
        load += c
    if cur:
        out.append(cur)
    return out
