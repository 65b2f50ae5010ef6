"""Utterance sharding across the GPUs of one box: independent units, no collective.

Each rank (one process per GPU) holds a full weight replica and processes its own utterances; a
greedy longest-first assignment balances the summed length per rank.  Deterministic, so every rank
computes the same partition from the same length list without communicating.
"""
from __future__ import annotations

from typing import List, Sequence


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Return ``world_size`` lists of utterance indices (each sorted ascending).

    Greedy LPT: visit utterances by decreasing length (ties by index), give each to the currently
    lightest rank (ties by rank id).
    """
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in sorted(range(len(lengths)), key=lambda j: (-int(lengths[j]), j)):
        r = min(range(world_size), key=lambda q: (loads[q], q))
        shards[r].append(i)
        loads[r] += int(lengths[i])
    for s in shards:
        s.sort()
    return shards
