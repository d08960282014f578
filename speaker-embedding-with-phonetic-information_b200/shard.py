"""Utterance sharding across GPUs / ranks.

The reference shards extraction with `utils/split_data.sh $data $nj` (per speaker,
not length balanced) and one process per shard
(egs/sre/v5/sid/nnet3/xvector/extract_xvectors_new.sh:72-74,91-93); there is no
cross-utterance reduction, so there is no collective.  Here shards are balanced by
frame count with the greedy longest-processing-time rule, an utterance's chunks
stay on one rank (its weighted average is local), and the merge is a
concatenation of the per-rank outputs in input order (:97-100).
"""
from __future__ import annotations

import heapq
from typing import List, Sequence

import numpy as np


def lpt_shards(lengths: Sequence[int], num_shards: int) -> List[np.ndarray]:
    """Greedy LPT: utterances sorted by length (desc, ties by index) go to the
    currently lightest shard.  Returns, per shard, the utterance indices in
    ascending (input) order.  Deterministic, so every rank computes the same
    assignment without communicating."""
    if num_shards < 1:
        raise ValueError("num_shards must be >= 1")
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.lexsort((np.arange(len(lengths)), -lengths))
    heap = [(0, s) for s in range(num_shards)]
    heapq.heapify(heap)
    out: List[List[int]] = [[] for _ in range(num_shards)]
    for i in order:
        load, s = heapq.heappop(heap)
        out[s].append(int(i))
        heapq.heappush(heap, (load + int(lengths[i]), s))
    return [np.array(sorted(o), dtype=np.int64) for o in out]


def shard_loads(lengths: Sequence[int], shards: List[np.ndarray]) -> np.ndarray:
    lengths = np.asarray(lengths, dtype=np.int64)
    return np.array([int(lengths[s].sum()) for s in shards], dtype=np.int64)


def merge_in_input_order(num_utts: int, shards: List[np.ndarray], parts: List[np.ndarray]) -> np.ndarray:
    """Inverse of the sharding for per-utterance results ([len(shard), dim] each)."""
    dim = parts[0].shape[1] if parts and parts[0].ndim == 2 else 0
    out = np.zeros((num_utts, dim), dtype=parts[0].dtype if parts else np.float32)
    for idx, p in zip(shards, parts):
        out[idx] = p
    return out
