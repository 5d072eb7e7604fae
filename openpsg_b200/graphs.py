"""CUDA-graph bookkeeping shared by the relation-query pipeline and the LLM decode loop.

Real PSG images change object count, token-grid shape and instruction length almost every image, so capturing a graph
the first time a signature is seen would pay warm-up + synchronise + capture for graphs that are never replayed, and
evict graphs that are.  A signature is therefore captured only once it has been seen ``capture_after`` times (first
sightings run eagerly, through the same kernels), and the cache evicts the least recently USED entry.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Hashable, Optional


class GraphCache:
    def __init__(self, max_entries: int = 4, capture_after: int = 2, max_tracked: int = 256):
        self.max_entries = max_entries
        self.capture_after = capture_after
        self.max_tracked = max_tracked
        self.entries: "OrderedDict[Hashable, Any]" = OrderedDict()
        self.seen: "OrderedDict[Hashable, int]" = OrderedDict()
        self.captures = 0
        self.evictions = 0

    def lookup(self, key) -> Optional[Any]:
        e = self.entries.get(key)
        if e is not None:
            self.entries.move_to_end(key)
        return e

    def should_capture(self, key) -> bool:
        """Count one sighting of ``key``; True once it has been seen ``capture_after`` times."""
        n = self.seen.pop(key, 0) + 1
        self.seen[key] = n
        while len(self.seen) > self.max_tracked:
            self.seen.popitem(last=False)
        return n >= self.capture_after

    def insert(self, key, entry) -> None:
        while len(self.entries) >= self.max_entries:
            self.entries.popitem(last=False)           # least recently used; frees its private memory pool
            self.evictions += 1
        self.entries[key] = entry
        self.captures += 1

    def clear(self) -> None:
        self.entries.clear()
        self.seen.clear()

    def __len__(self):
        return len(self.entries)
