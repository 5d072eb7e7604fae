"""CPU tests of the CUDA-graph cache policy (openpsg_b200/graphs.py): capture on the second sighting, LRU eviction."""
from openpsg_b200.graphs import GraphCache


def test_capture_after_second_sighting_and_lru():
    c = GraphCache(max_entries=2, capture_after=2)
    assert c.lookup("a") is None and not c.should_capture("a")      # first sighting runs eagerly
    assert c.should_capture("a")
    c.insert("a", 1)
    assert not c.should_capture("b") and c.should_capture("b")
    c.insert("b", 2)
    assert c.lookup("a") == 1                                       # "a" becomes most recently used
    assert not c.should_capture("c") and c.should_capture("c")
    c.insert("c", 3)                                                # evicts "b", the least recently used
    assert c.lookup("b") is None and c.lookup("a") == 1 and c.lookup("c") == 3
    assert c.evictions == 1 and c.captures == 3 and len(c) == 2


def test_signatures_that_never_repeat_are_never_captured():
    c = GraphCache(max_entries=2, capture_after=2, max_tracked=8)
    for i in range(100):
        assert c.lookup(i) is None and not c.should_capture(i)
    assert len(c.seen) <= 8 and len(c) == 0
