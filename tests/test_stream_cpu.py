"""Host logic of the mixed-size stream scheduler (etch_b200/stream.py): pure Python, no GPU."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_plan_covers_every_scan_once_and_batches_share_a_size():
    from etch_b200 import stream
    rng = random.Random(3)
    sizes = [rng.choice([5000, 10000, 20000]) for _ in range(101)]
    for world in (1, 2, 8):
        for batch in (1, 8, 16):
            plan = stream.plan_stream(sizes, world, batch)
            assert len(plan) == world
            seen = []
            for rank_batches in plan:
                for n, ids in rank_batches:
                    assert 1 <= len(ids) <= batch
                    assert all(sizes[i] == n for i in ids)
                    assert ids == sorted(ids)
                    seen += ids
            assert sorted(seen) == list(range(len(sizes)))


def test_plan_is_balanced_and_deterministic():
    from etch_b200 import stream
    rng = random.Random(5)
    sizes = [rng.choice([5000, 10000, 20000]) for _ in range(256)]
    plan = stream.plan_stream(sizes, 8, 8)
    loads = [sum(stream.batch_cost(len(ids), n) for n, ids in rb) for rb in plan]
    biggest = max(stream.batch_cost(len(ids), n) for rb in plan for n, ids in rb)
    assert max(loads) - min(loads) <= biggest          # LPT bound: no rank is ahead by more than one batch
    assert plan == stream.plan_stream(sizes, 8, 8)
    # longest first within a rank
    for rb in plan:
        costs = [stream.batch_cost(len(ids), n) for n, ids in rb]
        assert costs == sorted(costs, reverse=True)


def test_plan_edge_cases():
    import pytest
    from etch_b200 import stream
    assert stream.plan_stream([], 4, 8) == [[], [], [], []]
    assert stream.plan_stream([777], 2, 8) == [[(777, [0])], []]
    with pytest.raises(ValueError):
        stream.plan_stream([5000, 0], 1, 8)
    with pytest.raises(ValueError):
        stream.plan_stream([5000], 0, 8)
