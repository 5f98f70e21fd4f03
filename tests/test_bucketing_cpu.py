"""Row 8f #4 host logic: the max-frames bucketing restatement against batches formed by the reference's own functions
(tests/golden/bucketing.json, oracle/make_golden_bucketing.py) and the rank assignment policies."""
import json
import os

from auto_avsr_b200.bucketing import assign_to_ranks, max_frames_batches, padded_frames

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bucketing.json")


def test_batches_equal_the_reference_custom_bucket_dataset():
    cases = json.load(open(GOLDEN))
    assert len(cases) == 5
    for c in cases:
        got = max_frames_batches(c["lengths"], c["max_frames"], c["num_buckets"])
        assert got == c["batches"]
        flat = sorted(i for b in got for i in b)
        assert flat == list(range(len(c["lengths"])))                     # every utterance exactly once
        assert all(sum(c["lengths"][i] for i in b) <= c["max_frames"] for b in got)


def test_rank_assignment_policies():
    c = json.load(open(GOLDEN))[2]
    lengths, batches = c["lengths"], c["batches"]
    for W in (1, 2, 4, 8):
        ref = assign_to_ranks(batches, lengths, W, "reference")
        bal = assign_to_ranks(batches, lengths, W, "balanced")
        assert [i for r in range(W) for i in ref[r]].__len__() == len(batches)
        assert ref == [list(range(r, len(batches), W)) for r in range(W)]  # batch i -> rank i % W
        assert sorted(i for r in bal for i in r) == list(range(len(batches)))
        load = lambda plan: max(sum(padded_frames(batches[i], lengths) for i in r) for r in plan)   # noqa: E731
        assert load(bal) <= load(ref)


def test_empty_and_too_long():
    import pytest
    assert max_frames_batches([], 1600) == []
    with pytest.raises(ValueError):
        max_frames_batches([10, 2000], 1600)
