"""CPU checks of oracle/retrieval.py: its canonical definitions against direct fp64 evaluation and torch."""
import numpy as np
import torch

from oracle import retrieval as R


def test_canonical_dot_is_a_correctly_ordered_fp64_sum():
    rng = np.random.default_rng(0)
    a = rng.standard_normal((7, 512)).astype(np.float32)
    b = rng.standard_normal((7, 512)).astype(np.float32)
    got = R.canonical_dot(a, b)
    ref = np.array([np.float32(sum(float(x) * float(y) for x, y in zip(ar, br))) for ar, br in zip(a, b)])
    assert np.max(np.abs(got.astype(np.float64) - ref)) <= 2e-6  # same value up to the final float32 rounding
    # explicit restatement of the fixed order for one row
    p = a[0].astype(np.float64) * b[0].astype(np.float64)
    lanes = [sum_seq(p[l::32]) for l in range(32)]
    for off in (16, 8, 4, 2, 1):
        lanes = [lanes[l] + lanes[l ^ off] for l in range(32)]
    assert np.float32(lanes[0]) == got[0]


def sum_seq(v):
    acc = np.float64(0.0)
    for x in v:
        acc = acc + x
    return acc


def test_l2_normalize_matches_torch_to_two_ulp():
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((64, 512)) * 5).astype(np.float32)
    x[3] = 0
    got = R.l2_normalize(x)
    ref = torch.nn.functional.normalize(torch.from_numpy(x)).numpy()
    assert np.all(np.abs(got - ref) <= np.spacing(np.abs(ref)).astype(np.float32) * 2.0 + 1e-45)
    assert np.all(got[3] == 0)


def test_flat_ip_search_matches_fp64_ranking_and_pads():
    rng = np.random.default_rng(2)
    q = R.l2_normalize(rng.standard_normal((9, 128)).astype(np.float32))
    g = R.l2_normalize(rng.standard_normal((500, 128)).astype(np.float32))
    s, i = R.flat_ip_search(q, g, 20)
    ref = np.argsort(-(q.astype(np.float64) @ g.astype(np.float64).T), axis=1, kind="stable")[:, :20]
    assert np.array_equal(i, ref)
    assert np.all(np.diff(s, axis=1) <= 0)
    s2, i2 = R.flat_ip_search_candidates(q, g, 20)
    assert np.array_equal(i, i2) and np.array_equal(s, s2)
    s3, i3 = R.flat_ip_search(q, g[:5], 8)
    assert np.all(i3[:, 5:] == -1) and np.all(s3[:, 5:] == R.FLT_LOWEST)


def test_tie_rule_is_id_ascending_and_merge_is_consistent():
    rng = np.random.default_rng(3)
    base = R.l2_normalize(rng.standard_normal((30, 64)).astype(np.float32))
    g = np.concatenate([base, base])
    q = R.l2_normalize(rng.standard_normal((4, 64)).astype(np.float32))
    s, i = R.flat_ip_search(q, g, 10)
    for r in range(4):
        for j in range(9):
            assert s[r, j] > s[r, j + 1] or i[r, j] < i[r, j + 1]
    parts = [R.flat_ip_search(q, g[a:b], 10, id_offset=a) for a, b in ((0, 17), (17, 41), (41, 60))]
    ms, mi = R.merge_topk([p[0] for p in parts], [p[1] for p in parts], 10)
    assert np.array_equal(mi, i) and np.array_equal(ms, s)
