"""visiondk_b200/metrics.py (tensor CBIR metrics) against values produced by the REFERENCE's own CBIRMetrics / compute_metrics
(tests/golden/cbir_metrics.npz, minted by oracle/make_golden.py: cbir_metrics from engine/cbir/evaluation.py's source, with
sklearn's roc_auc_score / ndcg_score) — including tied scores."""
import os

import numpy as np
import pytest
import torch

from visiondk_b200.metrics import cbir_metrics, relevance_from_labels, relevance_from_positive_lists

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load():
    z = np.load(os.path.join(GOLD, "cbir_metrics.npz"))
    ids, scores = torch.from_numpy(z["ids"]), torch.from_numpy(z["scores"])
    return z, ids, scores, torch.from_numpy(z["q_label"]), torch.from_numpy(z["gal_label"])


def test_metrics_match_the_reference_values_from_labels():
    z, ids, scores, ql, gl = load()
    rel, n_pos = relevance_from_labels(ids, ql, gl)
    got = cbir_metrics(ids, scores, rel, n_pos, [int(c) for c in z["cutoffs"]])
    ref = dict(zip([str(n) for n in z["metric_names"]], z["metric_values"]))
    assert set(got) == set(ref)
    for name, v in ref.items():
        assert abs(got[name] - v) <= 1e-12 + 1e-9 * abs(v), (name, got[name], v)


def test_positive_lists_give_the_same_relevance_as_labels():
    z, ids, scores, ql, gl = load()
    rel, n_pos = relevance_from_labels(ids, ql, gl)
    pos = [np.flatnonzero(z["gal_label"] == z["q_label"][q]) for q in range(ids.shape[0])]
    ptr = torch.tensor(np.concatenate([[0], np.cumsum([len(p) for p in pos])]), dtype=torch.int64)
    idx = torch.tensor(np.concatenate(pos), dtype=torch.int64)
    rel2, n_pos2 = relevance_from_positive_lists(ids, ptr, idx, len(z["gal_label"]))
    assert torch.equal(rel, rel2) and torch.equal(n_pos, n_pos2)


def test_padding_and_error_cases():
    z, ids, scores, ql, gl = load()
    ids = ids.clone()
    ids[0, -3:] = -1  # fewer than k results for a query: -1 never counts as relevant
    rel, n_pos = relevance_from_labels(ids, ql, gl)
    assert not rel[0, -3:].any()
    m = cbir_metrics(ids, scores, rel, n_pos, [1, 3, 10], metrics=("mrr", "precision", "recall", "ndcg"))
    assert all(0.0 <= v <= 1.0 for v in m.values())
    with pytest.raises(ValueError):
        cbir_metrics(ids, scores, rel, n_pos, [1, 3, 10], metrics=("auc",))
    with pytest.raises(ValueError):
        cbir_metrics(ids, scores, rel, torch.zeros_like(n_pos), [1], metrics=("recall",))


def test_face_verification_matches_the_reference_evaluator():
    """tests/golden/face_verification.npz: mean / std produced by the reference's own Evaluator.test_one_model on 6000 synthetic
    pairs (oracle/make_golden.py: face_verification)."""
    from visiondk_b200.metrics import face_verification_accuracy
    z = np.load(os.path.join(GOLD, "face_verification.npz"))
    feats, pairs = z["feats"], z["pairs"]
    scores = np.array([np.dot(feats[a], feats[b]) for a, b, _ in pairs], dtype=np.float32)  # the reference's np.dot per pair
    mean, std = face_verification_accuracy(torch.from_numpy(scores), torch.from_numpy(pairs[:, 2]))
    assert abs(mean - float(z["mean"])) <= 1e-12 and abs(std - float(z["std"])) <= 1e-12, (mean, std, float(z["mean"]), float(z["std"]))
    with pytest.raises(ValueError):
        face_verification_accuracy(torch.from_numpy(scores[:-1]), torch.from_numpy(pairs[:-1, 2]))
