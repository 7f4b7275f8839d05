"""visiondk_b200/metrics.py on DEVICE tensors (where the product path runs it: engine/cbir/evaluation.py::valuate keeps ids and
scores on the GPU) against the values produced by the reference's own CBIRMetrics / compute_metrics + sklearn and
Evaluator.test_one_model (tests/golden/cbir_metrics.npz, face_verification.npz; the same pins as tests/test_metrics_cpu.py)."""
import os

import numpy as np
import pytest
import torch

from engine.cbir.evaluation import compute_metrics
from visiondk_b200.metrics import cbir_metrics, face_verification_accuracy, relevance_from_labels

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_cbir_metrics_on_device_match_the_reference_values(lib):
    z = np.load(os.path.join(GOLD, "cbir_metrics.npz"))
    ids, scores = torch.from_numpy(z["ids"]).cuda(), torch.from_numpy(z["scores"]).cuda()
    ql, gl = torch.from_numpy(z["q_label"]).cuda(), torch.from_numpy(z["gal_label"]).cuda()
    rel, n_pos = relevance_from_labels(ids, ql, gl)
    assert rel.is_cuda and n_pos.is_cuda
    cut = [int(c) for c in z["cutoffs"]]
    ref = dict(zip([str(n) for n in z["metric_names"]], z["metric_values"]))
    for got in (cbir_metrics(ids, scores, rel, n_pos, cut), compute_metrics(ids, scores, ql, gl, cutoffs=cut)):
        assert set(got) == set(ref)
        for name, v in ref.items():
            assert abs(got[name] - v) <= 1e-12 + 1e-9 * abs(v), (name, got[name], v)


def test_face_verification_on_device_matches_the_reference_evaluator(lib):
    z = np.load(os.path.join(GOLD, "face_verification.npz"))
    feats, pairs = torch.from_numpy(z["feats"]).cuda(), torch.from_numpy(z["pairs"]).cuda()
    # pair scores = dot products of unit features (engine/faceX/evaluation.py:61-66); fp32 on the device like np.dot on the host
    scores = (feats[pairs[:, 0]].double() * feats[pairs[:, 1]].double()).sum(dim=1).float()
    mean, std = face_verification_accuracy(scores, pairs[:, 2])
    assert abs(mean - float(z["mean"])) <= 2e-4 and abs(std - float(z["std"])) <= 2e-4, (mean, std)
