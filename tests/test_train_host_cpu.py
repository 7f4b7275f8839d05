"""CPU checks of the train-step host logic: the closed-form cosine_with_warm schedule against the learning rates the
reference's own scheduler produced (tests/golden/step_sgd_ema.npz, engine/scheduler.py:47-57)."""
import os

import numpy as np

from visiondk_b200.train import cosine_with_warm_lr

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_cosine_with_warm_matches_reference_scheduler():
    z = np.load(os.path.join(GOLD, "step_sgd_ema.npz"))
    lrs = z["lrs"]  # lr in effect during step s, groups with base lr 0.01 and 0.1, warm 4, total 20, lrf 0.1
    for s in range(lrs.shape[0]):
        got = [cosine_with_warm_lr(s, base, 0.01, 4, 20, 0.1) for base in (0.01, 0.1)]
        np.testing.assert_allclose(got, lrs[s], rtol=1e-6)
