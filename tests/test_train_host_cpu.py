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


def test_backward_sections_partition_the_parameters_for_every_architecture():
    """Host logic of the DDP overlap (no GPU): the unit ranges tile [0, units) in order and every parameter is handed over exactly
    once, in runs that are contiguous in named_parameters() order (= contiguous slices of the fused optimizer's flat buffer)."""
    from visiondk_b200.backbone import CONVNEXT_ARCHS, TimmWrapper
    from visiondk_b200.vit import VIT_ARCHS, ViTWrapper
    models = [TimmWrapper(name, 64, 64, pretrained=False) for name in ("convnext_atto", "convnext_tiny", "convnext_base")]
    models += [ViTWrapper(name, 64, 224, pretrained=False) for name in ("vit_tiny_patch16_224", "vit_base_patch16_224")]
    models.append(ViTWrapper("x", 64, 64, pretrained=False, patch=16, dim=128, depth=1, heads=2))
    for m in models:
        names = [n for n, _ in m.named_parameters()]
        pos = {n: i for i, n in enumerate(names)}
        sec = m.backward_sections()
        units = (1 + sum(m.model.depths) + 4) if hasattr(m.model, "depths") else m.model.depth + 2
        assert sec[0][0][0] == 0 and sec[-1][0][1] == units
        for (a, b), (c, d) in zip([s[0] for s in sec], [s[0] for s in sec][1:]):
            assert a < b == c < d
        seen = [n for _, ns in sec for n in ns]
        assert sorted(seen) == sorted(names)
        for _, ns in sec:
            idx = sorted(pos[n] for n in ns)
            assert idx == list(range(idx[0], idx[0] + len(idx))), "section is not a contiguous run of parameters"
