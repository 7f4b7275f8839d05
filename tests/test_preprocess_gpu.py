"""The device preprocessing (csrc/preprocess.cu through visiondk_b200.preprocess) against oracle/preprocess.py — itself pinned
bit for bit against Pillow + torchvision by tests/test_oracle_preprocess_cpu.py: byte / fixed-point work, so the bar is
bit-exact fp32 outputs for every image of a mixed-size batch (up- and down-scaling, portrait / landscape / square, extreme
aspect ratios, a side that needs no resampling)."""
import numpy as np
import pytest
import torch

from oracle import preprocess as P
from visiondk_b200.preprocess import ImagePreprocessor, resize_pad_normalize

pytestmark = pytest.mark.gpu

SHAPES = [(640, 480), (100, 300), (224, 224), (1000, 37), (50, 60), (333, 500), (7, 5), (225, 224), (1920, 1080), (31, 97), (224, 100)]


@pytest.mark.parametrize("size", [224, 336, 64])
def test_mixed_batch_is_bit_exact(lib, size):
    rng = np.random.default_rng(size)
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for (w, h) in SHAPES]
    imgs[2][:] = 255   # saturated image: clip8 / rounding at the top of the range
    imgs[4][:] = 0
    out = resize_pad_normalize(imgs, size=size).cpu().numpy()
    assert out.shape == (len(imgs), 3, size, size)
    for i, im in enumerate(imgs):
        ref = P.resize_pad_normalize(im, size)
        assert np.array_equal(out[i].view(np.uint32), ref.view(np.uint32)), (i, im.shape, np.abs(out[i] - ref).max())


def test_preprocessor_reuse_and_custom_statistics(lib):
    rng = np.random.default_rng(3)
    pre = ImagePreprocessor(112, mean=(0.5, 0.5, 0.5), std=(0.25, 0.5, 1.0))
    for n in (3, 1, 5):
        imgs = [rng.integers(0, 256, (int(rng.integers(20, 400)), int(rng.integers(20, 400)), 3), dtype=np.uint8) for _ in range(n)]
        out = pre(imgs).cpu().numpy()
        for i, im in enumerate(imgs):
            ref = P.resize_pad_normalize(im, 112, (0.5, 0.5, 0.5), (0.25, 0.5, 1.0))
            assert np.array_equal(out[i].view(np.uint32), ref.view(np.uint32))
    assert pre([]).shape == (0, 3, 112, 112)
    with pytest.raises(ValueError):
        pre([np.zeros((4, 4), np.uint8)])


def test_preprocessed_batch_feeds_the_extractor(lib):
    """decoded images -> device preprocessing -> TimmWrapper.embed: the extract path fed from bytes instead of synthetic tensors."""
    from visiondk_b200.backbone import TimmWrapper
    rng = np.random.default_rng(5)
    imgs = [rng.integers(0, 256, (int(rng.integers(40, 300)), int(rng.integers(40, 300)), 3), dtype=np.uint8) for _ in range(6)]
    x = resize_pad_normalize(imgs, size=64)
    model = TimmWrapper("toy", 32, 64, pretrained=False, depths=(1, 1, 1, 1), dims=(64, 64, 128, 128)).cuda().eval()
    emb = model.embed(x, l2_normalize=True)
    assert emb.shape == (6, 32) and torch.isfinite(emb).all()
