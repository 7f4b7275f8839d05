"""oracle/preprocess.py (the reference's val transform list restated: ResizeAndPadding2Square -> ToTensor -> Normalize,
dataset/transforms.py:325-365, 466-477) pinned against the libraries the reference calls — the installed Pillow's
`Image.resize(..., BILINEAR)` / `ImageOps.expand` and torchvision's `ToTensor` / `Normalize` — bit for bit."""
import numpy as np
import pytest
import torch

from oracle import preprocess as P

PIL = pytest.importorskip("PIL")
from PIL import Image, ImageOps  # noqa: E402

SHAPES = [(640, 480, 224), (100, 300, 224), (224, 224, 224), (1000, 37, 224), (50, 60, 224), (333, 500, 336), (7, 5, 64),
          (225, 224, 224), (1920, 1080, 224), (31, 97, 112)]


def reference_transform(img, size, mean, std):
    """The reference's own steps, executed with PIL / torchvision (dataset/transforms.py:340-362 restated call by call)."""
    import torchvision.transforms as T
    image = Image.fromarray(img)
    width, height = image.size
    scale_factor = size / max(width, height)
    new_width, new_height = int(width * scale_factor), int(height * scale_factor)
    image = image.resize((new_width, new_height), Image.BILINEAR)
    pad_width, pad_height = (size - new_width) // 2, (size - new_height) // 2
    padded = ImageOps.expand(image, (pad_width, pad_height, size - new_width - pad_width, size - new_height - pad_height), fill=(0, 0, 0))
    return T.Normalize(mean=mean, std=std)(T.ToTensor()(padded)).numpy()


@pytest.mark.parametrize("w,h,size", SHAPES)
def test_resize_matches_pillow_bit_for_bit(w, h, size):
    rng = np.random.default_rng(w * 7 + h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    nw, nh, left, top = P.resized_shape(w, h, size)
    ref = np.asarray(Image.fromarray(img).resize((nw, nh), Image.BILINEAR))
    assert np.array_equal(P.pil_resize_bilinear(img, nw, nh), ref)


@pytest.mark.parametrize("w,h,size", SHAPES[:6])
def test_whole_transform_matches_pil_and_torchvision(w, h, size):
    rng = np.random.default_rng(w + 13 * h)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    mean, std = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    ref = reference_transform(img, size, mean, std)
    got = P.resize_pad_normalize(img, size, mean, std)
    assert got.shape == (3, size, size) and got.dtype == np.float32
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_oracle_matches_the_reference_pipeline_golden():
    """tests/golden/preprocess_ref.npz was minted by executing the reference's OWN dataset/transforms.py
    (`create_AugTransforms` on the val augment list of its configs/faceX/cbir.yaml, oracle/make_golden.py::preprocess):
    the oracle must reproduce every stored tensor bit for bit."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess_ref.npz"))
    mean, std = tuple(float(v) for v in z["mean"]), tuple(float(v) for v in z["std"])
    assert int(z["cfg_size"]) == 224 and int(z["count"]) >= 12
    for n in range(int(z["count"])):
        img, size, ref = z[f"img{n}"], int(z[f"size{n}"]), z[f"out{n}"]
        got = P.resize_pad_normalize(img, size, mean, std)
        assert got.shape == ref.shape and np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (n, img.shape, size)
