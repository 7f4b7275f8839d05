"""Oracle for the eval-time image preprocessing (TEST INFRASTRUCTURE — see oracle/__init__.py).

Reference behaviour restated here (the `val.augment` list of configs/faceX/{face,cbir}.yaml):
  * dataset/transforms.py:325-365  ResizeAndPadding2Square(size, training=False): scale the longer side to `size` with
        PIL `Image.resize((new_w, new_h), Image.BILINEAR)` (new_w = int(w * size / max(w, h)), same for h), then
        `ImageOps.expand` with black borders, image centred (left/top pad = (size - new) // 2);
  * dataset/transforms.py:466-468  T.ToTensor(): uint8 HWC -> float32 CHW / 255;
  * dataset/transforms.py:474-477  T.Normalize(mean, std): (x - mean) / std per channel (the black border is normalised too).

The resize is Pillow's (third-party, not vendored by the reference; pinned by tests/test_oracle_preprocess_cpu.py against the
installed Pillow itself).  Its published algorithm for 8-bit images (Pillow src/libImaging/Resample.c, ImagingResampleInner):
separable, horizontal pass then vertical pass with a uint8 intermediate image; for every output coordinate `xx`
    scale = in_size / out_size;  filterscale = max(scale, 1);  support = 1.0 * filterscale          (bilinear: support 1)
    center = (xx + 0.5) * scale;  xmin = max(int(center - support + 0.5), 0);  xmax = min(int(center + support + 0.5), in_size)
    w[x] = triangle((x + xmin - center + 0.5) / filterscale), normalised to sum 1                       (double precision)
    k[x] = int(w[x] * 2^22 +- 0.5)                                                                       (fixed point)
    out  = clip8((2^21 + sum_x in[xmin + x] * k[x]) >> 22)
A pass is skipped when its size does not change.
"""
from __future__ import annotations

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resize_coeffs(in_size: int, out_size: int):
    """(xmin [out], count [out], k int32 [out, kmax]) of one resampling pass — Resample.c precompute_coeffs + normalize_coeffs_8bpc."""
    scale = float(in_size) / float(out_size)
    filterscale = scale if scale > 1.0 else 1.0
    support = 1.0 * filterscale
    kmax = int(np.ceil(support)) * 2 + 1
    xmins = np.zeros(out_size, np.int32)
    counts = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, kmax), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        n = xmax - xmin
        w = np.zeros(n, np.float64)
        ww = 0.0
        for x in range(n):
            a = (x + xmin - center + 0.5) * ss
            if a < 0.0:
                a = -a
            v = 1.0 - a if a < 1.0 else 0.0
            w[x] = v
            ww += v
        for x in range(n):
            if ww != 0.0:
                w[x] /= ww
        for x in range(n):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(v - 0.5) if w[x] < 0 else int(v + 0.5)
        xmins[xx], counts[xx] = xmin, n
    return xmins, counts, kk


def _resample_axis0(img: np.ndarray, out_size: int) -> np.ndarray:
    """One pass along axis 0 of a uint8 [n, m, c] image."""
    xmins, counts, kk = resize_coeffs(img.shape[0], out_size)
    out = np.empty((out_size,) + img.shape[1:], np.uint8)
    src = img.astype(np.int64)
    for xx in range(out_size):
        n = counts[xx]
        acc = (1 << (PRECISION_BITS - 1)) + np.tensordot(kk[xx, :n].astype(np.int64), src[xmins[xx]:xmins[xx] + n], axes=(0, 0))
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return out


def pil_resize_bilinear(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    """PIL `Image.fromarray(img).resize((new_w, new_h), Image.BILINEAR)` for a uint8 [h, w, c] array: horizontal pass, then vertical."""
    h, w = img.shape[:2]
    out = img
    if new_w != w:
        out = np.ascontiguousarray(_resample_axis0(np.ascontiguousarray(out.transpose(1, 0, 2)), new_w).transpose(1, 0, 2))
    if new_h != h:
        out = _resample_axis0(out, new_h)
    return out


def resized_shape(w: int, h: int, size: int):
    """new (w, h) and (left, top) padding of ResizeAndPadding2Square (transforms.py:344-357)."""
    scale_factor = size / max(w, h)
    new_w, new_h = int(w * scale_factor), int(h * scale_factor)
    return new_w, new_h, (size - new_w) // 2, (size - new_h) // 2


def resize_pad_normalize(img: np.ndarray, size: int, mean=(0.485, 0.456, 0.406), std=(0.229, 0.224, 0.225)) -> np.ndarray:
    """uint8 [h, w, 3] -> float32 [3, size, size]: the reference's val transform list, bit for bit."""
    h, w = img.shape[:2]
    new_w, new_h, left, top = resized_shape(w, h, size)
    canvas = np.zeros((size, size, 3), np.uint8)
    canvas[top:top + new_h, left:left + new_w] = pil_resize_bilinear(img, new_w, new_h)
    x = canvas.transpose(2, 0, 1).astype(np.float32) / np.float32(255.0)          # T.ToTensor
    m = np.asarray(mean, np.float32).reshape(3, 1, 1)
    s = np.asarray(std, np.float32).reshape(3, 1, 1)
    return ((x - m) / s).astype(np.float32)                                       # T.Normalize
