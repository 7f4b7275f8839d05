"""The evaluation ORCHESTRATION on image folders and pair files (engine/cbir/evaluation.py::valuate with a folder root,
engine/faceX/evaluation.py::valuate) run end to end on the CPU with test doubles standing in for the three device pieces —
the preprocessing kernels (played by the oracle), the backbone (a pooling "model") and the flat index (torch.topk).  What is
under test is the host logic between them: file order -> ids -> labels -> metrics, limits, the vis return, pair lookup, the
task dispatch.  (The real pieces have their own GPU tests; there is no CPU path in the product.)"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import preprocess as P

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402

SIZE = 32
AUGMENT = [{"resize_and_padding": {"size": SIZE, "training": False}}, {"to_tensor": "no_params"},
           {"normalize": {"mean": [0.485, 0.456, 0.406], "std": [0.229, 0.224, 0.225]}}]


class PoolModel:
    """Stands in for TimmWrapper: 4 x 4 average pooling of the normalised image as the embedding."""
    image_size, feat_dim = SIZE, 48

    def eval(self):
        return self

    def to(self, device):
        return self

    def embed(self, x, l2_normalize=False):
        f = F.adaptive_avg_pool2d(x.float(), 4).flatten(1)
        return F.normalize(f) if l2_normalize else f


class OraclePreprocessor:
    def __init__(self, size, mean, std, device="cpu"):
        self.size, self.mean, self.std = size, mean, std

    def __call__(self, images):
        if not images:
            return torch.empty((0, 3, self.size, self.size))
        return torch.from_numpy(np.stack([P.resize_pad_normalize(im, self.size, self.mean, self.std) for im in images]))


class TorchIndex:
    def __init__(self, emb):
        self.emb = emb

    def search_device(self, q, k, resolve_overflow=False):
        s, i = (q @ self.emb.T).topk(k, dim=1)
        return s, i


class Extractor:
    def __init__(self, model):
        self.model = model

    def extract_cbir_device(self, loader, device):
        return torch.cat([self.model.embed(b, l2_normalize=True) for b in loader])

    def extract_face(self, loader, device):
        out = {}
        for _, tensors, paths in loader:
            f = self.model.embed(tensors, l2_normalize=True).numpy()
            out.update({os.path.join(os.path.basename(os.path.dirname(p)), os.path.basename(p)): f[i] for i, p in enumerate(paths)})
        return out


def identity_image(rng, proto):
    w = h = int(rng.integers(30, 80))  # square: the 4 x 4 pattern survives the resize, so the pooling "model" recovers the identity
    img = np.asarray(Image.fromarray(proto).resize((w, h), Image.NEAREST)).astype(np.int16)
    return np.clip(img + rng.integers(-12, 13, img.shape), 0, 255).astype(np.uint8)


@pytest.fixture()
def doubles(monkeypatch):
    import engine.cbir.evaluation as EC
    import engine.faceX.evaluation as EF
    import visiondk_b200.preprocess as VP
    monkeypatch.setattr(VP, "ImagePreprocessor", OraclePreprocessor)
    monkeypatch.setattr(EC, "FeatureExtractor", Extractor)
    monkeypatch.setattr(EC, "index", lambda extractor, loader, device, logger=None: TorchIndex(extractor.extract_cbir_device(loader, device)))
    monkeypatch.setattr(EF, "FeatureExtractor", Extractor)
    return EC, EF


def test_cbir_valuate_on_a_folder(doubles, tmp_path):
    EC, _ = doubles
    rng = np.random.default_rng(2)
    for i in range(5):
        proto = rng.integers(0, 256, (4, 4, 3), dtype=np.uint8)
        for split, per in (("gallery", 4), ("query", 2)):
            os.makedirs(tmp_path / split / f"id{i}")
            for j in range(per):
                Image.fromarray(identity_image(rng, proto)).save(tmp_path / split / f"id{i}" / f"{j}.png")
    cfg = {"root": str(tmp_path), "nw": 2,
           "val": {"bs": 3, "augment": AUGMENT, "metrics": {"metrics": ["mrr", "recall", "precision", "auc", "ndcg"], "cutoffs": [1, 3]}}}
    m = EC.valuate(PoolModel(), cfg, "cpu", image_size=SIZE)
    assert set(m) == {"MRR@1", "MRR@3", "Recall@1", "Recall@3", "Precision@1", "Precision@3", "AUC@3", "nDCG@1", "nDCG@3"}
    assert m["MRR@1"] == 1.0 and m["Precision@3"] == 1.0 and m["Recall@3"] == 0.75  # every query finds its identity's images first
    results, scores, positives, queries = EC.valuate(PoolModel(), cfg, "cpu", vis=True, image_size=SIZE)
    assert len(results) == len(queries) == 10 and scores.shape == (10, 3)
    for r, pos, q in zip(results, positives, queries):
        assert len(pos) == 4 and set(r) <= set(pos)
        assert os.path.basename(os.path.dirname(q)) == os.path.basename(os.path.dirname(r[0]))
    # limits: the first 2 identities' gallery files and their queries
    m2 = EC.valuate(PoolModel(), cfg, "cpu", image_size=SIZE, gallery_limit=8, query_limit=4)
    assert m2["MRR@1"] == 1.0
    with pytest.raises(ValueError, match="resizes to 32"):
        EC.valuate(PoolModel(), cfg, "cpu", image_size=64)
    with pytest.raises(ValueError, match="neither a synthetic"):
        EC.valuate(PoolModel(), dict(cfg, root="some/hub-dataset"), "cpu", image_size=SIZE)


def test_face_valuate_on_a_pair_file_and_on_synthetic_identities(doubles, tmp_path):
    _, EF = doubles
    rng = np.random.default_rng(4)
    for i in range(6):
        proto = rng.integers(0, 256, (4, 4, 3), dtype=np.uint8)
        os.makedirs(tmp_path / "val" / f"p{i}")
        for j in range(3):
            Image.fromarray(identity_image(rng, proto)).save(tmp_path / "val" / f"p{i}" / f"{j}.png")
    rows = []
    for n in range(60):  # 10 folds of 6: three genuine, three impostor
        i = int(rng.integers(0, 6))
        if n % 6 < 3:
            a, b = rng.choice(3, 2, replace=False)
            rows.append(f"p{i}/{a}.png p{i}/{b}.png 1")
        else:
            o = int((i + 1 + rng.integers(0, 5)) % 6)
            rows.append(f"p{i}/{int(rng.integers(0, 3))}.png p{o}/{int(rng.integers(0, 3))}.png 0")
    (tmp_path / "pairs.txt").write_text("\n".join(rows) + "\n")
    cfg = {"root": str(tmp_path), "nw": 2, "val": {"bs": 4, "pair_txt": str(tmp_path / "pairs.txt"), "augment": AUGMENT}}
    mean, std = EF.valuate(PoolModel(), cfg, "cpu", image_size=SIZE)
    assert mean > 0.9 and 0.0 <= std < 0.1  # distinct patterns, mild noise: the folds separate (almost) perfectly
    (tmp_path / "pairs.txt").write_text("\n".join(rows[:-1]) + "\n")
    with pytest.raises(AssertionError, match="multiple of 10"):
        EF.valuate(PoolModel(), cfg, "cpu", image_size=SIZE)

    syn = {"root": "synthetic://cbir?ids=20&per_id=4&queries=5&noise=0.2", "nw": 0, "val": {"bs": 25, "augment": AUGMENT}}
    mean, std = EF.valuate(PoolModel(), syn, "cpu", image_size=SIZE, n_pairs=200)
    assert mean > 0.95 and 0.0 <= std < 0.05
