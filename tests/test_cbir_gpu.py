"""GPU test of the extract -> index -> search glue (visiondk_b200/cbir.py), the seams of face_model.py:120-144 and
engine/cbir/evaluation.py:106-200: embeddings within the stated tolerance of the fp32 oracle, and the search over
those embeddings bit-exact against the oracle's flat inner-product search."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import retrieval as R
from oracle.convnext import TimmWrapperOracle, randomize_
from visiondk_b200.backbone import TimmWrapper
from visiondk_b200.cbir import FeatureExtractor, index, search

pytestmark = pytest.mark.gpu


class HostLoader:
    """Pinned host batches, like DataLoader(pin_memory=True) over CBIRDatasets (one tensor per batch)."""

    def __init__(self, x, bs):
        self.x, self.bs = x, bs

    def __iter__(self):
        for a in range(0, self.x.shape[0], self.bs):
            yield self.x[a:a + self.bs].pin_memory()


def test_extract_index_search_pipeline(lib, tmp_path):
    depths, dims = (1, 1, 2, 1), (32, 64, 128, 256)
    oracle = randomize_(TimmWrapperOracle("toy", 64, 64, depths=depths, dims=dims), seed=5).eval()
    model = TimmWrapper("toy", 64, 64, pretrained=False, depths=depths, dims=dims)
    model.load_state_dict(oracle.state_dict(), strict=True)
    torch.manual_seed(0)
    gallery_x, query_x = torch.randn(300, 3, 64, 64), torch.randn(37, 3, 64, 64)
    ext = FeatureExtractor(model)
    emb = ext.extract_cbir(HostLoader(gallery_x, 64), "cuda")
    assert isinstance(emb, np.ndarray) and emb.shape == (300, 64) and emb.dtype == np.float32
    with torch.no_grad():
        ref = F.normalize(oracle(gallery_x)).numpy()
    cos = (emb * ref).sum(1)
    assert cos.min() >= 0.999 and np.allclose(np.linalg.norm(emb, axis=1), 1.0, atol=1e-5)
    # order is the dataloader's order, independent of batch size
    emb2 = ext.extract_cbir(HostLoader(gallery_x, 37), "cuda")
    assert np.allclose(emb, emb2, atol=2e-3)

    idx = index(ext, HostLoader(gallery_x, 64), "cuda", memmap_save_path=str(tmp_path / "gallery.f32"))
    assert idx.ntotal == 300
    scores, ids = search(ext, HostLoader(query_x, 64), idx, "cuda", k=10)
    q_emb = ext.extract_cbir(HostLoader(query_x, 64), "cuda")
    g_emb = np.memmap(tmp_path / "gallery.f32", mode="r", dtype=np.float32).reshape(-1, 64)
    ref_s, ref_i = R.flat_ip_search(q_emb, np.asarray(g_emb), 10)
    assert np.array_equal(ids, ref_i) and np.array_equal(scores.view(np.uint32), ref_s.view(np.uint32))
    # reload path (engine/cbir/evaluation.py:124-130)
    idx2 = index(ext, None, "cuda", memmap_feat_dim=64, memmap_dtype=np.float32, memmap_save_path=str(tmp_path / "gallery.f32"),
                 memmap_load_embedding=True)
    s2, i2 = idx2.search(q_emb, 10)
    assert np.array_equal(i2, ref_i)
