"""CBIR evaluation of the reference (engine/cbir/evaluation.py:106-291) on the B200 kernels — same function names and
argument meaning: `index`, `search`, `compute_metrics`, `valuate`.

`valuate` (:226-291) = build query / gallery loaders (shuffle=False) -> FeatureExtractor -> index(gallery) -> search(queries,
k = cutoffs[-1]) -> map returned ids to gallery entries, dropping -1 (:269-272) -> compute_metrics against every query's
positives (:274-291) -> {name: float}.  Embeddings, ids and scores stay on the device; relevance is an integer comparison of
identity labels (visiondk_b200.metrics) instead of the reference's per-query Python loops over path strings."""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

from engine.synthetic import SyntheticFaceData, is_synthetic
from visiondk_b200 import metrics as M
from visiondk_b200.cbir import FeatureExtractor, index, search  # noqa: F401  (reference exports the same names)


def compute_metrics(ids: torch.Tensor, scores: torch.Tensor, query_label: torch.Tensor, gallery_label: torch.Tensor,
                    metrics: Sequence[str] = ("mrr", "precision", "recall", "auc", "ndcg"),
                    cutoffs: Sequence[int] = (1, 3, 10)) -> Dict[str, float]:
    """engine/cbir/evaluation.py:202-224 with the retrieval result as tensors: `ids`/`scores` [Nq,k] as `search` returns them
    (-1 padded), identity labels of queries and gallery items.  Keys and values as the reference's metrics dict."""
    rel, n_pos = M.relevance_from_labels(ids, query_label, gallery_label)
    out: Dict[str, float] = {}
    for m in metrics:
        try:
            out.update(M.cbir_metrics(ids, scores, rel, n_pos, cutoffs, [m]))
        except ValueError as e:
            # sklearn's roc_auc_score raises when the retrieved pairs are all relevant (or all irrelevant), which would abort the
            # reference's in-training eval on an easy dataset; report the metric as undefined instead of losing the checkpoint
            if m != "auc" or "one class" not in str(e):
                raise
            out[f"AUC@{list(cutoffs)[-1]}"] = float("nan")
    return out


def valuate(model, data_cfg: dict, device, logger=None, vis: bool = False, image_size: Optional[int] = None,
            gallery_limit: Optional[int] = None, query_limit: Optional[int] = None):
    """engine/cbir/evaluation.py:226-291.  `data_cfg['root']` must be a synthetic:// URL (engine/synthetic.py) — reading image
    folders is the dataset layer the hot-path scope leaves out; callers with their own loaders use index / search /
    compute_metrics directly with their label tensors.  `gallery_limit` / `query_limit` bound the in-training eval."""
    device = torch.device(device)
    root = str(data_cfg["root"])
    if not is_synthetic(root):
        raise NotImplementedError("valuate: only synthetic:// data roots are built (dataset IO is outside the B200 hot-path scope); "
                                  "use index() / search() / compute_metrics() with your own loaders and label tensors")
    size = image_size if image_size is not None else getattr(model, "image_size")
    data = SyntheticFaceData(root, size, data_cfg["val"]["bs"], device)
    extractor = FeatureExtractor(model)
    faiss_index = index(extractor, data.gallery_batches(gallery_limit), device, logger=logger)
    cutoffs = list(data_cfg["val"]["metrics"]["cutoffs"])
    q = extractor.extract_cbir_device(data.query_batches(query_limit), device)
    if logger is not None:
        logger.console("Searching ...")
    scores, ids = faiss_index.search_device(q, cutoffs[-1], resolve_overflow=True)
    q_label, g_label = data.query_labels(query_limit), data.gallery_labels(gallery_limit)
    if vis:  # (retrieval results, scores, ground truths, queries) like :279
        return ids, scores, g_label, q_label
    out = compute_metrics(ids, scores, q_label, g_label, metrics=data_cfg["val"]["metrics"]["metrics"], cutoffs=cutoffs)
    return {k: float(v) for k, v in out.items()}
