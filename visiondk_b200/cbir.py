"""CBIR eval path on B200: extract -> index -> search (Seams 5 and 6).

Mirrors, by name and argument meaning:
  FeatureExtractor.extract_cbir   models/faceX/face_model.py:120-144
  index(...)                      engine/cbir/evaluation.py:106-169   (cbir_eval.py:35-96)
  search(...)                     engine/cbir/evaluation.py:171-200   (cbir_eval.py:98-122)
with the arithmetic on the sm_100a kernels: the backbone's `embed()` fuses F.normalize into the neck epilogue,
embeddings stay in HBM between extraction and search (the reference copies every batch to the host,
face_model.py:140), and the faiss objects are replaced by visiondk_b200.retrieval.FlatIPIndex.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch

from .retrieval import FlatIPIndex


class FeatureExtractor:
    """models/faceX/face_model.py:88-144 (CBIR branch)."""

    def __init__(self, model):
        self.model = model

    @torch.no_grad()
    def extract_cbir_device(self, dataloader, device) -> torch.Tensor:
        """Device-resident fp32 [N, feat_dim], rows L2-normalised, dataloader order preserved."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("visiondk_b200 extraction runs on CUDA (sm_100a) only; there is no CPU fallback")
        model = self.model
        model.eval()
        model.to(device)
        feats = []
        for tensors in dataloader:
            if isinstance(tensors, (list, tuple)):
                tensors = tensors[0]
            tensors = tensors.to(device, non_blocking=True)
            feats.append(model.embed(tensors, l2_normalize=True))
        if not feats:
            return torch.empty((0, model.feat_dim), dtype=torch.float32, device=device)
        return torch.cat(feats, dim=0)

    def extract_cbir(self, dataloader, device) -> np.ndarray:
        """Reference signature: numpy float32 [N, feat_dim] (one device->host copy at the end, not one per batch)."""
        return self.extract_cbir_device(dataloader, device).cpu().numpy()


def index(extractor: FeatureExtractor, gallery_dataloader, device, logger=None, index_factory: str = "Flat",
          memmap_feat_dim: Optional[int] = None, memmap_dtype=np.float16, memmap_save_path: Optional[str] = None,
          memmap_load_embedding: bool = False) -> FlatIPIndex:
    """engine/cbir/evaluation.py:106-169: encode the gallery, build the flat inner-product index (resident on
    `device`), optionally save / load the embeddings as a raw np.memmap (:124-152)."""
    if index_factory != "Flat":
        raise ValueError("only the 'Flat' (exact inner product) index of the reference's CBIR path is built")
    device = torch.device(device)
    if memmap_load_embedding:
        emb = np.memmap(memmap_save_path, mode="r", dtype=memmap_dtype).reshape(-1, memmap_feat_dim)
        emb = torch.from_numpy(np.ascontiguousarray(emb, dtype=np.float32)).to(device)
    else:
        emb = extractor.extract_cbir_device(gallery_dataloader, device)
        if memmap_save_path is not None:
            if logger is not None:
                logger.console(f"saving embeddings at {memmap_save_path}...")
            host = emb.cpu().numpy()
            mm = np.memmap(memmap_save_path, shape=host.shape, mode="w+", dtype=host.dtype)
            mm[:] = host
            mm.flush()
    faiss_index = FlatIPIndex(emb.shape[-1], device)
    if logger is not None:
        logger.console("Adding embeddings...")
    faiss_index.train(emb)
    faiss_index.add(emb)
    return faiss_index


def search(extractor: FeatureExtractor, query_dataloader, faiss_index: FlatIPIndex, device, logger=None, k: int = 100,
           batch_size: int = 256):
    """engine/cbir/evaluation.py:171-200.  The reference searches in `batch_size` slices because faiss wants host
    arrays per call; here the whole query block is scored in one pass (each gallery tile is read once for all
    queries), so `batch_size` is accepted and ignored.  Returns numpy (scores [n,k] desc, ids [n,k], -1 padded)."""
    q = extractor.extract_cbir_device(query_dataloader, device)
    if logger is not None:
        logger.console("Searching ...")
    scores, ids = faiss_index.search_device(q, k)
    faiss_index.check_status()
    return scores.cpu().numpy(), ids.cpu().numpy()
