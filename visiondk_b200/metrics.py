"""Evaluation metrics on tensors (the step right after top-k / pair scoring): CBIR MRR / Precision / Recall @cutoffs, AUC, nDCG, and
the LFW-style face pair-verification accuracy (`face_verification_accuracy`, engine/faceX/evaluation.py:34-113).

Replaces `CBIRMetrics` + `compute_metrics` (engine/cbir/evaluation.py:14-224), which walk Python lists of gallery PATH strings
per query (`x in label`, `np.intersect1d`, `np.isin`) and call sklearn's `roc_auc_score` / `ndcg_score` on the host.  Here the
retrieval result stays what the search returns — `ids [Nq, k]` (int64, -1 padded) and `scores [Nq, k]` — and relevance is an
integer comparison against identity labels (`gallery_label[ids] == query_label`), or a membership test against explicit
positive lists in CSR form; everything is a handful of batched tensor ops on whichever device the ids live on (no kernels of
ours: O(Nq k) integer / fp64 work).  Definitions follow the reference exactly, including sklearn's tie handling:

  MRR@c        1 / rank of the first relevant item if that rank <= c, else 0; mean over queries
  Precision@c  |relevant in top-c| / min(c, |positives of the query|)
  Recall@c     |relevant in top-c| / |positives of the query|
  AUC@c_last   roc_auc_score over ALL (relevance, score) pairs of all queries (Mann-Whitney U with tie-averaged ranks)
  nDCG@c       sklearn.metrics.ndcg_score(relevance, scores, k=c): ranking by score with tie-averaged gains, log2 discounts,
               ideal DCG from the relevances of the SAME k retrieved items; mean over queries
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch


def relevance_from_labels(ids: torch.Tensor, query_label: torch.Tensor, gallery_label: torch.Tensor):
    """rel [Nq, k] bool (False at -1 padding) and the number of positives of every query in the WHOLE gallery."""
    valid = ids >= 0
    gl = gallery_label.to(ids.device)[ids.clamp_min(0)]
    rel = (gl == query_label.to(ids.device)[:, None]) & valid
    # positives per query = count of its label in the gallery
    uniq, counts = torch.unique(gallery_label, return_counts=True)
    pos = torch.searchsorted(uniq, query_label.to(uniq.device))
    pos = pos.clamp_max(uniq.numel() - 1)
    n_pos = torch.where(uniq[pos] == query_label.to(uniq.device), counts[pos], torch.zeros_like(counts[pos]))
    return rel, n_pos.to(ids.device)


def relevance_from_positive_lists(ids: torch.Tensor, pos_ptr: torch.Tensor, pos_idx: torch.Tensor, n_gallery: int):
    """Explicit ground truth (the reference's `query_dataset.data['pos']`) in CSR form: positives of query q are
    pos_idx[pos_ptr[q]:pos_ptr[q+1]] (gallery indices)."""
    nq = ids.shape[0]
    q_of = torch.repeat_interleave(torch.arange(nq, device=pos_idx.device), (pos_ptr[1:] - pos_ptr[:-1]).to(pos_idx.device))
    keys = torch.sort(q_of * int(n_gallery) + pos_idx).values
    probe = torch.arange(nq, device=ids.device)[:, None] * int(n_gallery) + ids.clamp_min(0)
    rel = torch.isin(probe, keys.to(ids.device)) & (ids >= 0)
    return rel, (pos_ptr[1:] - pos_ptr[:-1]).to(ids.device)


def _tie_averaged_ranks(x: torch.Tensor) -> torch.Tensor:
    """1-based ranks of a 1-D fp64 tensor in ascending order, ties receiving the mean of their positions (scipy rankdata 'average')."""
    order = torch.argsort(x, stable=True)
    xs = x[order]
    n = x.numel()
    new_group = torch.ones(n, dtype=torch.bool, device=x.device)
    new_group[1:] = xs[1:] != xs[:-1]
    gid = torch.cumsum(new_group.to(torch.int64), 0) - 1
    pos = torch.arange(1, n + 1, dtype=torch.float64, device=x.device)
    ng = int(gid[-1].item()) + 1 if n else 0
    sums = torch.zeros(ng, dtype=torch.float64, device=x.device).index_add_(0, gid, pos)
    cnts = torch.zeros(ng, dtype=torch.float64, device=x.device).index_add_(0, gid, torch.ones_like(pos))
    ranks = torch.empty(n, dtype=torch.float64, device=x.device)
    ranks[order] = (sums / cnts)[gid]
    return ranks


def _ndcg(rel: torch.Tensor, scores: torch.Tensor, c: int) -> torch.Tensor:
    """sklearn.metrics.ndcg_score(rel, scores, k=c) per query (fp64)."""
    nq, k = rel.shape
    relf = rel.to(torch.float64)
    disc = 1.0 / torch.log2(torch.arange(k, dtype=torch.float64, device=rel.device) + 2.0)
    disc[c:] = 0.0
    order = torch.argsort(-scores.to(torch.float64), dim=1, stable=True)
    s_sorted = torch.gather(scores.to(torch.float64), 1, order)
    r_sorted = torch.gather(relf, 1, order)
    # tie groups of equal score along each row: every member gets the mean relevance of its group
    new_group = torch.ones((nq, k), dtype=torch.bool, device=rel.device)
    new_group[:, 1:] = s_sorted[:, 1:] != s_sorted[:, :-1]
    gid = torch.cumsum(new_group.to(torch.int64), dim=1) - 1 + torch.arange(nq, device=rel.device)[:, None] * k
    flat_gid = gid.reshape(-1)
    sums = torch.zeros(nq * k, dtype=torch.float64, device=rel.device).index_add_(0, flat_gid, r_sorted.reshape(-1))
    cnts = torch.zeros(nq * k, dtype=torch.float64, device=rel.device).index_add_(0, flat_gid, torch.ones(nq * k, dtype=torch.float64,
                                                                                                         device=rel.device))
    mean_rel = (sums / cnts.clamp_min(1.0))[flat_gid].reshape(nq, k)
    dcg = (mean_rel * disc[None, :]).sum(dim=1)
    ideal = (torch.sort(relf, dim=1, descending=True).values * disc[None, :]).sum(dim=1)
    return torch.where(ideal > 0, dcg / ideal.clamp_min(1e-300), torch.zeros_like(dcg))


def cbir_metrics(ids: torch.Tensor, scores: torch.Tensor, rel: torch.Tensor, n_pos: torch.Tensor, cutoffs: Sequence[int],
                 metrics: Sequence[str] = ("mrr", "precision", "recall", "auc", "ndcg")) -> Dict[str, float]:
    """compute_metrics(preds, preds_scores, labels, metrics, cutoffs) of the reference on tensors; keys and values as the
    reference's dict (`MRR@c`, `Precision@c`, `Recall@c`, `AUC@<last cutoff>`, `nDCG@c`)."""
    if (n_pos <= 0).any():
        raise ValueError("every query needs at least one positive in the gallery (the reference divides by len(label))")
    nq, k = ids.shape
    out: Dict[str, float] = {}
    relf = rel.to(torch.float64)
    csum = torch.cumsum(relf, dim=1)
    npos = n_pos.to(torch.float64)
    for m in metrics:
        if m == "mrr":
            any_hit = rel.any(dim=1)
            first = torch.argmax(rel.to(torch.int8), dim=1) + 1  # 1-based rank of the first hit (garbage where no hit)
            rr = torch.where(any_hit, 1.0 / first.to(torch.float64), torch.zeros(nq, dtype=torch.float64, device=ids.device))
            for c in cutoffs:
                out[f"MRR@{c}"] = float((rr * (first <= c)).sum().item() / nq)
        elif m == "precision":
            for c in cutoffs:
                hits = csum[:, min(c, k) - 1]
                out[f"Precision@{c}"] = float((hits / torch.minimum(torch.full_like(npos, float(c)), npos)).sum().item() / nq)
        elif m == "recall":
            for c in cutoffs:
                out[f"Recall@{c}"] = float((csum[:, min(c, k) - 1] / npos).sum().item() / nq)
        elif m == "auc":
            if (ids < 0).any():
                raise ValueError("AUC needs k valid results for every query (the reference flattens the unfiltered score matrix)")
            y = rel.reshape(-1)
            r = _tie_averaged_ranks(scores.reshape(-1).to(torch.float64))
            n1 = float(y.sum().item())
            n0 = float(y.numel()) - n1
            if n1 == 0 or n0 == 0:
                raise ValueError("AUC is undefined when only one class is present (sklearn raises as well)")
            out[f"AUC@{cutoffs[-1]}"] = float(((r[y].sum().item() - n1 * (n1 + 1) / 2.0) / (n1 * n0)))
        elif m == "ndcg":
            for c in cutoffs:
                out[f"nDCG@{c}"] = float(_ndcg(rel, scores, int(c)).sum().item() / nq)
        else:
            raise ValueError(f"{m} is not supported")
    return out


def face_verification_accuracy(scores: torch.Tensor, labels: torch.Tensor, num_thresholds: int = 1000, folds: int = 10):
    """Evaluator.test_one_model + getThreshold (engine/faceX/evaluation.py:34-113): LFW-style verification accuracy.

    `scores [N]` float32 cosine similarities of the N pairs in protocol order (N a multiple of `folds`; the reference hard-codes 600
    pairs per fold), `labels [N]` 1 = same identity.  For every fold the threshold is chosen on the other nine — the one of
    `num_thresholds` evenly spaced values in (min, max] of the TRAINING scores that maximises TPR - FPR (first maximum) — and the
    held-out fold is scored with it (score > t for genuine, score < t for impostor pairs).  Returns (mean accuracy, standard
    error of the mean with ddof = 1), computed like the reference: float32 scores, float64 thresholds.
    All folds are evaluated at once as [folds, thresholds, pairs] comparisons on the device the scores live on."""
    n = scores.numel()
    if n % folds != 0:
        raise ValueError("make sure the number of pairs is a multiple of 10 (check_nps, evaluation.py:110-113)")
    per = n // folds
    s32 = scores.reshape(folds, per).to(torch.float32)
    lab = labels.reshape(folds, per).to(torch.bool)
    dev = scores.device
    accs = []
    steps = torch.arange(1, num_thresholds + 1, dtype=torch.float64, device=dev)
    for f in range(folds):
        keep = torch.ones(folds, dtype=torch.bool, device=dev)
        keep[f] = False
        tr_s, tr_l = s32[keep].reshape(-1), lab[keep].reshape(-1)
        smin, smax = tr_s.min(), tr_s.max()
        step = ((smax - smin) / num_thresholds)                       # float32, like numpy's float32 / int
        thr = smin.to(torch.float64) + step.to(torch.float64) * steps  # float64 thresholds
        pos, neg = tr_s[tr_l].to(torch.float64), tr_s[~tr_l].to(torch.float64)
        tpr = (pos[None, :] > thr[:, None]).sum(dim=1).to(torch.float64) / pos.numel()
        fpr = (neg[None, :] > thr[:, None]).sum(dim=1).to(torch.float64) / neg.numel()
        best = thr[torch.argmax(tpr - fpr)]
        te_s, te_l = s32[f].to(torch.float64), lab[f]
        correct = (te_s[te_l] > best).sum() + (te_s[~te_l] < best).sum()
        accs.append(correct.to(torch.float64) / per)
    acc = torch.stack(accs)
    mean = acc.mean()
    std = acc.std(unbiased=True) / (folds ** 0.5)
    return float(mean.item()), float(std.item())

