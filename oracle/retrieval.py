"""Oracle for the CBIR retrieval path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Reference behaviour restated here:
  * models/faceX/face_model.py:137-140   features = F.normalize(model(x)); L2, dim=1, eps=1e-12
  * engine/cbir/evaluation.py:155-168    faiss.index_factory(dim, "Flat", METRIC_INNER_PRODUCT); add()
  * engine/cbir/evaluation.py:190-195    index.search(q, k) -> (scores desc float32 [n,k], ids int64 [n,k], -1 pad)
faiss itself is a third-party dependency (faiss-gpu 1.8.0, README.md:34) that is NOT vendored under
/root/reference and not installed here: PARITY UNPINNED at that boundary.  Flat/IP is exact brute force, so
its published algorithm is "all inner products, keep the k largest"; faiss leaves tie order and the
floating-point summation order unspecified (its GPU path even rounds inputs to fp16), so "bit-exact" needs a
canonical definition.  Ours:

  canonical_score(q, g) = float32( fp64 sum of q_i*g_i in a FIXED order )
      order: 32 partial sums, partial l takes i = l, l+32, l+64, ... sequentially; then the partials are
      combined by an xor butterfly (16, 8, 4, 2, 1).  Products of two float32 are exact in float64.
  ranking = score descending, then id ascending;   padding = (-FLT_MAX, -1) like faiss' heap neutral.

The CUDA path must reproduce ids AND scores bit for bit.
"""
from __future__ import annotations

import numpy as np

FLT_LOWEST = np.float32(-3.4028234663852886e38)


def _butterfly(acc: np.ndarray) -> np.ndarray:
    """acc[..., 32] float64 -> xor-butterfly sum, every lane ends with the same value; returns [...]."""
    idx = np.arange(32)
    for off in (16, 8, 4, 2, 1):
        acc = acc + acc[..., idx ^ off]
    return acc[..., 0]


def _lane_sums(prod: np.ndarray) -> np.ndarray:
    """prod[..., D] float64 -> [..., 32] sequential per-lane sums (lane l: l, l+32, ...)."""
    d = prod.shape[-1]
    pad = (-d) % 32
    if pad:
        prod = np.concatenate([prod, np.zeros(prod.shape[:-1] + (pad,), prod.dtype)], axis=-1)
    p = prod.reshape(prod.shape[:-1] + (-1, 32))
    acc = np.zeros(prod.shape[:-1] + (32,), np.float64)
    for j in range(p.shape[-2]):
        acc = acc + p[..., j, :]
    return acc


def canonical_dot(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Row-wise canonical score of float32 arrays a[..., D], b[..., D] (broadcasting) -> float32[...]."""
    prod = a.astype(np.float64) * b.astype(np.float64)
    return _butterfly(_lane_sums(prod)).astype(np.float32)


def l2_normalize(x: np.ndarray, eps: float = 1e-12) -> np.ndarray:
    """F.normalize(x, p=2, dim=1, eps) with the canonical reduction (face_model.py:139).

    ||x|| = float32(sqrt(fp64 fixed-order sum of squares)); out = x / max(||x||, eps) in float32.
    torch's own reduction order is unspecified; results agree with it to 2 ulp (1 in the norm, 1 in the division).
    """
    x = np.ascontiguousarray(x, dtype=np.float32)
    ss = _butterfly(_lane_sums(x.astype(np.float64) ** 2))
    nrm = np.sqrt(ss).astype(np.float32)
    den = np.maximum(nrm, np.float32(eps))
    return (x / den[:, None]).astype(np.float32)


def canonical_scores(q: np.ndarray, g: np.ndarray, block: int = 64) -> np.ndarray:
    """All-pairs canonical scores float32 [nq, ng] (small sizes only: O(nq*ng*D) float64 work)."""
    q = np.ascontiguousarray(q, np.float32)
    g = np.ascontiguousarray(g, np.float32)
    out = np.empty((q.shape[0], g.shape[0]), np.float32)
    g64 = g.astype(np.float64)
    for i in range(0, q.shape[0], block):
        qb = q[i:i + block].astype(np.float64)
        prod = qb[:, None, :] * g64[None, :, :]
        out[i:i + block] = _butterfly(_lane_sums(prod)).astype(np.float32)
    return out


def topk_from_scores(scores: np.ndarray, k: int, id_offset: int = 0):
    """(score desc, id asc) top-k of a float32 score matrix; pads with (-FLT_MAX, -1)."""
    nq, ng = scores.shape
    out_s = np.full((nq, k), FLT_LOWEST, np.float32)
    out_i = np.full((nq, k), -1, np.int64)
    kk = min(k, ng)
    if kk == 0:
        return out_s, out_i
    ids = np.arange(ng, dtype=np.int64)
    for r in range(nq):
        order = np.lexsort((ids, -scores[r].astype(np.float64)))[:kk]  # last key is primary
        out_s[r, :kk] = scores[r, order]
        out_i[r, :kk] = order + id_offset
    return out_s, out_i


def flat_ip_search(q: np.ndarray, g: np.ndarray, k: int, id_offset: int = 0):
    """IndexFlatIP.search restated (engine/cbir/evaluation.py:193) with the canonical score/tie rules."""
    return topk_from_scores(canonical_scores(q, g), k, id_offset)


def flat_ip_search_candidates(q: np.ndarray, g: np.ndarray, k: int, over: int = 4):
    """Same result as flat_ip_search at sizes where all-pairs fp64 is too slow: float32 BLAS scores pick
    `over*k + 64` candidates per query, which are then re-scored canonically.  The float32 GEMM error
    (<~1e-5 for unit rows) is orders of magnitude below the score gap at rank over*k, so the candidate set
    contains the true top-k; the function asserts the safety margin it relies on."""
    q = np.ascontiguousarray(q, np.float32)
    g = np.ascontiguousarray(g, np.float32)
    nq, ng = q.shape[0], g.shape[0]
    c = min(ng, over * k + 64)
    out_s = np.full((nq, k), FLT_LOWEST, np.float32)
    out_i = np.full((nq, k), -1, np.int64)
    kk = min(k, ng)
    for i in range(0, nq, 256):
        s32 = q[i:i + 256] @ g.T
        if c < ng:
            cand = np.argpartition(-s32, c - 1, axis=1)[:, :c]
        else:
            cand = np.tile(np.arange(ng), (s32.shape[0], 1))
        for r in range(s32.shape[0]):
            ids = np.sort(cand[r])
            sc = canonical_dot(q[i + r][None, :], g[ids])
            order = np.lexsort((ids, -sc.astype(np.float64)))[:kk]
            if c < ng and kk > 0:
                worst_in = s32[r, ids].min()
                assert sc[order[-1]] - worst_in > 1e-4 or kk == c, "candidate margin too thin for the fp32 prefilter"
            out_s[i + r, :kk] = sc[order]
            out_i[i + r, :kk] = ids[order]
    return out_s, out_i


def merge_topk(scores_list, ids_list, k: int):
    """Merge per-shard ordered lists with the same (score desc, id asc) rule (IndexShards semantics)."""
    s = np.concatenate(scores_list, axis=1)
    i = np.concatenate(ids_list, axis=1)
    nq = s.shape[0]
    out_s = np.full((nq, k), FLT_LOWEST, np.float32)
    out_i = np.full((nq, k), -1, np.int64)
    for r in range(nq):
        valid = i[r] >= 0
        sr, ir = s[r][valid], i[r][valid]
        order = np.lexsort((ir, -sr.astype(np.float64)))[:k]
        out_s[r, :len(order)] = sr[order]
        out_i[r, :len(order)] = ir[order]
    return out_s, out_i


def sketch_ranks(k: int, world: int):
    """Ranks a shard of a `world`-way search reports: k, ceil(k/2), ceil(k/4), ... until `world` shards of the smallest rank
    cover k (visiondk_b200.retrieval._Exchange; no counterpart in the reference, which replicates the index)."""
    ranks, r, w = [int(k)], int(k), 1
    while w < world and len(ranks) < 8 and r > 1:
        r, w = -(-r // 2), w * 2
        ranks.append(r)
    return ranks


def rank_sketch(shard_scores: np.ndarray, ranks) -> np.ndarray:
    """shard_scores [nq, n_rows] (this shard's scores, any order) -> [nq, len(ranks)]: the ranks[i]-th largest score of every
    query (-inf where the shard holds fewer rows): what `vdk_ip_topk_rank_sketch` lower-bounds (by eps, the fp16 pass' error)."""
    nq, n = shard_scores.shape
    desc = -np.sort(-shard_scores.astype(np.float64), axis=1)
    out = np.full((nq, len(ranks)), -np.inf, np.float32)
    for j, r in enumerate(ranks):
        if r <= n:
            out[:, j] = desc[:, r - 1]
    return out


def bound_from_sketches(sketches: np.ndarray, ranks, k: int, bound: np.ndarray) -> np.ndarray:
    """sketches [n_shards, nq, n_ranks] of shards holding DISJOINT rows -> max(bound, largest reported score t such that
    sum over shards of max{ranks[i] : sketch[s, q, i] >= t} >= k): a lower bound of the global k-th score
    (`vdk_topk_bound_from_sketches`).  Entries need not decrease with the rank index (a rank-k entry may carry an older global
    bound), hence the max over i."""
    out = np.array(bound, dtype=np.float32, copy=True)
    n_shards, nq, nr = sketches.shape
    for q in range(nq):
        for t in sketches[:, q, :].ravel():
            if not np.isfinite(t):
                continue
            cnt = 0
            for s_ in range(n_shards):
                ok = [ranks[i] for i in range(nr) if sketches[s_, q, i] >= t]
                cnt += max(ok) if ok else 0
            if cnt >= k:
                out[q] = max(out[q], t)
    return out


def synthetic_gallery(n_ids: int, per_id: int, dim: int = 512, seed: int = 2, noise: float = 0.5):
    """SURVEY.md §8(d) synthetic CBIR data: identity centres N(0,I) x members = centre + noise*N(0,I),
    L2-normalised float32; one query per identity built the same way.  Returns (gallery, queries, labels)."""
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((n_ids, dim)).astype(np.float32)
    g = np.repeat(centres, per_id, axis=0) + noise * rng.standard_normal((n_ids * per_id, dim)).astype(np.float32)
    qs = centres + noise * rng.standard_normal((n_ids, dim)).astype(np.float32)
    labels = np.repeat(np.arange(n_ids), per_id)
    return l2_normalize(g), l2_normalize(qs), labels
