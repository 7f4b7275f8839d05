"""Debug helper: reruns the planted-identities scenario and dumps the workspace state of overflowed rows."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import retrieval as R
from visiondk_b200.retrieval import FlatIPIndex

g, q, labels = R.synthetic_gallery(n_ids=300, per_id=100, dim=512, seed=2)
rng = np.random.default_rng(9)
g_raw = g * rng.uniform(0.5, 4.0, (g.shape[0], 1)).astype(np.float32)
q_raw = q * rng.uniform(0.5, 4.0, (q.shape[0], 1)).astype(np.float32)
idx = FlatIPIndex(512, "cuda", normalize=True)
idx.add(g_raw)
s, i = idx.search(q_raw, 100)
print("status", idx.last_status.cpu().tolist())
nq, segs, cc = 300, 8192, 2048
ws = idx._ws.cpu().numpy()
a256 = lambda v: (v + 255) & ~255
off = 0
seg = ws[off:off + nq * segs * 8].view(np.uint32).reshape(nq, segs, 2); off += a256(nq * segs * 8)
carry = []
for _ in range(2):
    carry.append(ws[off:off + nq * cc * 8].view(np.uint32).reshape(nq, cc, 2)); off += a256(nq * cc * 8)
seg_cnt = ws[off:off + nq * 32 * 4].view(np.uint32).reshape(nq, 32); off += a256(nq * 32 * 4)
carry_cnt = ws[off:off + nq * 4].view(np.uint32); off += a256(nq * 4)
tau = ws[off:off + nq * 4].view(np.float32); off += a256(nq * 4)
eps = ws[off:off + nq * 4].view(np.float32)
print("seg_cnt max per row (top 5):", np.sort(seg_cnt.max(1))[-5:], "rows", np.argsort(seg_cnt.max(1))[-5:])
bad = np.nonzero(seg_cnt.max(1) > 315)[0]
for r in bad[:4]:
    print("row", r, "seg_cnt", seg_cnt[r].tolist(), "tau", tau[r], "eps", eps[r], "carry_cnt", carry_cnt[r])
print("tau stats", np.nanmin(tau), np.nanmax(tau), "nan", np.isnan(tau).sum(), "eps", eps.min(), eps.max())
print("carry_cnt", carry_cnt.min(), carry_cnt.max())
ref_s, ref_i = R.flat_ip_search_candidates(R.l2_normalize(q_raw), R.l2_normalize(g_raw), 100)
print("rows differing", np.nonzero((i != ref_i).any(1))[0][:10])
