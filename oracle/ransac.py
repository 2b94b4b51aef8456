"""Oracle (test infrastructure): correspondence RANSAC, float64 NumPy.

Follows the reference's call site ``src/vfm-reg/src/registration_node.py:319-327``
(``registration_ransac_based_on_correspondence(src, tgt, corres, 10000,
TransformationEstimationPointToPoint(False), ransac_n=3, RANSACConvergenceCriteria(50000, 1))``).
The arithmetic lives in Open3D 0.18.0 (``Dockerfile:81``) / Eigen ``umeyama`` -- third
party, absent -> **parity unpinned**; the published algorithm is restated:

  * draw 3 correspondences with replacement; rigid fit without scale:
    means, Sigma = sum (q - qm)(p - pm)^T, Sigma = U diag(s) V^T,
    S = diag(1, 1, sign(det U det V)), R = U S V^T, t = qm - R pm;
  * score over the correspondence list: inlier iff ||R p + t - q||^2 < tau^2 (strict);
  * best = max inlier count, then min rmse, then lowest hypothesis id; no final refit
    (the winning 3-point transform is returned as is), optional ``refit`` over the inliers.

The Kabsch sign rule is cross-checked against the reference's only in-tree Kabsch,
``src/vfm-reg/src/pointdsc/common.py:7-47`` (tests/golden/kabsch_pointdsc.npz).
"""
from __future__ import annotations

import numpy as np

DEGENERATE_REL = 1e-12   # lambda_2 <= DEGENERATE_REL * lambda_1  (eigenvalues of Sigma^T Sigma)
DEGENERATE_ABS = 1e-300
SUMQ_BITS = 40           # d^2 is accumulated as rint(d^2 * 2^40 / tau^2) (order-independent)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def sample_indices(seed: int, n_hyp: int, n_corr: int) -> np.ndarray:
    """Counter-based sampler shared by the CUDA path (csrc/ransac.cu) and the C oracle:
    splitmix64 finaliser of ``seed * GOLDEN + (3 h + j)``, range-reduced by multiply-high.
    (Open3D's own sampler is a mutex-protected mt19937 drawn from OpenMP threads and is
    not reproducible)"""
    if n_corr <= 0:
        return np.zeros((n_hyp, 3), dtype=np.int32)
    with np.errstate(over="ignore"):
        ctr = np.arange(n_hyp * 3, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) * np.uint64(0x9E3779B97F4A7C15) + ctr
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        hi = z >> np.uint64(32)
        idx = (hi * np.uint64(n_corr)) >> np.uint64(32)
    return idx.astype(np.int32).reshape(n_hyp, 3)


def kabsch(p: np.ndarray, q: np.ndarray):
    """Least-squares rigid transform q ~ R p + t (Umeyama without scale, A.3).
    Returns (R, t, valid); ``valid`` is False for rank-deficient inputs (collinear or
    repeated points) where the rotation is not determined."""
    p = np.asarray(p, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    pm, qm = p.mean(0), q.mean(0)
    sigma = (q - qm).T @ (p - pm)
    u, s, vt = np.linalg.svd(sigma)
    valid = bool(s[0] ** 2 > DEGENERATE_ABS and s[1] ** 2 > DEGENERATE_REL * s[0] ** 2)
    d = np.sign(np.linalg.det(u) * np.linalg.det(vt))
    if d == 0:
        d = 1.0
    r = u @ np.diag([1.0, 1.0, d]) @ vt
    t = qm - r @ pm
    return r, t, valid


def kabsch3_batch(p3: np.ndarray, q3: np.ndarray):
    """Batched 3-point Kabsch: p3, q3 (H, 3, 3) -> R (H,3,3), t (H,3), valid (H,)."""
    p3 = np.asarray(p3, dtype=np.float64)
    q3 = np.asarray(q3, dtype=np.float64)
    pm, qm = p3.mean(1), q3.mean(1)
    sigma = np.einsum("hki,hkj->hij", q3 - qm[:, None], p3 - pm[:, None])
    u, s, vt = np.linalg.svd(sigma)
    valid = (s[:, 0] ** 2 > DEGENERATE_ABS) & (s[:, 1] ** 2 > DEGENERATE_REL * s[:, 0] ** 2)
    d = np.sign(np.linalg.det(u) * np.linalg.det(vt))
    d[d == 0] = 1.0
    smat = np.zeros_like(sigma)
    smat[:, 0, 0] = 1.0
    smat[:, 1, 1] = 1.0
    smat[:, 2, 2] = d
    r = u @ smat @ vt
    t = qm - np.einsum("hij,hj->hi", r, pm)
    return r, t, valid


def score(r, t, src_c, tgt_c, thresh: float, block: int = 256):
    """Per-hypothesis inlier count, quantised sum of squared inlier residuals, and the
    float64 sum (A.4).  r (H,3,3), t (H,3), src_c/tgt_c (K,3)."""
    h = r.shape[0]
    tau2 = float(thresh) * float(thresh)
    scale = np.float64(2.0 ** SUMQ_BITS) / np.float64(tau2)
    counts = np.zeros(h, dtype=np.int64)
    sumq = np.zeros(h, dtype=np.int64)
    sumsq = np.zeros(h, dtype=np.float64)
    for s in range(0, h, block):
        x = np.einsum("hij,kj->hki", r[s:s + block], src_c) + t[s:s + block, None, :] - tgt_c[None]
        d2 = (x * x).sum(-1)
        inl = d2 < tau2
        counts[s:s + block] = inl.sum(1)
        sumq[s:s + block] = np.where(inl, np.rint(d2 * scale), 0.0).astype(np.int64).sum(1)
        sumsq[s:s + block] = np.where(inl, d2, 0.0).sum(1)
    return counts, sumq, sumsq


def select_best(counts, sumq, valid):
    """max count, then min residual sum (== min rmse at equal count), then lowest id."""
    c = np.where(valid, counts, -1)
    cmax = c.max() if c.size else -1
    if cmax < 0:
        return -1
    cand = np.nonzero(c == cmax)[0]
    return int(cand[np.argmin(sumq[cand])])  # first occurrence == lowest id


def ransac(src_xyz, tgt_xyz, corr, sample_idx, thresh: float, refit: bool = False):
    """Full solve on a correspondence list.

    src_xyz (N,3), tgt_xyz (M,3), corr (K,2) int, sample_idx (H,3) int into ``corr``.
    Returns dict(T 4x4 f64, best, counts, sumq, valid, mask, fitness, rmse).
    K < 3 or no valid hypothesis -> identity transform, fitness 0 (Open3D returns an
    empty RegistrationResult for |corres| < ransac_n)."""
    src_xyz = np.asarray(src_xyz, dtype=np.float64)
    tgt_xyz = np.asarray(tgt_xyz, dtype=np.float64)
    corr = np.asarray(corr).reshape(-1, 2)
    k = corr.shape[0]
    h = sample_idx.shape[0]
    out = dict(T=np.eye(4), best=-1, counts=np.full(h, -1, dtype=np.int64),
               sumq=np.zeros(h, dtype=np.int64), valid=np.zeros(h, dtype=bool),
               mask=np.zeros(k, dtype=bool), fitness=0.0, rmse=0.0)
    if k < 3 or h == 0:
        return out
    src_c = src_xyz[corr[:, 0]]
    tgt_c = tgt_xyz[corr[:, 1]]
    r, t, valid = kabsch3_batch(src_c[sample_idx], tgt_c[sample_idx])
    counts, sumq, sumsq = score(r, t, src_c, tgt_c, thresh)
    best = select_best(counts, sumq, valid)
    out.update(counts=np.where(valid, counts, -1), sumq=np.where(valid, sumq, 0), valid=valid, best=best)
    if best < 0:
        return out
    rb, tb = r[best], t[best]
    x = src_c @ rb.T + tb - tgt_c
    d2 = (x * x).sum(1)
    mask = d2 < float(thresh) ** 2
    if refit and mask.sum() >= 3:
        r2, t2, ok = kabsch(src_c[mask], tgt_c[mask])
        if ok:
            rb, tb = r2, t2
    tm = np.eye(4)
    tm[:3, :3] = rb
    tm[:3, 3] = tb
    n_in = int(mask.sum())
    out.update(T=tm, mask=mask, fitness=n_in / k,
               rmse=float(np.sqrt(d2[mask].sum() / n_in)) if n_in else 0.0)
    return out


def ransac_nn_all(src_xyz, tgt_xyz, corr, sample_idx, max_dist: float, fit=None):
    """The hypothesis score of Open3D 0.18's ``registration_ransac_based_on_correspondence`` as recalled in SURVEY.md A.8
    (call site registration_node.py:312-327; Open3D's source is not in the reference tree -> **parity unpinned**):
    every hypothesis transforms the whole source cloud; fitness = share of transformed points whose nearest target point
    (KD-tree query) is closer than ``max_dist``, inlier_rmse over those distances; best = higher fitness, then lower rmse
    (then lower hypothesis id); the winning 3-point transform is returned as is.

    ``fit(p3, q3) -> (R, t, valid)`` replaces the SVD-based 3-point fit (the tests pass the C restatement's so that the
    hypotheses are bit-identical to the CUDA path's)."""
    from scipy.spatial import cKDTree
    src_xyz = np.asarray(src_xyz, dtype=np.float64)
    tgt_xyz = np.asarray(tgt_xyz, dtype=np.float64)
    corr = np.asarray(corr).reshape(-1, 2)
    h = sample_idx.shape[0]
    out = dict(T=np.eye(4), best=-1, inliers=np.full(h, -1, dtype=np.int64), sum_d2=np.zeros(h), fitness=0.0, rmse=0.0)
    if corr.shape[0] < 3 or h == 0:
        return out
    src_c, tgt_c = src_xyz[corr[:, 0]], tgt_xyz[corr[:, 1]]
    if fit is None:
        r, t, valid = kabsch3_batch(src_c[sample_idx], tgt_c[sample_idx])
    else:
        fits = [fit(src_c[s], tgt_c[s]) for s in sample_idx]
        r, t, valid = np.stack([f[0] for f in fits]), np.stack([f[1] for f in fits]), np.array([f[2] for f in fits])
    tree = cKDTree(tgt_xyz)
    cnt = np.full(h, -1, dtype=np.int64)
    s2 = np.zeros(h)
    for i in range(h):
        if not valid[i]:
            continue
        x = src_xyz @ r[i].T + t[i]
        d, _ = tree.query(x, k=1)
        inl = d < max_dist
        cnt[i] = int(inl.sum())
        s2[i] = float((d[inl] ** 2).sum())
    out.update(inliers=cnt, sum_d2=s2)
    ok = np.nonzero(cnt > 0)[0]
    if ok.size == 0:
        return out
    order = np.lexsort((ok, s2[ok] / cnt[ok], -cnt[ok]))
    best = int(ok[order[0]])
    tm = np.eye(4)
    tm[:3, :3], tm[:3, 3] = r[best], t[best]
    out.update(T=tm, best=best, fitness=cnt[best] / src_xyz.shape[0], rmse=float(np.sqrt(s2[best] / cnt[best])))
    return out
