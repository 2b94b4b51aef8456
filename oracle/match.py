"""Oracle (test infrastructure): descriptor nearest-neighbour matching, float64 NumPy.

Follows (paths relative to /root/reference):
  * ``src/kiss-icp/cpp/kiss_icp/core/VoxelHashMap.cpp:461-626``  GetVFMCorrespondences:
    per-row float32 ``fvec_renorm_L2`` (:469-482), ``IndexFlatIP.search(k=1)`` (:486-495),
    reject ``D < min_cosine`` (:501-511), emit xyz pairs of the valid rows in query order
    (:587-600).  faiss itself is third-party and absent (unpinned git HEAD,
    ``Dockerfile:54-62``) -> its published semantics are restated: inner-product top-1,
    first (lowest) index wins exact ties.
  * ``src/vfm-reg/src/registration_node.py:482-538``  find_correspondences (mutual NN).
  * ``src/vfm-reg/src/registration_node.py:191-214``  brute-force
    ``sqrt(2 - 2 a.b + 1e-6)`` + argmin block of the PointDSC path.

The "truth" here is the float64 evaluation of the inner products of the float32
L2-renormalised rows; ``gap`` (top-1 minus top-2) lets the parity tests separate
unambiguous queries (gap > 1e-5) from near-ties.
"""
from __future__ import annotations

import numpy as np


def renorm_l2(x: np.ndarray) -> np.ndarray:
    """faiss ``fvec_renorm_L2`` (VoxelHashMap.cpp:474,480): float32 rows scaled by
    1/sqrt(sum x^2) when the sum is > 0; all-zero rows stay zero."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    s = np.einsum("ij,ij->i", x, x, dtype=np.float32)
    inv = np.ones_like(s)
    nz = s > 0
    inv[nz] = (1.0 / np.sqrt(s[nz], dtype=np.float32)).astype(np.float32)
    return (x * inv[:, None]).astype(np.float32)


def top2_ip(a: np.ndarray, b: np.ndarray, block: int = 512):
    """For every row of ``a``: argmax_j <a_i, b_j> in float64 (lowest j on exact ties),
    the top-1 value and the top-2 value (-inf when b has a single row).

    Blocked over queries like the reference's own brute-force path
    (registration_node.py:191-209)."""
    a64 = np.asarray(a, dtype=np.float64)
    b64t = np.asarray(b, dtype=np.float64).T.copy()
    n, m = a64.shape[0], b64t.shape[1]
    idx = np.zeros(n, dtype=np.int64)
    best = np.full(n, -np.inf)
    second = np.full(n, -np.inf)
    if m == 0:
        return idx - 1, best, second
    for s in range(0, n, block):
        ip = a64[s:s + block] @ b64t
        i1 = np.argmax(ip, axis=1)  # first occurrence == lowest index on ties
        rows = np.arange(ip.shape[0])
        v1 = ip[rows, i1]
        if m > 1:
            ip[rows, i1] = -np.inf
            v2 = ip.max(axis=1)
        else:
            v2 = np.full(ip.shape[0], -np.inf)
        idx[s:s + block] = i1
        best[s:s + block] = v1
        second[s:s + block] = v2
    return idx, best, second


def match_nn(a: np.ndarray, b: np.ndarray, normalize: bool = True, mutual: bool = False):
    """Top-1 inner-product search a->b (and b->a when ``mutual``).

    Returns dict(idx01, sim01, sec01[, idx10, sim10, sec10]); ``sec*`` is the runner-up
    similarity (used by the ratio test and for the ambiguity report)."""
    af = renorm_l2(a) if normalize else np.ascontiguousarray(a, dtype=np.float32)
    bf = renorm_l2(b) if normalize else np.ascontiguousarray(b, dtype=np.float32)
    out = {}
    out["idx01"], out["sim01"], out["sec01"] = top2_ip(af, bf)
    if mutual:
        out["idx10"], out["sim10"], out["sec10"] = top2_ip(bf, af)
    return out


def filter_correspondences(idx01, sim01, sec01=None, idx10=None, *, min_cos=None,
                           mutual=False, ratio=None):
    """Correspondence list (K, 2) int32 = (query index, map index), in query order.

      * cosine gate: keep iff sim >= min_cos   (VoxelHashMap.cpp:503 rejects ``D < min``)
      * mutual:      keep iff idx10[idx01[i]] == i   (registration_node.py:530)
      * ratio:       Lowe test on squared L2 distances of unit vectors,
                     (1 - s1) < ratio^2 (1 - s2)   (no reference counterpart)
    """
    idx01 = np.asarray(idx01)
    n = idx01.shape[0]
    keep = idx01 >= 0
    if min_cos is not None:
        keep &= np.asarray(sim01, dtype=np.float32) >= np.float32(min_cos)
    if mutual:
        idx10 = np.asarray(idx10)
        keep &= idx10[np.clip(idx01, 0, None)] == np.arange(n)
    if ratio is not None:
        s1 = np.asarray(sim01, dtype=np.float32)
        s2 = np.asarray(sec01, dtype=np.float32)
        r2 = np.float32(ratio) * np.float32(ratio)
        keep &= (np.float32(1) - s1) < r2 * (np.float32(1) - s2)
    q = np.nonzero(keep)[0]
    return np.stack([q, idx01[q]], axis=1).astype(np.int32)


def get_vfm_correspondences(points: np.ndarray, map_points: np.ndarray, min_cosine: float):
    """Restates VoxelHashMap::GetVFMCorrespondences (VoxelHashMap.cpp:461-626) on the
    reference's own array layout: rows are ``[x, y, z, f_0 .. f_{D-1}]``.

    Returns (src_xyz[K,3] f64, tgt_xyz[K,3] f64) for the rows whose top-1 cosine is
    >= ``min_cosine``; the MAD filter of :546-584 is computed-but-disabled in the
    reference (body commented out) and is therefore a no-op here."""
    points = np.asarray(points)
    map_points = np.asarray(map_points)
    r = match_nn(points[:, 3:].astype(np.float32), map_points[:, 3:].astype(np.float32))
    corr = filter_correspondences(r["idx01"], r["sim01"], min_cos=min_cosine)
    src = points[corr[:, 0], :3].astype(np.float64)
    tgt = map_points[corr[:, 1], :3].astype(np.float64)
    return src, tgt


def find_correspondences(feats0: np.ndarray, feats1: np.ndarray, n_points: int = 5000,
                         mutual_filter: bool = True):
    """Restates the nested ``find_correspondences`` (registration_node.py:482-538) with a
    brute-force float64 L2 search instead of cKDTree (identical results: exact 1-NN)."""
    f0 = np.asarray(feats0, dtype=np.float64)
    f1 = np.asarray(feats1, dtype=np.float64)

    def knn(x, y):
        d2 = (x * x).sum(1)[:, None] + (y * y).sum(1)[None, :] - 2.0 * (x @ y.T)
        i = np.argmin(d2, axis=1)
        return i, np.sqrt(np.maximum(d2[np.arange(len(x)), i], 0.0))

    nns01, dists = knn(f0, f1)
    idx0 = np.arange(len(nns01))
    if not mutual_filter:
        n = min(n_points, len(dists) - 1)
        top = np.argpartition(dists, n)[:n]
        return idx0[top], nns01[top]
    nns10, _ = knn(f1, f0)
    m = nns10[nns01] == idx0
    return idx0[m], nns01[m]


def l2_block_argmin(src: np.ndarray, tgt: np.ndarray, batch: int = 1000):
    """Restates the PointDSC-path brute-force block (registration_node.py:191-209):
    ``sqrt(2 - 2 s.t + 1e-6)``, per-row argmin and min."""
    idx, dis = [], []
    for s in range(0, src.shape[0], batch):
        d = np.sqrt(2 - 2 * (src[s:s + batch] @ tgt.T) + 1e-6)
        idx.append(np.argmin(d, axis=1))
        dis.append(np.min(d, axis=1))
    return np.concatenate(idx), np.concatenate(dis)
