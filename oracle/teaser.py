"""Oracle (test infrastructure, NumPy float64) for csrc/teaser.cu: the TEASER++ solve of registration_node.py:91-131.

PARITY UNPINNED: teaserpp_python is not installable offline and TEASER++ is not in the reference tree (Dockerfile:74 clones its
HEAD); this restates the published algorithm (Yang, Shi, Carlone, T-RO 2020; teaser/registration.cc as recalled): TIM
compatibility graph with beta = 2 noise_bound sqrt(cbar2), maximum clique, GNC-TLS rotation on the clique's chain of TIMs,
component-wise TLS translation by adaptive voting.  Independent of the product code: the Kabsch step is np.linalg.svd."""
from __future__ import annotations

import numpy as np


def tim_graph(src: np.ndarray, tgt: np.ndarray, noise_bound: float, cbar2: float = 1.0) -> np.ndarray:
    """(K, K) bool: | ||b_i - b_j|| - ||a_i - a_j|| | <= beta (teaser ScaleInliersSelector), no self loops."""
    da = src[:, None, :] - src[None, :, :]
    db = tgt[:, None, :] - tgt[None, :, :]
    na = np.sqrt((da[..., 0] * da[..., 0] + da[..., 1] * da[..., 1]) + da[..., 2] * da[..., 2])
    nb = np.sqrt((db[..., 0] * db[..., 0] + db[..., 1] * db[..., 1]) + db[..., 2] * db[..., 2])
    g = np.abs(na - nb) <= 2.0 * noise_bound * np.sqrt(cbar2)
    np.fill_diagonal(g, False)
    return g


def max_clique(adj: np.ndarray) -> np.ndarray:
    """A maximum clique (ascending vertex indices): Bron-Kerbosch with pivoting over Python sets -- an algorithm unrelated to
    the product's coloured branch and bound.  Maximum cliques need not be unique: tests compare SIZES, and vertex sets only
    where the planted clique is the unique maximum."""
    n = adj.shape[0]
    nbr = [set(np.nonzero(adj[i])[0].tolist()) for i in range(n)]
    best: list = []

    def bk(r, p, x):
        nonlocal best
        if not p and not x:
            if len(r) > len(best):
                best = sorted(r)
            return
        if len(r) + len(p) <= len(best):
            return
        u = max(p | x, key=lambda v: len(nbr[v] & p))
        for v in list(p - nbr[u]):
            bk(r + [v], p & nbr[v], x & nbr[v])
            p = p - {v}
            x = x | {v}
    bk([], set(range(n)), set())
    return np.asarray(best, dtype=np.int64)


def kabsch_weighted(a: np.ndarray, b: np.ndarray, w: np.ndarray) -> np.ndarray:
    """R minimising sum w ||b - R a||^2 (teaser svdRot: H = X W Y^T, R = V U^T with the reflection fix)."""
    h = (a * w[:, None]).T @ b
    u, _, vt = np.linalg.svd(h)
    v = vt.T
    if np.linalg.det(u) * np.linalg.det(v) < 0:
        v[:, 2] *= -1
    return v @ u.T


def gnc_tls_rotation(a, b, noise_bound, gnc_factor=1.4, max_iterations=10000, cost_threshold=1e-16):
    n = a.shape[0]
    w = np.ones(n)
    nb2 = noise_bound ** 2
    mu, prev = 1.0, np.inf
    rot = np.eye(3)
    it = 0
    for it in range(max_iterations):
        rot = kabsch_weighted(a, b, w)
        r2 = ((b - a @ rot.T) ** 2).sum(axis=1)
        if it == 0:
            mu = 1.0 / (2.0 * r2.max() / nb2 - 1.0)
            if mu <= 0:
                break
        th1, th2 = (mu + 1) / mu * nb2, mu / (mu + 1) * nb2
        cost = float((w * r2).sum())
        mid = np.sqrt(nb2 * mu * (mu + 1) / np.maximum(r2, 1e-300)) - mu
        w = np.where(r2 >= th1, 0.0, np.where(r2 <= th2, 1.0, mid))
        diff = abs(cost - prev)
        mu *= gnc_factor
        prev = cost
        if diff < cost_threshold:
            break
    return rot, w >= 0.5


def tls_scalar(x: np.ndarray, rng: float) -> float:
    """teaser ScalarTLSEstimator::estimate with equal ranges: the consensus interval with the lowest truncated cost."""
    n = len(x)
    ev = sorted([(x[i] - rng, i + 1) for i in range(n)] + [(x[i] + rng, -(i + 1)) for i in range(n)], key=lambda e: e[0])
    wgt = 1.0 / rng ** 2
    ris, dxw, dw, sx, sx2, card = rng * n, 0.0, 0.0, 0.0, 0.0, 0
    best_cost, best = np.inf, 0.0
    for _, tag in ev:
        i, eps = abs(tag) - 1, (1.0 if tag > 0 else -1.0)
        card += 1 if tag > 0 else -1
        dw += eps * wgt
        dxw += eps * wgt * x[i]
        ris -= eps * rng
        sx += eps * x[i]
        sx2 += eps * x[i] * x[i]
        if card <= 0:
            continue
        xh = dxw / dw
        cost = (card * xh * xh + sx2 - 2 * sx * xh) + ris
        if cost < best_cost:
            best_cost, best = cost, xh
    return best


def teaser_solve(src, tgt, noise_bound=0.2, cbar2=1.0, gnc_factor=1.4, max_iterations=10000, cost_threshold=1e-16, clique=None):
    src, tgt = np.asarray(src, dtype=np.float64), np.asarray(tgt, dtype=np.float64)
    T = np.eye(4)
    if len(src) < 2:
        return T, np.zeros(0, dtype=np.int64)
    cl = max_clique(tim_graph(src, tgt, noise_bound, cbar2)) if clique is None else np.asarray(clique)
    if len(cl) <= 1:
        return T, cl
    a, b = np.diff(src[cl], axis=0), np.diff(tgt[cl], axis=0)     # CHAIN
    rot, _ = gnc_tls_rotation(a, b, noise_bound, gnc_factor, max_iterations, cost_threshold)
    raw = tgt[cl] - src[cl] @ rot.T
    T[:3, :3] = rot
    T[:3, 3] = [tls_scalar(raw[:, c], noise_bound * np.sqrt(cbar2)) for c in range(3)]
    return T, cl
