"""GPU parity (run with -m gpu on the B200 box): the CUDA path, called through the C ABI, against the oracle on the
same seeded inputs.  Bit-exact against the canonical-order C restatement (indices, similarities, inlier counts,
masks, transforms); against the float64 NumPy oracle on unambiguous queries and within 1e-4 Frobenius on T."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import cref, match, ransac  # noqa: E402
from vfm_registration_b200 import synth  # noqa: E402


@pytest.fixture(scope="module")
def vfm():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import vfm_registration_b200 as v
    v.get_context(0)  # raises if libvfmreg_b200.so is missing: no fallback
    return v


def _cmp_match(g, c, o, key, n_expected_clear=0.9):
    idx = g["idx" + key].cpu().numpy()
    sim = g["sim" + key].cpu().numpy()
    sec = g["sec" + key].cpu().numpy()
    assert np.array_equal(idx, c["idx" + key]), f"idx{key} differs from the C oracle at {np.nonzero(idx != c['idx' + key])[0][:5]}"
    assert np.array_equal(sim, c["sim" + key]) and np.array_equal(sec, c["sec" + key])
    clear = (o["sim" + key] - o["sec" + key]) > 1e-5
    assert clear.mean() >= n_expected_clear
    assert np.array_equal(idx[clear], o["idx" + key][clear])
    assert np.abs(sim - o["sim" + key]).max() < 1e-5  # fp32 accumulation tolerance vs float64


@pytest.mark.parametrize("algo", ["simt", "tc"])
@pytest.mark.parametrize("n,m,d", [(1000, 3000, 384), (257, 1031, 100), (64, 129, 7), (1, 5, 384), (5, 1, 16), (300, 20000, 384),
                                   (4096, 4096, 384), (700, 900, 768)])
def test_match_nn_vs_oracles(vfm, n, m, d, algo):
    rng = np.random.default_rng(n * 7 + m)
    a = rng.standard_normal((n, d)).astype(np.float32)
    b = rng.standard_normal((m, d)).astype(np.float32)
    if n > 10:
        a[3] = 0
        b[min(m - 1, 9)] = b[2]  # duplicate map row -> exact tie
    g = vfm.match_nn(a, b, mutual=True, algo=algo)
    gd = {k: getattr(g, k) for k in ("idx01", "sim01", "sec01", "idx10", "sim10", "sec10")}
    c = cref.match_nn(a, b, mutual=True)
    o = match.match_nn(a, b, mutual=True)
    _cmp_match(gd, c, o, "01", 0.9 if m > 1 else 0.0)
    _cmp_match(gd, c, o, "10", 0.9 if n > 1 else 0.0)


def test_match_tc_pathological_ties(vfm):
    """All database rows identical (every score ties) + zero rows: the candidate lists overflow and the exact
    fallback scan must still return the lowest index and the tied runner-up."""
    rng = np.random.default_rng(8)
    a = rng.standard_normal((300, 128)).astype(np.float32)
    a[::7] = 0
    b = np.tile(rng.standard_normal((1, 128)).astype(np.float32), (2000, 1))
    b[1500:] = rng.standard_normal((500, 128)).astype(np.float32)
    b[100] = 0
    g = vfm.match_nn(a, b, mutual=True, algo="tc")
    c = cref.match_nn(a, b, mutual=True)
    for k in ("idx01", "sim01", "sec01", "idx10", "sim10", "sec10"):
        assert np.array_equal(getattr(g, k).cpu().numpy(), c[k]), k


def test_match_tc_equals_simt_full_size(vfm):
    """BASELINE configs[1] size (10k x 50k x 384): the tensor-core path returns bit-identical indices and values to
    the exact fp32 kernel (itself pinned to the oracle at smaller sizes); planted matches are recovered."""
    s = synth.make_pair(2, 50_000, 10_000, 384)
    a, b = torch.from_numpy(s["scan_feat"]).cuda(), torch.from_numpy(s["map_feat"]).cuda()
    g = vfm.match_nn(a, b, mutual=True, algo="tc")
    e = vfm.match_nn(a, b, mutual=True, algo="simt")
    for k in ("idx01", "sim01", "sec01", "idx10", "sim10", "sec10"):
        assert torch.equal(getattr(g, k), getattr(e, k)), k
    inl = np.nonzero(s["perm"] >= 0)[0]
    assert np.array_equal(g.idx01.cpu().numpy()[inl], s["perm"][inl])


def test_match_no_normalize(vfm):
    rng = np.random.default_rng(5)
    a = rng.standard_normal((200, 64)).astype(np.float32) * 3
    b = rng.standard_normal((300, 64)).astype(np.float32) * 0.5
    g = vfm.match_nn(a, b, normalize=False)
    c = cref.match_nn(a, b, normalize=False)
    assert np.array_equal(g.idx01.cpu().numpy(), c["idx01"]) and np.array_equal(g.sim01.cpu().numpy(), c["sim01"])


def test_filter_correspondences(vfm):
    s = synth.make_pair(21, 4000, 1500, 128)
    g = vfm.match_nn(s["scan_feat"], s["map_feat"], mutual=True)
    c = cref.match_nn(s["scan_feat"], s["map_feat"], mutual=True)
    for kw in (dict(min_cos=0.8), dict(mutual=True), dict(min_cos=0.5, mutual=True), dict(ratio=0.9), dict(),
               dict(min_cos=0.8, ratio=0.8, mutual=True), dict(min_cos=1.1)):
        corr = vfm.filter_correspondences(g, **kw).cpu().numpy()
        want = match.filter_correspondences(c["idx01"], c["sim01"], c["sec01"], c["idx10"], **kw)
        assert np.array_equal(corr, want), kw
    inl = np.nonzero(s["perm"] >= 0)[0]
    corr = vfm.filter_correspondences(g, min_cos=0.8).cpu().numpy()
    assert np.array_equal(corr[:, 0], inl) and np.array_equal(corr[:, 1], s["perm"][inl])


def _corr_for(s, k, n_good, rng):
    good = np.nonzero(s["perm"] >= 0)[0]
    bad = np.nonzero(s["perm"] < 0)[0]
    q = np.sort(np.concatenate([good[:n_good], bad[: k - n_good]]))
    j = np.where(s["perm"][q] >= 0, s["perm"][q], rng.integers(0, len(s["map_xyz"]), len(q)))
    return np.stack([q, j], 1).astype(np.int32)


@pytest.mark.parametrize("thresh,refit,f64", [(1.0, False, False), (1e4, False, False), (1.0, True, True), (0.5, False, True)])
def test_ransac_bit_exact_vs_c_oracle(vfm, thresh, refit, f64):
    s = synth.make_pair(31, 6000, 2500, 16)
    rng = np.random.default_rng(9)
    corr = _corr_for(s, 1500, 400, rng)
    si = ransac.sample_indices(3, 8192, len(corr))
    si[5] = [7, 7, 9]      # repeated sample -> degenerate hypothesis
    si[6] = [1, 1, 1]
    src = s["scan_xyz"].astype(np.float64) if f64 else s["scan_xyz"]
    tgt = s["map_xyz"].astype(np.float64) if f64 else s["map_xyz"]
    g = vfm.ransac_kabsch(src, tgt, corr, sample_idx=si, thresh=thresh, refit=refit)
    c = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, si, thresh, refit=refit)
    assert np.array_equal(g.counts.cpu().numpy(), c["counts"])
    assert np.array_equal(g.sumq.cpu().numpy(), c["sumq"])
    assert g.counts[5].item() == -1 and g.counts[6].item() == -1
    assert g.best == c["best"] and g.n_inliers == c["n_inliers"]
    assert np.array_equal(g.mask.cpu().numpy(), c["mask"])
    if refit:
        assert np.linalg.norm(g.T - c["T"]) < 1e-9   # refit sums are reduced in a different order on the device
    else:
        assert np.array_equal(g.T, c["T"])           # bit-exact 4x4
    o = ransac.ransac(s["scan_xyz"], s["map_xyz"], corr, si, thresh, refit=refit)
    assert np.linalg.norm(g.T - o["T"]) < 1e-4 and g.best == o["best"] and np.array_equal(g.mask.cpu().numpy(), o["mask"])
    if thresh <= 1.0:
        rte, rre = synth.pose_errors(g.T, s["T_gt"])
        assert rte < 0.5 and rre < 1.0


def test_ransac_device_rng_and_small_k(vfm):
    s = synth.make_pair(33, 2000, 900, 16)
    rng = np.random.default_rng(1)
    corr = _corr_for(s, 500, 200, rng)
    g = vfm.ransac_kabsch(s["scan_xyz"], s["map_xyz"], corr, n_hyp=3000, seed=123, thresh=1.0)
    c = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, None, 1.0, seed=123, n_hyp=3000)
    assert g.best == c["best"] and np.array_equal(g.T, c["T"]) and np.array_equal(g.counts.cpu().numpy(), c["counts"])
    for k in (0, 1, 2):
        g = vfm.ransac_kabsch(s["scan_xyz"], s["map_xyz"], corr[:k], n_hyp=64, thresh=1.0)
        assert g.best == -1 and np.array_equal(g.T, np.eye(4)) and g.fitness == 0.0


@pytest.mark.parametrize("host", [True, False])
@pytest.mark.parametrize("kw", [dict(min_cos=0.8, inlier_thresh=1.0), dict(min_cos=0.8), dict(min_cos=None, mutual=True, inlier_thresh=1.0),
                                dict(min_cos=0.3, ratio=0.9, mutual=True, inlier_thresh=1.0, refit=True)])
def test_register_end_to_end(vfm, host, kw):
    """Config-1 style pair (smaller): whole path through vfmreg_register[_host] vs the oracle chain."""
    s = synth.make_pair(1, 4096, 2048, 384)
    h = 1024
    args = (s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"])
    if not host:
        args = tuple(torch.from_numpy(x).cuda() for x in args)
    r = vfm.register(*args, ransac_iters=h, seed=42, **kw)
    m = cref.match_nn(s["scan_feat"], s["map_feat"], mutual=True)
    corr = match.filter_correspondences(m["idx01"], m["sim01"], m["sec01"], m["idx10"], min_cos=kw.get("min_cos"),
                                        mutual=kw.get("mutual", False), ratio=kw.get("ratio"))
    assert np.array_equal(r.corr, corr)
    c = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, None, kw.get("inlier_thresh", 1e4), refit=kw.get("refit", False),
                    seed=42, n_hyp=h)
    assert r.best_hyp == c["best"] and np.array_equal(r.inlier_mask, c["mask"]) and r.n_inliers == c["n_inliers"]
    assert np.linalg.norm(r.T - c["T"]) < (1e-9 if kw.get("refit") else 1e-300)
    assert abs(r.fitness - c["fitness"]) < 1e-12 and abs(r.rmse - c["rmse"]) < 1e-9
    if kw.get("inlier_thresh", 1e4) <= 1.0:
        rte, rre = synth.pose_errors(r.T, s["T_gt"])
        assert rte < 1.0 and rre < 5.0   # recall@(1 m, 5 deg)


def test_register_sample_idx_and_errors(vfm):
    s = synth.make_pair(2, 1500, 600, 64)
    m = cref.match_nn(s["scan_feat"], s["map_feat"])
    corr = match.filter_correspondences(m["idx01"], m["sim01"], min_cos=0.8)
    si = np.random.default_rng(4).integers(0, len(corr), (2048, 3)).astype(np.int32)
    r = vfm.register(s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"], sample_idx=si, inlier_thresh=1.0)
    c = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, si, 1.0)
    assert r.best_hyp == c["best"] and np.array_equal(r.T, c["T"])
    # all rejected -> K = 0 -> identity, fitness 0 (never UB on empty correspondences)
    r0 = vfm.register(s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"], min_cos=1.1, ransac_iters=128)
    assert np.array_equal(r0.T, np.eye(4)) and r0.fitness == 0.0 and len(r0.corr) == 0 and r0.best_hyp == -1
    with pytest.raises(ValueError, match="Invalid shape"):
        vfm.register(s["scan_xyz"][:, :2], s["map_xyz"], s["scan_feat"], s["map_feat"])
    with pytest.raises(ValueError, match="Invalid shape"):
        vfm.register(s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"][:, :32])
    with pytest.raises(vfm.VfmRegError):
        vfm.register(s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"], ransac_iters=0)


def test_compat_shim_reference_layout(vfm):
    """The reference's own call surface (one (N, 3+D) array per cloud) through the shim."""
    from vfm_registration_b200 import compat
    s = synth.make_pair(4, 3000, 1000, 64)
    vmap_arr = np.c_[s["map_xyz"], s["map_feat"]].astype(np.float32)
    scan_arr = np.c_[s["scan_xyz"], s["scan_feat"]].astype(np.float32)
    vm = compat.VoxelHashMap(1.0, 100.0, 20)
    vm.add_points(vmap_arr)
    src, tgt = vm.get_vfm_correspondences(scan_arr, 0.8)
    osrc, otgt = match.get_vfm_correspondences(scan_arr, vmap_arr, 0.8)
    assert src.dtype == np.float64 and np.array_equal(src, osrc) and np.array_equal(tgt, otgt)
    node = compat.RegistrationNode(ransac_iters=2048)
    pose, icp = node.ransac_registration(vmap_arr, scan_arr, "vfm")
    assert icp is None and pose.shape == (4, 4)
    # the reference's literal tau = 10000 selects the min-residual hypothesis; it still lands near the truth here
    t_err, r_err = node.compute_errors(pose, s["T_gt"], "vfm")
    assert node.compute_success_rate("vfm", 2, 5) in (0.0, 1.0)
    with pytest.raises(ValueError, match="Invalid method"):
        node.ransac_registration(vmap_arr, scan_arr, "bogus")
    with pytest.raises(ValueError, match="Invalid shape"):
        vm.get_vfm_correspondences(scan_arr[:, :10], 0.8)


def test_register_batch_equals_sequential(vfm):
    """The double-buffered batch entry point returns exactly what per-pair calls return (different sizes per pair)."""
    pairs, want = [], []
    for i, (m, n) in enumerate(((3000, 1000), (2000, 1500), (4096, 700), (1000, 1000), (2500, 300))):
        s = synth.make_pair(50 + i, m, n, 128)
        pr = (s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"])
        pairs.append(pr)
        want.append(vfm.register(*pr, min_cos=0.8, mutual=True, ransac_iters=1024, inlier_thresh=1.0, seed=9))
    got = vfm.register_batch(pairs, min_cos=0.8, mutual=True, ransac_iters=1024, inlier_thresh=1.0, seed=9)
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert np.array_equal(g.T, w.T) and np.array_equal(g.corr, w.corr) and np.array_equal(g.inlier_mask, w.inlier_mask)
        assert g.best_hyp == w.best_hyp and g.fitness == w.fitness and g.rmse == w.rmse
    dev_pairs = [tuple(torch.from_numpy(x).cuda() for x in pr) for pr in pairs]
    gd = vfm.register_batch(dev_pairs, min_cos=0.8, mutual=True, ransac_iters=1024, inlier_thresh=1.0, seed=9)
    for g, w in zip(gd, want):
        assert np.array_equal(g.T, w.T) and np.array_equal(g.corr.cpu().numpy(), w.corr)
        assert np.array_equal(g.inlier_mask.cpu().numpy(), w.inlier_mask) and g.best_hyp == w.best_hyp
    # run it twice more: stage reuse across batches
    again = vfm.register_batch(pairs[::-1], min_cos=0.8, mutual=True, ransac_iters=1024, inlier_thresh=1.0, seed=9)
    for g, w in zip(again, want[::-1]):
        assert np.array_equal(g.T, w.T) and np.array_equal(g.corr, w.corr)


def test_config1_full_vs_oracle(vfm):
    """BASELINE configs[0]: 4096 x 4096 points, 384-d, 1024 RANSAC iterations, no mutual filter -- the reference's own
    CPU-runnable case, compared end to end with the oracle (C restatement bit-exact, NumPy float64 within 1e-4)."""
    s = synth.make_pair(1, 4096, 4096, 384)
    r = vfm.register(s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"], min_cos=0.8, ransac_iters=1024, inlier_thresh=1.0, seed=1)
    m = cref.match_nn(s["scan_feat"], s["map_feat"])
    o = match.match_nn(s["scan_feat"], s["map_feat"])
    clear = (o["sim01"] - o["sec01"]) > 1e-5
    assert clear.mean() > 0.99 and np.array_equal(m["idx01"][clear], o["idx01"][clear])
    corr = match.filter_correspondences(m["idx01"], m["sim01"], min_cos=0.8)
    assert np.array_equal(r.corr, corr)
    c = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, None, 1.0, seed=1, n_hyp=1024)
    assert r.best_hyp == c["best"] and np.array_equal(r.T, c["T"]) and np.array_equal(r.inlier_mask, c["mask"])
    on = ransac.ransac(s["scan_xyz"], s["map_xyz"], corr, ransac.sample_indices(1, 1024, len(corr)), 1.0)
    assert np.linalg.norm(r.T - on["T"]) < 1e-4 and np.array_equal(r.inlier_mask, on["mask"])
    rte, rre = synth.pose_errors(r.T, s["T_gt"])
    assert rte < 1.0 and rre < 5.0


def test_config4_shape_ratio_test_768d(vfm):
    """BASELINE configs[3] semantics (768-d, ratio test, many hypotheses) at a size the oracle finishes in seconds, plus
    the full 200k x 20k x 768 size through a size-independent property: tensor-core path == exact fp32 path, bit for bit."""
    s = synth.make_pair(4, 20000, 4000, 768, sigma_f=0.02)
    kw = dict(min_cos=None, ratio=0.9, ransac_iters=16384, inlier_thresh=1.0, seed=4)
    r = vfm.register(s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"], **kw)
    m = cref.match_nn(s["scan_feat"], s["map_feat"])
    corr = match.filter_correspondences(m["idx01"], m["sim01"], m["sec01"], ratio=0.9)
    assert np.array_equal(r.corr, corr) and 800 < len(corr) < 4000
    c = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, None, 1.0, seed=4, n_hyp=16384)
    assert r.best_hyp == c["best"] and np.array_equal(r.T, c["T"]) and np.array_equal(r.inlier_mask, c["mask"])
    rte, rre = synth.pose_errors(r.T, s["T_gt"])
    assert rte < 1.0 and rre < 5.0
    g = torch.Generator(device="cuda").manual_seed(0)
    a = torch.randn(20000, 768, device="cuda", generator=g)
    b = torch.randn(200000, 768, device="cuda", generator=g)
    t = vfm.match_nn(a, b, algo="tc")
    e = vfm.match_nn(a, b, algo="simt")
    assert torch.equal(t.idx01, e.idx01) and torch.equal(t.sim01, e.sim01) and torch.equal(t.sec01, e.sec01)


def test_full_size_ransac_properties(vfm):
    """configs[1]/[3] RANSAC sizes through size-independent properties: planted SE(3) is recovered, the winner's count equals
    the mask population, re-running is deterministic, and doubling the hypothesis budget never lowers the best count."""
    s = synth.make_pair(6, 200_000, 20_000, 16)
    inl = np.nonzero(s["perm"] >= 0)[0]
    rng = np.random.default_rng(0)
    k = 20_000
    j = np.where(s["perm"] >= 0, s["perm"], rng.integers(0, 200_000, k))
    corr = np.stack([np.arange(k), j], 1).astype(np.int32)
    r1 = vfm.ransac_kabsch(s["scan_xyz"], s["map_xyz"], corr, n_hyp=32768, seed=7, thresh=1.0)
    r2 = vfm.ransac_kabsch(s["scan_xyz"], s["map_xyz"], corr, n_hyp=65536, seed=7, thresh=1.0)
    r3 = vfm.ransac_kabsch(s["scan_xyz"], s["map_xyz"], corr, n_hyp=65536, seed=7, thresh=1.0)
    assert r2.n_inliers >= r1.n_inliers and int(r2.mask.sum()) == r2.n_inliers
    assert r2.best == r3.best and np.array_equal(r2.T, r3.T) and torch.equal(r2.counts, r3.counts)
    assert torch.equal(r2.counts[:32768], r1.counts)
    assert r2.n_inliers >= 0.95 * len(inl)
    rte, rre = synth.pose_errors(r2.T, s["T_gt"])
    assert rte < 0.2 and rre < 0.5


@pytest.mark.parametrize("m,n,d,min_cos", [(50_000, 10_000, 384, 0.8), (50_000, 10_000, 384, None), (3000, 3000, 384, 0.8),
                                           (5000, 200, 64, 0.5), (900, 100, 768, None), (4000, 1000, 128, 0.999),
                                           (300, 1, 32, None)])
def test_pruned_mutual_equals_full_search(vfm, m, n, d, min_cos):
    """register() answers the mutual check from a reverse search restricted to the map rows that gated queries point at
    (row count on the device); the correspondence list must equal gate + full reverse search, at BASELINE configs[1]
    size and at the edges (no gate, no survivor, fewer rows than one tile, a single query)."""
    s = synth.make_pair(70 + n % 50, m, n, d)
    dev = [torch.from_numpy(s[k]).cuda() for k in ("scan_xyz", "map_xyz", "scan_feat", "map_feat")]
    if n > 50:
        dev[2][5] = 0                 # zero-norm query
        dev[3][11] = dev[3][3]        # duplicated map row: tie -> lowest index
    r = vfm.register(*dev, min_cos=min_cos, mutual=True, ransac_iters=256, inlier_thresh=1.0, seed=3)
    full = vfm.match_nn(dev[2], dev[3], mutual=True)
    want = vfm.filter_correspondences(full, min_cos=min_cos, mutual=True)
    got = r.corr if isinstance(r.corr, np.ndarray) else r.corr.cpu().numpy()
    want = want if isinstance(want, np.ndarray) else want.cpu().numpy()
    assert np.array_equal(got, want)
    if min_cos == 0.8 and n >= 3000:
        assert len(got) > 0.2 * n


# ---- BASELINE-size parity pinned directly to the oracle (not to the repo's own exact kernel) ---------------------------
def test_configs1_full_size_vs_c_oracle(vfm):
    """BASELINE configs[1] (10k scan x 50k map x 384, mutual + cos >= 0.8, 8192 hypotheses, tau = 1 m) against oracle/c on all
    six match arrays, then the correspondence list, winning hypothesis, inlier mask and transform of register()."""
    s = synth.make_pair(2, 50_000, 10_000, 384)
    c = cref.match_nn(s["scan_feat"], s["map_feat"], mutual=True)
    g = vfm.match_nn(torch.from_numpy(s["scan_feat"]).cuda(), torch.from_numpy(s["map_feat"]).cuda(), mutual=True)
    for k in ("idx01", "sim01", "sec01", "idx10", "sim10", "sec10"):
        assert np.array_equal(getattr(g, k).cpu().numpy(), c[k]), k
    corr = match.filter_correspondences(c["idx01"], c["sim01"], c["sec01"], c["idx10"], min_cos=0.8, mutual=True)
    o = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, None, 1.0, seed=42, n_hyp=8192)
    for host in (True, False):
        args = (s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"])
        if not host:
            args = tuple(torch.from_numpy(x).cuda() for x in args)
        r = vfm.register(*args, min_cos=0.8, mutual=True, ransac_iters=8192, inlier_thresh=1.0, seed=42)
        assert np.array_equal(r.corr, corr) and len(corr) > 2000
        assert r.best_hyp == o["best"] and np.array_equal(r.inlier_mask, o["mask"]) and np.array_equal(r.T, o["T"])
    # float64 NumPy oracle of the solve (independent arithmetic): same winner and mask, T within 1e-4 Frobenius
    on = ransac.ransac(s["scan_xyz"], s["map_xyz"], corr, cref.sample_indices(42, 8192, len(corr)), 1.0)
    assert on["best"] == r.best_hyp and np.array_equal(on["mask"], r.inlier_mask) and np.linalg.norm(on["T"] - r.T) < 1e-4


def test_configs3_full_size_vs_c_oracle(vfm):
    """BASELINE configs[3] (20k scan x 200k map x 768, ratio test 0.9, 65536 hypotheses) against oracle/c: top-2 arrays of
    the forward search, then correspondences / winner / mask / transform of register()."""
    s = synth.make_pair(4, 200_000, 20_000, 768, sigma_f=0.02)
    c = cref.match_nn(s["scan_feat"], s["map_feat"])
    a, b = torch.from_numpy(s["scan_feat"]).cuda(), torch.from_numpy(s["map_feat"]).cuda()
    g = vfm.match_nn(a, b)
    for k in ("idx01", "sim01", "sec01"):
        assert np.array_equal(getattr(g, k).cpu().numpy(), c[k]), k
    corr = match.filter_correspondences(c["idx01"], c["sim01"], c["sec01"], ratio=0.9)
    o = cref.ransac(s["scan_xyz"], s["map_xyz"], corr, None, 1.0, seed=4, n_hyp=65536)
    r = vfm.register(torch.from_numpy(s["scan_xyz"]).cuda(), torch.from_numpy(s["map_xyz"]).cuda(), a, b, min_cos=None, ratio=0.9,
                     ransac_iters=65536, inlier_thresh=1.0, seed=4)
    assert np.array_equal(r.corr, corr) and len(corr) > 3000
    assert r.best_hyp == o["best"] and np.array_equal(r.inlier_mask, o["mask"]) and np.array_equal(r.T, o["T"])
    rte, rre = synth.pose_errors(r.T, s["T_gt"])
    assert rte < 1.0 and rre < 5.0


def test_reference_shape_vs_c_oracle(vfm):
    """The reference's own problem shape (~300 voxelised queries against a 200k-point map, 384-d, cosine gate)."""
    rng = np.random.default_rng(77)
    b = rng.standard_normal((200_000, 384)).astype(np.float32)
    a = rng.standard_normal((300, 384)).astype(np.float32)
    a[:120] = b[rng.integers(0, 200_000, 120)] + 0.02 * rng.standard_normal((120, 384)).astype(np.float32)
    c = cref.match_nn(a, b)
    g = vfm.match_nn(a, b)
    for k in ("idx01", "sim01", "sec01"):
        assert np.array_equal(getattr(g, k).cpu().numpy(), c[k]), k


# ---- maps shared by the scans of a scene -----------------------------------------------------------------------------------
def _scene(seed, m, ns, d):
    """One map and len(ns) scans of it (each with its own planted pose)."""
    base = synth.make_pair(seed, m, ns[0], d, scan_seed=seed * 100)
    scans = [(base["scan_xyz"], base["scan_feat"], base["T_gt"])]
    for k, n in enumerate(ns[1:]):
        s = synth.make_pair(seed, m, n, d, scan_seed=seed * 100 + k + 1)
        assert np.array_equal(s["map_feat"], base["map_feat"])
        scans.append((s["scan_xyz"], s["scan_feat"], s["T_gt"]))
    return base["map_xyz"], base["map_feat"], scans


@pytest.mark.parametrize("kw", [dict(min_cos=0.8, mutual=True, inlier_thresh=1.0), dict(min_cos=0.8, inlier_thresh=1.0),
                                dict(min_cos=None, ratio=0.9, inlier_thresh=1.0), dict(min_cos=0.5, mutual=True)])
def test_shared_map_batches_equal_sequential(vfm, kw):
    """register_batch with pairs that share their target objects (host and device), ResidentMap + register_scans (host and
    device): every route returns exactly what per-pair register() calls return, and those equal the oracle chain."""
    scenes = [_scene(300, 6000, (1500, 900, 2000), 128), _scene(301, 4000, (1000, 1000), 128), _scene(302, 5000, (700,), 128)]
    pairs, want = [], []
    for mx, mf, scans in scenes:
        for sx, sf, _ in scans:
            pairs.append((sx, mx, sf, mf))
            want.append(vfm.register(sx, mx, sf, mf, ransac_iters=1024, seed=11, **kw))
    # the sequential results against the oracle chain (first scene)
    mx, mf, scans = scenes[0]
    for (sx, sf, _), w in zip(scans, want):
        m = cref.match_nn(sf, mf, mutual=True)
        corr = match.filter_correspondences(m["idx01"], m["sim01"], m["sec01"], m["idx10"], min_cos=kw.get("min_cos"),
                                            mutual=kw.get("mutual", False), ratio=kw.get("ratio"))
        o = cref.ransac(sx, mx, corr, None, kw.get("inlier_thresh", 1e4), seed=11, n_hyp=1024)
        assert np.array_equal(w.corr, corr) and w.best_hyp == o["best"] and np.array_equal(w.T, o["T"])
        assert np.array_equal(w.inlier_mask, o["mask"])

    def same(got):
        assert len(got) == len(want)
        for g, w in zip(got, want):
            gc = g.corr if isinstance(g.corr, np.ndarray) else g.corr.cpu().numpy()
            gm = g.inlier_mask if isinstance(g.inlier_mask, np.ndarray) else g.inlier_mask.cpu().numpy()
            assert np.array_equal(g.T, w.T) and np.array_equal(gc, w.corr) and np.array_equal(gm, w.inlier_mask)
            assert g.best_hyp == w.best_hyp and g.fitness == w.fitness and g.rmse == w.rmse

    same(vfm.register_batch(pairs, ransac_iters=1024, seed=11, **kw))
    cache = {}
    dev = lambda x: cache.setdefault(id(x), torch.from_numpy(x).cuda())  # noqa: E731  shared objects stay shared
    same(vfm.register_batch([tuple(dev(x) for x in pr) for pr in pairs], ransac_iters=1024, seed=11, **kw))
    vfm.get_context(0).set_lanes(2)      # more runs than ring slots: slot reuse across lanes
    same(vfm.register_batch([tuple(dev(x) for x in pr) for pr in pairs * 3], ransac_iters=1024, seed=11, **kw)[len(pairs):2 * len(pairs)])
    vfm.get_context(0).set_lanes(5)
    i = 0
    for mx, mf, scans in scenes:
        for host in (True, False):
            rm = vfm.ResidentMap(mx, mf) if host else vfm.ResidentMap(dev(mx), dev(mf))
            sc = [(sx, sf) if host else (dev(sx), dev(sf)) for sx, sf, _ in scans]
            got = vfm.register_scans(rm, sc, ransac_iters=1024, seed=11, **kw)
            for g, w in zip(got, want[i:i + len(scans)]):
                gc = g.corr if isinstance(g.corr, np.ndarray) else g.corr.cpu().numpy()
                assert np.array_equal(g.T, w.T) and np.array_equal(gc, w.corr) and g.best_hyp == w.best_hyp
            rm.close()
        i += len(scans)


def test_resident_map_match_and_gate_floor(vfm):
    """ResidentMap.match: with the runner-up it equals match_nn; with a cosine gate handed to the search, every query whose
    oracle best reaches the gate gets exactly the oracle's (index, score), every other query reports something the gate
    drops (index -1 / -inf, or a score below the gate)."""
    s = synth.make_pair(88, 30_000, 4000, 384)
    a = s["scan_feat"].copy()
    a[7] = 0
    # queries close to the gate: blend a map row with noise so that the cosine lands around 0.8
    rng = np.random.default_rng(3)
    for i, t in zip(range(100, 160), np.linspace(0.65, 0.85, 60)):
        v = rng.standard_normal(384).astype(np.float32)
        v /= np.linalg.norm(v)
        a[i] = s["map_feat"][i] + t * v
    c = cref.match_nn(a, s["map_feat"])
    rm = vfm.ResidentMap(s["map_xyz"], s["map_feat"])
    g = rm.match(a)
    for k in ("idx01", "sim01", "sec01"):
        assert np.array_equal(getattr(g, k).cpu().numpy(), c[k]), k
    for gate in (0.8, 0.3, 0.05, 0.99):
        f = rm.match(a, min_cos=gate, second=False)
        idx, sim = f.idx01.cpu().numpy(), f.sim01.cpu().numpy()
        above = c["sim01"] >= gate
        assert np.array_equal(idx[above], c["idx01"][above]) and np.array_equal(sim[above], c["sim01"][above])
        assert np.all(sim[~above] < gate)
        corr = vfm.filter_correspondences(f, min_cos=gate).cpu().numpy()
        assert np.array_equal(corr, match.filter_correspondences(c["idx01"], c["sim01"], min_cos=gate))
    assert 20 < int((c["sim01"][100:160] >= 0.8).sum()) < 50   # the gate really cuts through the blended queries
    with pytest.raises(ValueError, match="Invalid shape"):
        rm.match(a[:, :100])
    with pytest.raises(ValueError, match="Invalid shape"):
        vfm.ResidentMap(s["map_xyz"][:10], s["map_feat"])


# ---- a8 / a7 remainder: L2 distances and "keep the n_points smallest" -------------------------------------------------------
def test_l2_distances_and_select_smallest(vfm):
    """registration_node.py:191-214 (brute-force L2 block + n_points selection) and :510-518 (non-mutual branch of
    find_correspondences) against the oracle restatements; the latter is pinned to the reference by find_corr.npz."""
    import os
    from vfm_registration_b200 import compat
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "find_corr.npz"))
    f0 = (g["f0"] / np.linalg.norm(g["f0"], axis=1, keepdims=True)).astype(np.float32)
    f1 = (g["f1"] / np.linalg.norm(g["f1"], axis=1, keepdims=True)).astype(np.float32)
    # the reference's own function on these unit vectors (oracle restatement, pinned by the golden fixture on raw features)
    j0, j1 = match.find_correspondences(f0, f1, n_points=100, mutual_filter=False)
    o = np.argsort(j0)
    k0, k1 = compat.find_correspondences(f0, f1, n_points=100, mutual_filter=False)
    assert np.array_equal(k0, j0[o]) and np.array_equal(k1, j1[o])
    i0, i1 = match.find_correspondences(f0, f1, mutual_filter=True)
    m0, m1 = compat.find_correspondences(f0, f1, mutual_filter=True)
    assert np.array_equal(m0, i0) and np.array_equal(m1, i1)
    # distances of the brute-force block
    idx, dis = match.l2_block_argmin(f0, f1)
    mt = vfm.match_nn(f0, f1)
    d = vfm.l2_distances(mt).cpu().numpy()
    assert np.array_equal(mt.idx01.cpu().numpy(), idx) and np.abs(d ** 2 - dis ** 2).max() < 2e-6   # 2 - 2 s: fp32 summation order of s
    # larger case with ties at the boundary and unmatched queries: the n smallest (distance, index) pairs, in query order
    rng = np.random.default_rng(5)
    n = 40_000
    sim = torch.from_numpy(np.round(rng.uniform(-0.2, 1.0, n), 3).astype(np.float32)).cuda()   # 1201 distinct values -> many ties
    idx01 = torch.from_numpy(rng.integers(0, 1000, n).astype(np.int32)).cuda()
    idx01[::97] = -1
    mt2 = vfm.api.MatchResult(idx01, sim, None)
    simh, idxh = sim.cpu().numpy(), idx01.cpu().numpy()
    valid = np.nonzero(idxh >= 0)[0]
    for keep in (0, 1, 5000, len(valid) - 1, len(valid), n + 5):
        corr, dist = vfm.select_smallest(mt2, keep, return_distance=True)
        corr, dist = corr.cpu().numpy(), dist.cpu().numpy()
        order = valid[np.lexsort((valid, -simh[valid]))][:min(keep, len(valid))]   # largest similarity first, then lowest index
        want = np.sort(order)
        assert np.array_equal(corr[:, 0], want) and np.array_equal(corr[:, 1], idxh[want])
        assert np.allclose(dist, np.sqrt(2 - 2 * simh[want].astype(np.float64) + 1e-6), atol=2e-6)


# ---- the Open3D-style hypothesis score (nearest neighbour of every source point) ---------------------------
def test_kdtree_nearest_exact(vfm):
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(2)
    for n, nq in ((1, 10), (7, 50), (5000, 4000), (200_000, 20_000)):
        pts = np.c_[rng.uniform(-50, 50, (n, 2)), rng.uniform(-2, 8, n)]
        if n > 100:
            pts[5] = pts[3]                                   # duplicate point
        q = np.r_[np.c_[rng.uniform(-60, 60, (nq - 3, 2)), rng.uniform(-30, 40, nq - 3)], [[900.0, -700.0, 300.0]], pts[:2] if n > 1 else pts[[0, 0]]]
        tree = vfm.KdTree(pts)
        idx, d2 = tree.nearest(q)
        d, j = cKDTree(pts).query(q)
        assert np.array_equal(np.sum((pts[idx.cpu().numpy()] - q) ** 2, 1), d2.cpu().numpy())   # the reported point, exactly
        assert np.allclose(np.sqrt(d2.cpu().numpy()), d, rtol=1e-12, atol=1e-12)                   # ... is the nearest one
        idx2, _ = tree.nearest(q, max_dist=1.0)
        assert np.array_equal(idx2.cpu().numpy() >= 0, d < 1.0)


@pytest.mark.parametrize("n_map,n_scan,n_hyp,max_dist", [(3000, 700, 512, 1e4), (20_000, 3000, 2048, 1e4), (20_000, 3000, 1024, 0.5),
                                                         (60_000, 9000, 512, 1e4)])
def test_ransac_nn_all_vs_oracle(vfm, n_map, n_scan, n_hyp, max_dist):
    """Winner, per-hypothesis inlier counts and distance sums against the cKDTree restatement (oracle/ransac.py), with the
    hypotheses of the C restatement so that both sides score bit-identical transforms."""
    # a scan whose every point lies on the map (the chamfer criterion is about geometry: uniform clutter thrown outside the
    # map's bounding box by the true pose would dominate it); two thirds of the correspondences are wrong
    s = synth.make_pair(41, n_map, n_scan, 16, inlier_frac=1.0)
    rng = np.random.default_rng(3)
    k = min(600, n_scan // 2)
    q = np.sort(rng.permutation(n_scan)[:k])
    j = s["perm"][q].copy()
    wrong = rng.permutation(k)[: 2 * k // 3]
    j[wrong] = rng.integers(0, n_map, len(wrong))
    corr = np.stack([q, j], 1).astype(np.int32)
    si = ransac.sample_indices(5, n_hyp, len(corr))
    si[3] = [2, 2, 2]                                        # degenerate sample
    fit = lambda p, q: cref.kabsch3(p, q)                      # noqa: E731
    o = ransac.ransac_nn_all(s["scan_xyz"], s["map_xyz"], corr, si, max_dist, fit=fit)
    g = vfm.ransac_nn_all(s["scan_xyz"], s["map_xyz"], corr, sample_idx=si, max_dist=max_dist)
    assert np.array_equal(g.inliers.cpu().numpy(), o["inliers"]) and g.inliers[3].item() == -1
    assert np.allclose(g.sum_d2.cpu().numpy(), o["sum_d2"], rtol=1e-11, atol=1e-9)
    assert g.best == o["best"] and np.array_equal(g.T, o["T"])
    assert abs(g.fitness - o["fitness"]) < 1e-12 and abs(g.rmse - o["rmse"]) < 1e-9
    if max_dist >= 1e4:
        assert g.fitness == 1.0
    rte, rre = synth.pose_errors(g.T, s["T_gt"])
    assert rte < 1.0 and rre < 5.0
    # device-side sampler, and K < 3 -> identity
    g2 = vfm.ransac_nn_all(s["scan_xyz"], s["map_xyz"], corr, n_hyp=256, seed=9, max_dist=max_dist)
    o2 = ransac.ransac_nn_all(s["scan_xyz"], s["map_xyz"], corr, cref.sample_indices(9, 256, len(corr)), max_dist, fit=fit)
    assert g2.best == o2["best"] and np.array_equal(g2.T, o2["T"])
    g3 = vfm.ransac_nn_all(s["scan_xyz"], s["map_xyz"], corr[:2], n_hyp=16, max_dist=max_dist)
    assert g3.best == -1 and np.array_equal(g3.T, np.eye(4)) and g3.fitness == 0.0


def test_compat_node_default_score_is_open3d_style(vfm):
    """At the reference's literal max_correspondence_distance = 10000 the shim scores hypotheses as Open3D does (nearest
    neighbour of every scan point); checked against the oracle chain, with and without the reference's pre-processing."""
    from vfm_registration_b200 import compat
    s = synth.make_pair(9, 5000, 1500, 64, inlier_frac=0.5)
    vmap_arr = np.c_[s["map_xyz"], s["map_feat"]].astype(np.float32)
    scan_arr = np.c_[s["scan_xyz"], s["scan_feat"]].astype(np.float32)
    node = compat.RegistrationNode(ransac_iters=512, preprocess=False)
    assert node.score == "nn_all" and compat.RegistrationNode(max_correspondence_distance=1.0).score == "corr"
    pose, _ = node.ransac_registration(vmap_arr, scan_arr, "vfm")
    m = cref.match_nn(s["scan_feat"], s["map_feat"])
    corr = match.filter_correspondences(m["idx01"], m["sim01"], min_cos=0.8)
    o = ransac.ransac_nn_all(s["scan_xyz"], s["map_xyz"], corr, cref.sample_indices(42, 512, len(corr)), 1e4, fit=lambda p, q: cref.kabsch3(p, q))
    assert np.array_equal(pose, o["T"])
    rte, rre = synth.pose_errors(pose, s["T_gt"])
    assert rte < 0.5 and rre < 2.0
    node2 = compat.RegistrationNode(ransac_iters=512)          # with the voxel pre-processing of registration_node.py:287-291, 396-425
    pose2, icp = node2.ransac_registration(vmap_arr, scan_arr, "vfm", run_icp=True)
    rte2, rre2 = synth.pose_errors(icp, s["T_gt"])
    assert pose2.shape == (4, 4) and rte2 < 0.2 and rre2 < 0.5


# ---- SURVEY 8f row 4, second half: the TEASER++-style solve (csrc/teaser.cu) against oracle/teaser.py ---------------------
def _teaser_problem(seed, n_in, n_out, noise=0.02):
    rng = np.random.default_rng(seed)
    from scipy.spatial.transform import Rotation as R
    rot = R.from_euler("zyx", rng.uniform(-180, 180, 3) * [1, 0.05, 0.05], degrees=True).as_matrix()
    t = rng.normal(0, 10, 3)
    src = rng.uniform(-40, 40, (n_in + n_out, 3))
    tgt = src @ rot.T + t + rng.normal(0, noise, (n_in + n_out, 3))
    tgt[n_in:] = rng.uniform(-40, 40, (n_out, 3))
    perm = rng.permutation(n_in + n_out)
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = rot, t
    return src[perm], tgt[perm], T


@pytest.mark.parametrize("seed,n_in,n_out", [(1, 40, 110), (2, 25, 175), (3, 120, 60)])
def test_teaser_solve_vs_oracle(vfm, seed, n_in, n_out):
    """Compatibility graph (GPU), exact maximum clique, GNC-TLS rotation and TLS translation against the NumPy restatement
    (an unrelated clique algorithm and np.linalg.svd): same clique, pose within 1e-9 (float64, different summation order)."""
    from oracle import teaser as ot
    src, tgt, T_gt = _teaser_problem(seed, n_in, n_out)
    r = vfm.teaser_solve(src, tgt)
    T_o, cl_o = ot.teaser_solve(src, tgt)
    assert r.exact and len(r.clique) == len(cl_o)
    if not np.array_equal(r.clique, cl_o):            # maximum cliques of equal size: solve the oracle on the product's clique
        T_o, _ = ot.teaser_solve(src, tgt, clique=r.clique)
    assert np.abs(r.T - T_o).max() < 1e-9, float(np.abs(r.T - T_o).max())
    assert np.linalg.norm(r.T[:3, 3] - T_gt[:3, 3]) < 0.05
    assert np.degrees(np.arccos(np.clip((np.trace(r.T[:3, :3].T @ T_gt[:3, :3]) - 1) / 2, -1, 1))) < 0.2
    # the graph itself, through the clique: every pair of clique members is compatible in the oracle's graph
    g = ot.tim_graph(src, tgt, 0.2)
    assert g[np.ix_(r.clique, r.clique)].sum() == len(r.clique) * (len(r.clique) - 1)


def test_teaser_solve_edge_cases_and_size(vfm):
    r = vfm.teaser_solve(np.zeros((0, 3)), np.zeros((0, 3)))
    assert np.array_equal(r.T, np.eye(4)) and len(r.clique) == 0
    r = vfm.teaser_solve(np.ones((1, 3)), np.ones((1, 3)))
    assert np.array_equal(r.T, np.eye(4))
    with pytest.raises(ValueError, match="Invalid shape"):
        vfm.teaser_solve(np.zeros((4, 3)), np.zeros((5, 3)))
    # the size the reference feeds it (a few thousand correspondences of the 1 m-voxelised scan), 60 % outliers
    src, tgt, T_gt = _teaser_problem(7, 1200, 1800)
    r = vfm.teaser_solve(src, tgt)
    assert r.exact and 1200 <= len(r.clique) <= 1210
    assert np.linalg.norm(r.T[:3, 3] - T_gt[:3, 3]) < 0.02


def test_compat_teaser_registration(vfm):
    """compat.RegistrationNode.teaser_registration (registration_node.py:91-160) on a synthetic descriptor-carrying pair."""
    from vfm_registration_b200 import compat, synth
    s = synth.make_pair(5, 6000, 3000, 64, inlier_frac=0.5)
    node = compat.RegistrationNode(preprocess=False)
    vmap = np.c_[s["map_xyz"], s["map_feat"]]
    scan = np.c_[s["scan_xyz"], s["scan_feat"]]
    pose, icp = node.teaser_registration(vmap, scan, "vfm")
    assert icp is None
    rte, rre = synth.pose_errors(pose, s["T_gt"])
    assert rte < 0.2 and rre < 0.5, (rte, rre)
    with pytest.raises(ValueError, match="Invalid method"):
        node.teaser_registration(vmap, scan, "fpfh")
