"""One process, two devices: kernel attributes (dynamic shared memory opt-in, resident cluster counts, staging rings) are kept per
context = per device, so a second GPU used after the first one launches the same kernels (round-1 advisor finding: process-wide
`static bool attr_set` flags).  Skipped on single-GPU boxes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_second_device_after_first():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import vfm_registration_b200 as v
    from vfm_registration_b200 import synth
    s = synth.make_pair(3, 6000, 3000, 384)
    out = []
    for dev in (0, 1):
        r = v.register(s["scan_xyz"], s["map_xyz"], s["scan_feat"], s["map_feat"], min_cos=0.8, mutual=True, ransac_iters=1024,
                       inlier_thresh=1.0, seed=5, device=dev)
        with pytest.warns(RuntimeWarning):
            f = v.ViTFeaturizer("vits14", seed=3, random_init=True, device=dev)
        imgs = np.random.default_rng(1).integers(0, 255, (2, 224, 224, 3), dtype=np.uint8)
        tok = f.forward(imgs).cpu().numpy()
        out.append((r.T, np.asarray(r.corr.cpu() if hasattr(r.corr, "cpu") else r.corr), tok))
        f.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert np.array_equal(out[0][2], out[1][2])     # the forward is deterministic: same weights, same images, same bits
