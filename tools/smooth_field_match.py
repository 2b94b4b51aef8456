"""How the candidate search behaves on descriptors sampled from a low-resolution token grid (what the reference's pipeline
produces: bilinear samples of a 16 x Wp DINOv2 grid) instead of independent random rows: many near-duplicate map rows."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vfm_registration_b200 as v

rng = np.random.default_rng(0)
d, cams, g = 384, 6, 16
tok = rng.standard_normal((cams, g, g, d)).astype(np.float32)
tok /= np.linalg.norm(tok, axis=-1, keepdims=True)

def sample(c, y, x):
    y0, x0 = np.floor(y).astype(int), np.floor(x).astype(int)
    y1, x1 = np.minimum(y0 + 1, g - 1), np.minimum(x0 + 1, g - 1)
    wy, wx = (y - y0)[:, None], (x - x0)[:, None]
    return ((1 - wy) * ((1 - wx) * tok[c, y0, x0] + wx * tok[c, y0, x1]) + wy * ((1 - wx) * tok[c, y1, x0] + wx * tok[c, y1, x1])).astype(np.float32)

m, n = 50_000, 10_000
cm, ym, xm = rng.integers(0, cams, m), rng.uniform(0, g - 1, m), rng.uniform(0, g - 1, m)
map_feat = sample(cm, ym, xm)
pick = rng.choice(m, n, replace=False)
for jitter in (0.0, 0.01, 0.05):
    scan_feat = sample(cm[pick], np.clip(ym[pick] + rng.normal(0, jitter, n), 0, g - 1), np.clip(xm[pick] + rng.normal(0, jitter, n), 0, g - 1))
    rm = v.ResidentMap(torch.zeros(m, 3, device="cuda"), torch.from_numpy(map_feat).cuda())
    a = torch.from_numpy(scan_feat).cuda()
    ctx = v.get_context(0)
    for name, kw in (("top-1 + gate 0.8", dict(min_cos=0.8, second=False)), ("top-2", dict(second=True))):
        for _ in range(3):
            r = rm.match(a, **kw)
        torch.cuda.synchronize()
        ctx.enable_timing(True)
        for _ in range(5):
            r = rm.match(a, **kw)
        torch.cuda.synchronize()
        gm, gl = ctx.group_time_ms(0)
        rr, rl = ctx.group_time_ms(6)
        ctx.enable_timing(False)
        hit = float((r.idx01.cpu().numpy() == pick).mean())
        print(f"jitter {jitter:.2f} cells, {name:18s}: search kernel {gm / max(gl, 1) * 1e3:7.1f} us, re-rank {rr / max(rl, 1) * 1e3:6.1f} us, "
              f"query finds its own map point {100 * hit:5.1f} %")
    rm.close()
