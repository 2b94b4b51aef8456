"""The ViT / preprocessing oracle against (i) the reference's own transform (golden) and (ii) transformers.Dinov2Model
with the same seeded random weights (architecture-level parity; hub weights are unreachable offline)."""
import numpy as np
import pytest
import torch

from oracle import vit


def test_preprocess_matches_reference_transform(golden):
    g = golden("preprocess.npz")
    x = vit.preprocess(g["image"])
    assert tuple(x.shape) == tuple(g["shape"]) and vit.patch_grid(*g["image"].shape[:2])[1] == int(g["patch_w"])
    assert np.abs(x[:, ::7, ::9].numpy() - g["sub"]).max() < 1e-6
    assert abs(float(x.double().sum()) - float(g["total"])) < 1e-2
    assert abs(float(x.double().abs().sum()) - float(g["abs_total"])) < 1e-2


@pytest.mark.parametrize("hw", [(224, 224), (224, 252)])
def test_forward_matches_transformers_dinov2(hw):
    transformers = pytest.importorskip("transformers")
    cfg = vit.ViTConfig(depth=3, width=384, heads=6)
    sd = vit.make_weights(cfg, seed=3)
    hf_cfg = transformers.Dinov2Config(hidden_size=cfg.width, num_hidden_layers=cfg.depth, num_attention_heads=cfg.heads,
                                       mlp_ratio=cfg.mlp_ratio, patch_size=cfg.patch, image_size=518,
                                       layer_norm_eps=cfg.ln_eps, attn_implementation="eager")
    model = transformers.Dinov2Model(hf_cfg).eval()
    missing = model.load_state_dict(vit.to_hf_state_dict(sd, cfg), strict=True)
    x = torch.randn(2, 3, *hw, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        want = model(pixel_values=x).last_hidden_state          # (B, 1 + Np, W), after the final LayerNorm
    _, hidden = vit.forward(sd, cfg, x, return_hidden=True)
    assert hidden.shape == want.shape
    assert (hidden - want).abs().max() < 2e-4
    feats = vit.forward(sd, cfg, x)
    assert feats.shape == (2, hw[0] // 14, hw[1] // 14, cfg.width)
    # ChannelNorm = LayerNorm over C of the patch tokens
    ref = torch.nn.functional.layer_norm(want[:, 1:], (cfg.width,), sd["channel_norm.weight"], sd["channel_norm.bias"], cfg.cn_eps)
    assert (feats.reshape(2, -1, cfg.width) - ref).abs().max() < 5e-4
