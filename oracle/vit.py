"""Oracle (test infrastructure): the image-feature producer, torch-CPU float32.

Follows ``src/vfm-reg/src/vfm_reg/image_features.py:67-77`` (ToTensor -> bilinear Resize to (14*16, 14*patch_w),
antialias=False -> ImageNet Normalize; PINNED by tests/golden/preprocess.npz) and ``:95-101`` (``self.model.model(x)`` =
FeatUp's DINOv2 featurizer + ChannelNorm).  The network itself is third party (torch.hub "mhamilton723/FeatUp" ->
facebookresearch/dinov2, unpinned, needs network): **parity unpinned** w.r.t. the hub weights.  The published
architecture is restated: Conv2d(3, W, 14, 14) patch embedding, CLS token, position embedding
bicubically interpolated to the patch grid, depth x [LN(1e-6) -> MHA (qkv bias) -> LayerScale -> +; LN -> MLP(GELU)
-> LayerScale -> +], final LN, patch tokens only, then LayerNorm over channels (FeatUp ChannelNorm).  It is checked
against ``transformers.Dinov2Model`` with the same seeded random weights (tests/test_oracle_vit.py).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


@dataclass
class ViTConfig:
    depth: int = 12
    width: int = 384
    heads: int = 6
    mlp_ratio: int = 4
    patch: int = 14
    pos_grid: int = 37          # 518 / 14, the grid the position embedding was trained on
    ln_eps: float = 1e-6
    cn_eps: float = 1e-4        # FeatUp ChannelNorm = LayerNorm(dim, eps=1e-4) over the channel axis

    @property
    def mlp_dim(self):
        return self.width * self.mlp_ratio


CONFIGS = {"vits14": ViTConfig(12, 384, 6), "vitb14": ViTConfig(12, 768, 12), "vitl14": ViTConfig(24, 1024, 16)}


def make_weights(cfg: ViTConfig, seed: int = 0) -> dict:
    """Seeded random weights in dinov2-hub naming (no checkpoint is reachable offline)."""
    g = torch.Generator().manual_seed(seed)
    w = cfg.width

    def rn(*shape, std=0.02):
        return torch.randn(*shape, generator=g) * std

    def un(*shape, lo, hi):
        return torch.rand(*shape, generator=g) * (hi - lo) + lo

    sd = {"patch_embed.proj.weight": rn(w, 3, cfg.patch, cfg.patch, std=0.05), "patch_embed.proj.bias": rn(w, std=0.02),
          "cls_token": rn(1, 1, w, std=0.5), "pos_embed": rn(1, 1 + cfg.pos_grid ** 2, w, std=0.2),
          "norm.weight": un(w, lo=0.8, hi=1.2), "norm.bias": rn(w, std=0.05),
          "channel_norm.weight": un(w, lo=0.8, hi=1.2), "channel_norm.bias": rn(w, std=0.05)}
    for l in range(cfg.depth):
        p = f"blocks.{l}."
        sd[p + "norm1.weight"] = un(w, lo=0.8, hi=1.2)
        sd[p + "norm1.bias"] = rn(w, std=0.05)
        sd[p + "attn.qkv.weight"] = rn(3 * w, w, std=1.5 / math.sqrt(w))
        sd[p + "attn.qkv.bias"] = rn(3 * w, std=0.05)
        sd[p + "attn.proj.weight"] = rn(w, w, std=1.0 / math.sqrt(w))
        sd[p + "attn.proj.bias"] = rn(w, std=0.02)
        sd[p + "ls1.gamma"] = un(w, lo=0.2, hi=1.0)
        sd[p + "norm2.weight"] = un(w, lo=0.8, hi=1.2)
        sd[p + "norm2.bias"] = rn(w, std=0.05)
        sd[p + "mlp.fc1.weight"] = rn(cfg.mlp_dim, w, std=1.0 / math.sqrt(w))
        sd[p + "mlp.fc1.bias"] = rn(cfg.mlp_dim, std=0.05)
        sd[p + "mlp.fc2.weight"] = rn(w, cfg.mlp_dim, std=1.0 / math.sqrt(cfg.mlp_dim))
        sd[p + "mlp.fc2.bias"] = rn(w, std=0.02)
        sd[p + "ls2.gamma"] = un(w, lo=0.2, hi=1.0)
    return sd


def patch_grid(img_h: int, img_w: int, patch: int = 14, patch_h: int = 16):
    """image_features.py:67-69: scale = 14*16 / H; patch_w = int(scale * W / 14)."""
    scale = (patch * patch_h) / img_h
    return patch_h, int(scale * img_w / patch)


def preprocess(image_u8: np.ndarray, patch: int = 14, patch_h: int = 16) -> torch.Tensor:
    """HWC uint8 -> (3, 14*16, 14*patch_w) float32: /255, bilinear resize (align_corners=False, no antialias),
    ImageNet normalisation (image_features.py:71-77)."""
    gh, gw = patch_grid(image_u8.shape[0], image_u8.shape[1], patch, patch_h)
    x = torch.from_numpy(np.ascontiguousarray(image_u8)).permute(2, 0, 1).to(torch.float32) / 255.0
    x = F.interpolate(x[None], size=(patch * gh, patch * gw), mode="bilinear", align_corners=False, antialias=False)[0]
    mean = torch.tensor(IMAGENET_MEAN)[:, None, None]
    std = torch.tensor(IMAGENET_STD)[:, None, None]
    return (x - mean) / std


def interp_pos_embed(pos_embed: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """(1, 1 + G*G, W) -> (1 + gh*gw, W): bicubic, align_corners=False, size= (the transformers / recent dinov2
    formulation; the older dinov2 `scale_factor + 0.1` quirk is not reproduced)."""
    w = pos_embed.shape[-1]
    g = int(round(math.sqrt(pos_embed.shape[1] - 1)))
    cls_pos, patch_pos = pos_embed[0, :1], pos_embed[0, 1:]
    if (gh, gw) != (g, g):
        patch_pos = patch_pos.reshape(1, g, g, w).permute(0, 3, 1, 2)
        patch_pos = F.interpolate(patch_pos.float(), size=(gh, gw), mode="bicubic", align_corners=False)
        patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(gh * gw, w)
    return torch.cat([cls_pos, patch_pos], dim=0)


@torch.no_grad()
def forward(sd: dict, cfg: ViTConfig, x: torch.Tensor, channel_norm: bool = True, return_hidden: bool = False):
    """x (B, 3, H, W) normalised float32 -> patch features (B, H/14, W/14, C)."""
    b, _, h, w_img = x.shape
    gh, gw = h // cfg.patch, w_img // cfg.patch
    wd, nh = cfg.width, cfg.heads
    dh = wd // nh
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=cfg.patch)  # (B, W, gh, gw)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd["cls_token"].expand(b, -1, -1), t], dim=1) + interp_pos_embed(sd["pos_embed"], gh, gw)[None]
    for l in range(cfg.depth):
        p = f"blocks.{l}."
        y = F.layer_norm(t, (wd,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.ln_eps)
        qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(b, -1, 3, nh, dh).permute(2, 0, 3, 1, 4)
        att = torch.softmax((qkv[0] @ qkv[1].transpose(-1, -2)) / math.sqrt(dh), dim=-1) @ qkv[2]
        att = att.transpose(1, 2).reshape(b, -1, wd)
        t = t + sd[p + "ls1.gamma"] * F.linear(att, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        y = F.layer_norm(t, (wd,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], cfg.ln_eps)
        y = F.linear(F.gelu(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])), sd[p + "mlp.fc2.weight"],
                     sd[p + "mlp.fc2.bias"])
        t = t + sd[p + "ls2.gamma"] * y
    hidden = F.layer_norm(t, (wd,), sd["norm.weight"], sd["norm.bias"], cfg.ln_eps)
    feats = hidden[:, 1:]
    if channel_norm:
        feats = F.layer_norm(feats, (wd,), sd["channel_norm.weight"], sd["channel_norm.bias"], cfg.cn_eps)
    feats = feats.reshape(b, gh, gw, wd)
    return (feats, hidden) if return_hidden else feats


def to_hf_state_dict(sd: dict, cfg: ViTConfig) -> dict:
    """dinov2-hub names -> transformers.Dinov2Model names (for the cross-check only)."""
    w = cfg.width
    out = {"embeddings.cls_token": sd["cls_token"], "embeddings.mask_token": torch.zeros(1, w),
           "embeddings.position_embeddings": sd["pos_embed"],
           "embeddings.patch_embeddings.projection.weight": sd["patch_embed.proj.weight"],
           "embeddings.patch_embeddings.projection.bias": sd["patch_embed.proj.bias"],
           "layernorm.weight": sd["norm.weight"], "layernorm.bias": sd["norm.bias"]}
    for l in range(cfg.depth):
        p, q = f"blocks.{l}.", f"encoder.layer.{l}."
        for i, n in enumerate(("query", "key", "value")):
            out[q + f"attention.attention.{n}.weight"] = sd[p + "attn.qkv.weight"][i * w:(i + 1) * w]
            out[q + f"attention.attention.{n}.bias"] = sd[p + "attn.qkv.bias"][i * w:(i + 1) * w]
        out[q + "attention.output.dense.weight"] = sd[p + "attn.proj.weight"]
        out[q + "attention.output.dense.bias"] = sd[p + "attn.proj.bias"]
        out[q + "layer_scale1.lambda1"] = sd[p + "ls1.gamma"]
        out[q + "layer_scale2.lambda1"] = sd[p + "ls2.gamma"]
        for n in ("norm1", "norm2"):
            out[q + n + ".weight"] = sd[p + n + ".weight"]
            out[q + n + ".bias"] = sd[p + n + ".bias"]
        for n in ("fc1", "fc2"):
            out[q + f"mlp.{n}.weight"] = sd[p + f"mlp.{n}.weight"]
            out[q + f"mlp.{n}.bias"] = sd[p + f"mlp.{n}.bias"]
    return out
