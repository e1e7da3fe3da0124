"""fp32 restatement of DINOv2-with-registers ``forward_features`` (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows upstream facebookresearch/dinov2 (un-vendored; reference call site run/dino_features.py:58):
  dinov2/models/vision_transformer.py  prepare_tokens_with_masks, interpolate_pos_encoding, forward_features
  dinov2/layers/patch_embed.py         Conv2d(3, C, 14, 14) -> flatten -> transpose
  dinov2/layers/attention.py           qkv.reshape(B,N,3,H,C/H); softmax(q*scale @ k^T) @ v; proj
  dinov2/layers/block.py               x + ls1(attn(norm1(x))); x + ls2(mlp(norm2(x)))
  dinov2/layers/mlp.py / swiglu_ffn.py fc1-GELU-fc2  /  w12 -> silu(x1)*x2 -> w3
  dinov2/layers/layer_scale.py         x * gamma
  dinov2/hub/backbones.py              *_reg: num_register_tokens=4, interpolate_antialias=True,
                                       interpolate_offset=0.0, img_size=518, patch 14, init_values=1.0
Cross-check: transformers/models/dinov2_with_registers/modeling_dinov2_with_registers.py ("HF:").
Parameters come as a state dict with the upstream names.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def interpolate_pos_encoding(pos_embed: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """[1, 1+M*M, C] -> [1, 1+gh*gw, C]; bicubic, antialias, size-based (interpolate_offset == 0.0)."""
    n = pos_embed.shape[1] - 1
    m = int(math.sqrt(n))
    assert m * m == n
    if gh == m and gw == m:
        return pos_embed.float()
    pe = pos_embed.float()
    cls_pos, patch_pos = pe[:, :1], pe[:, 1:]
    dim = pe.shape[-1]
    patch_pos = F.interpolate(patch_pos.reshape(1, m, m, dim).permute(0, 3, 1, 2), size=(gh, gw), mode="bicubic",
                              antialias=True)
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, gh * gw, dim)
    return torch.cat([cls_pos, patch_pos], dim=1)


def _attention(x, sd, p, heads):
    B, N, C = x.shape
    qkv = F.linear(x, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]).reshape(B, N, 3, heads, C // heads)
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    attn = (q * (C // heads) ** -0.5) @ k.transpose(-2, -1)
    attn = attn.softmax(dim=-1)
    out = (attn @ v).transpose(1, 2).reshape(B, N, C)
    return F.linear(out, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])


def _ffn(x, sd, p):
    if p + "mlp.w12.weight" in sd:
        x12 = F.linear(x, sd[p + "mlp.w12.weight"], sd[p + "mlp.w12.bias"])
        x1, x2 = x12.chunk(2, dim=-1)
        return F.linear(F.silu(x1) * x2, sd[p + "mlp.w3.weight"], sd[p + "mlp.w3.bias"])
    h = F.gelu(F.linear(x, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    return F.linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])


@torch.no_grad()
def forward_features(sd: dict, x: torch.Tensor, num_heads: int, eps: float = 1e-6, return_blocks: bool = False) -> dict:
    """x: f32 [B, 3, H, W] with H, W multiples of 14."""
    sd = {k: v.float() for k, v in sd.items()}
    B, _, H, W = x.shape
    C = sd["cls_token"].shape[-1]
    R = sd["register_tokens"].shape[1]
    t = F.conv2d(x.float(), sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=14)
    gh, gw = t.shape[-2:]
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], dim=1)
    t = t + interpolate_pos_encoding(sd["pos_embed"], gh, gw)
    t = torch.cat([t[:, :1], sd["register_tokens"].expand(B, -1, -1), t[:, 1:]], dim=1)
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    per_block = []
    for i in range(depth):
        p = f"blocks.{i}."
        t = t + sd[p + "ls1.gamma"] * _attention(F.layer_norm(t, (C,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps), sd, p, num_heads)
        t = t + sd[p + "ls2.gamma"] * _ffn(F.layer_norm(t, (C,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps), sd, p)
        if return_blocks:
            per_block.append(t.clone())
    xn = F.layer_norm(t, (C,), sd["norm.weight"], sd["norm.bias"], eps)
    out = {"x_norm_clstoken": xn[:, 0], "x_norm_regtokens": xn[:, 1:R + 1], "x_norm_patchtokens": xn[:, R + 1:],
           "x_prenorm": t, "masks": None}
    if return_blocks:
        out["blocks"] = per_block
    return out


class OracleDino:
    """Object with the surface the reference uses of the hub model: .cuda(), .eval(), .forward_features()."""

    def __init__(self, sd: dict, num_heads: int):
        self.sd, self.num_heads = sd, num_heads

    def cuda(self):
        return self

    def eval(self):
        return self

    def forward_features(self, x):
        return forward_features(self.sd, x, self.num_heads)
