"""fp32 restatement of the CryoVIT 3-D head (TEST INFRASTRUCTURE, see oracle/__init__.py).

Follows src/cryovit/models/cryovit.py:
  CryoVIT.__init__      :18-34  Conv3d(1536,1024,1)+GELU; SynthesisBlock(1024,192,128,32,24), (128,64,32,16,12),
                                (32,32,32,8,4), (32,16,8,2,1); output Conv3d(8,8,3)+GELU+Conv3d(8,1,3)
  forward_volume        :36-40  layers -> output_layer -> clip(-5, 5)
  forward               :42-49  (B,D,C,H,W) -> permute (B,C,D,H,W) -> forward_volume -> squeeze(1) -> sigmoid
  SynthesisBlock        :52-83  GroupNorm(max(8, c1//8), c1, eps=1e-3) -> Conv3d(c1,c2,3,same,dil=(d1,1,1)) -> GELU
                                -> Conv3d(c2,c2,3,same,dil=(d2,1,1)) -> GELU -> ConvTranspose3d(c2,c3,(1,2,2),
                                stride=(1,2,2)) -> GELU
Parameters come as a state dict with the reference's names (layers.0.weight, layers.2.layers.1.weight, ...,
output_layer.2.bias). ``in_channels`` generalises the hard-wired 1536 (BASELINE config 1 uses ViT-S, 384).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BLOCKS = [(1024, 192, 128, 32, 24), (128, 64, 32, 16, 12), (32, 32, 32, 8, 4), (32, 16, 8, 2, 1)]


def random_state_dict(in_channels: int = 1536, seed: int = 0) -> dict:
    """Seeded parameters with torch's default Conv3d / ConvTranspose3d / GroupNorm initialisation order."""
    torch.manual_seed(seed)
    import torch.nn as nn

    sd = {}

    def put(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[prefix + k] = v.detach().clone()

    put("layers.0.", nn.Conv3d(in_channels, 1024, 1, padding="same"))
    for bi, (c1, c2, c3, d1, d2) in enumerate(BLOCKS):
        p = f"layers.{bi + 2}.layers."
        put(p + "0.", nn.GroupNorm(max(8, c1 // 8), c1, eps=1e-3))
        put(p + "1.", nn.Conv3d(c1, c2, 3, padding="same", dilation=(d1, 1, 1)))
        put(p + "3.", nn.Conv3d(c2, c2, 3, padding="same", dilation=(d2, 1, 1)))
        put(p + "5.", nn.ConvTranspose3d(c2, c3, (1, 2, 2), stride=(1, 2, 2)))
    put("output_layer.0.", nn.Conv3d(8, 8, 3, padding="same"))
    put("output_layer.2.", nn.Conv3d(8, 1, 3, padding="same"))
    return sd


@torch.no_grad()
def forward_volume(sd: dict, x: torch.Tensor, return_layers: bool = False):
    """x: f32 [B, C, D, h, w] -> clipped logits [B, 1, D, 16h, 16w]."""
    sd = {k: v.float() for k, v in sd.items()}
    layers = []
    x = F.gelu(F.conv3d(x.float(), sd["layers.0.weight"], sd["layers.0.bias"]))
    layers.append(x)
    for bi, (c1, c2, c3, d1, d2) in enumerate(BLOCKS):
        p = f"layers.{bi + 2}.layers."
        x = F.group_norm(x, max(8, c1 // 8), sd[p + "0.weight"], sd[p + "0.bias"], eps=1e-3)
        x = F.gelu(F.conv3d(x, sd[p + "1.weight"], sd[p + "1.bias"], padding="same", dilation=(d1, 1, 1)))
        x = F.gelu(F.conv3d(x, sd[p + "3.weight"], sd[p + "3.bias"], padding="same", dilation=(d2, 1, 1)))
        x = F.gelu(F.conv_transpose3d(x, sd[p + "5.weight"], sd[p + "5.bias"], stride=(1, 2, 2)))
        layers.append(x)
    x = F.gelu(F.conv3d(x, sd["output_layer.0.weight"], sd["output_layer.0.bias"], padding="same"))
    x = F.conv3d(x, sd["output_layer.2.weight"], sd["output_layer.2.bias"], padding="same")
    x = torch.clip(x, -5.0, 5.0)
    return (x, layers) if return_layers else x


@torch.no_grad()
def forward(sd: dict, tomo_batch: torch.Tensor) -> torch.Tensor:
    """tomo_batch: [B, D, C, h, w] (collate_fn layout) -> probabilities [B, D, 16h, 16w]."""
    x = tomo_batch.permute(0, 2, 1, 3, 4)
    return torch.sigmoid(forward_volume(sd, x).squeeze(1))
