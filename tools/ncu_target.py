"""Small fixed workload for ncu: one launch of each hot kernel at the ViT-g / 128-slice shapes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200 import ops  # noqa: E402
from cryovit_b200.vit import interleave_w12  # noqa: E402

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T, C, Fh, H = 1029, 1536, 4096, 24
M = B * T
bf = dict(device=dev, dtype=torch.bfloat16)
ln = torch.randn(M, C, **bf)
qkv_w, qkv_b = torch.randn(3 * C, C, **bf) * C**-0.5, torch.randn(3 * C, device=dev)
qkv = torch.empty(M, 3 * C, **bf)
attn = torch.empty(M, C, **bf)
proj_w, proj_b, g = torch.randn(C, C, **bf) * C**-0.5, torch.randn(C, device=dev), torch.ones(C, device=dev)
x = torch.randn(M, C, device=dev)
w12, b12 = torch.randn(2 * Fh, C, **bf) * C**-0.5, torch.randn(2 * Fh, device=dev)
w12i, b12i = interleave_w12(w12, b12)
hidden = torch.empty(M, Fh, **bf)
w3, b3 = torch.randn(C, Fh, **bf) * Fh**-0.5, torch.randn(C, device=dev)
n1w, n1b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
for _ in range(2):
    ops.layernorm(x, n1w, n1b, ln, 1e-6)
    ops.linear_bias(ln, qkv_w, qkv_b, qkv)
    ops.attention(qkv, attn, B, T, H)
    ops.linear_scale_residual(attn, proj_w, proj_b, g, x)
    ops.linear_swiglu(ln, w12i, b12i, hidden)
    ops.linear_scale_residual(hidden, w3, b3, g, x)
torch.cuda.synchronize()
print("ok")
