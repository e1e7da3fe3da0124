"""A/B of attention builds: time the tcgen05 attention kernel at the ViT-g shapes and report accuracy vs torch SDPA."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200 import build, ops  # noqa: E402

build.build()
B, T, H = 128, 1029, 24
qkv = (torch.randn(B * T, 3 * H * 64, device="cuda") * 1.5).bfloat16()
out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attention(qkv, out, B, T, H)
torch.cuda.synchronize()
ts = []
for _ in range(9):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.attention(qkv, out, B, T, H)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
q, k, v = qkv[: 4 * T].view(4, T, 3, H, 64).permute(2, 0, 3, 1, 4).float()
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(4 * T, H * 64)
err = (out[: 4 * T].float() - ref).abs().max().item()
import os
from cryovit_b200 import _lib
print(f"attention ({'exact' if os.environ.get('CVIT_FA_EXACT') == '1' else 'fast'} pass) {sorted(ts)[4]:.4f} ms (min {min(ts):.4f}), "
      f"max abs err vs fp32 SDPA {err:.3e}, exact-pass items {_lib.load().cvit_attention_redo_items()}")
qq, kk, vv = qkv.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
ts = []
for _ in range(6):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print(f"torch SDPA {sorted(ts)[3]:.4f} ms")
