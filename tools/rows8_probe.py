"""8 -> 8 full-resolution convolution: conv3d_rows8 (one voxel per MMA row) against conv3d_wpack8 (8 voxels per row, banded
weights) at the head's shape (128, 512, 512, 8): agreement and CUDA-event times."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200 import ops  # noqa: E402
from cryovit_b200.head import rows8_weight_image, wpack_weight_image  # noqa: E402

D, H, W = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (128, 512, 512)
g = torch.Generator().manual_seed(0)
x = torch.randn(D, H, W, 8, generator=g).bfloat16().cuda()
w = (torch.randn(8, 8, 3, 3, 3, generator=g) * (27 * 8) ** -0.5).bfloat16().cuda()
b = torch.randn(8, generator=g).cuda()
o1, o2 = torch.empty_like(x), torch.empty_like(x)
img_r, img_w = rows8_weight_image(w).bfloat16(), wpack_weight_image(w, 8).bfloat16()
b8 = b.repeat(8).contiguous()


def t(fn, n=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[n // 2]


for act in (1, 0):
    tr = t(lambda: ops.conv3d_rows8(x, img_r, b, o1, act=act))
    tw = t(lambda: ops.conv3d_wpack8_gelu(x, img_w, b8, o2, act=act))
    err = (o1.float() - o2.float()).abs().max().item()
    print(f"act {act}: rows8 {tr:.3f} ms | wpack8 {tw:.3f} ms | max abs diff {err:.3e}", flush=True)
w1 = (torch.randn(1, 8, 3, 3, 3, generator=g) * 0.3).bfloat16().cuda()
b1 = torch.randn(1, generator=g).cuda()
l1, p1, l2, p2 = (torch.empty(D, H, W, device="cuda") for _ in range(4))
if W % 16 == 0:
    img_rf, img_wf = rows8_weight_image(w1).bfloat16(), wpack_weight_image(w1, 16).bfloat16()
    tr = t(lambda: ops.conv3d_rows8_final(x, img_rf, b1, l1, p1))
    tw = t(lambda: ops.conv3d_wpack8_final(x, img_wf, b1.repeat(16).contiguous(), l2, p2))
    print(f"final: rows8 {tr:.3f} ms | wpack8 {tw:.3f} ms | max abs diff logits {(l1 - l2).abs().max().item():.3e} probs {(p1 - p2).abs().max().item():.3e}")
