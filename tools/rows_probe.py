"""16 / 32-channel dilated convolutions: conv3d_rows (one voxel per MMA row) against conv3d_wpackn at the head's shapes."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200 import ops  # noqa: E402
from cryovit_b200.head import rowsn_weight_image, wpackn_weight_image  # noqa: E402

g = torch.Generator().manual_seed(0)


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


for D, H, W, cin, cout, dil in [(128, 256, 256, 32, 16, 2), (128, 256, 256, 16, 16, 1), (128, 128, 128, 32, 32, 8), (128, 128, 128, 32, 32, 4),
                                (128, 256, 256, 16, 32, 2)]:
    x = torch.randn(D, H, W, cin, generator=g).bfloat16().cuda()
    w = (torch.randn(cout, cin, 3, 3, 3, generator=g) * (27 * cin) ** -0.5).bfloat16().cuda()
    b = torch.randn(cout, generator=g).cuda()
    o1, o2 = torch.empty(D, H, W, cout, device="cuda", dtype=torch.bfloat16), torch.empty(D, H, W, cout, device="cuda", dtype=torch.bfloat16)
    img_r = rowsn_weight_image(w).bfloat16()
    P = ops.wpackn_group(cin, cout)
    line = f"{cin}->{cout} d{dil} @{H}x{W}:"
    for act in (1, 0):
        tr = t(lambda: ops.conv3d_rows(x, img_r, b.repeat(64).contiguous(), o1, dil, act=act))
        line += f" act {act}: rows {tr:.3f} ms"
        if P:
            img_w = wpackn_weight_image(w, cout, P).bfloat16()
            tw = t(lambda: ops.conv3d_wpackn(x, img_w, b.repeat(64).contiguous(), o2, dil, cout, act=act))
            line += f" | wpackn {tw:.3f} ms | diff {(o1.float() - o2.float()).abs().max().item():.2e};"
    print(line, flush=True)
