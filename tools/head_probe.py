"""Per-layer timing of the CryoVIT 3-D head at BASELINE config 4 (1536-ch feature volume 128x32x32 -> 128x512x512)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200 import build, ops  # noqa: E402
from cryovit_b200.head import CryoVITHeadB200  # noqa: E402
from oracle import head as ohead  # noqa: E402  (only for the seeded random state dict)

build.build()
C, D, h, w = 1536, int(sys.argv[1]) if len(sys.argv) > 1 else 128, 32, 32
head = CryoVITHeadB200(C).load_state_dict(ohead.random_state_dict(C, seed=0)).cuda()
feats = (torch.randn(C, D, h, w, device="cuda") * 0.5).half()

records = {}
names = ["features_to_ndhwc", "linear_bias", "linear_bias_cfirst", "groupnorm_ndhwc", "conv3d_dilated", "conv3d_halo", "convT_1x2x2", "head_out_conv", "conv3d_wpack8_gelu", "conv3d_wpack8_final"]
orig = {n: getattr(ops, n) for n in names}
active = False
seq = []


def wrap(name):
    def f(*a, **k):
        if not active:
            return orig[name](*a, **k)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig[name](*a, **k)
        e.record()
        shape = tuple(a[0].shape)
        seq.append((name, shape, s, e))
        return r
    return f


for n in names:
    setattr(ops, n, wrap(n))

for _ in range(2):
    head.segment_volume(feats)
torch.cuda.synchronize()
active = True
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
head.segment_volume(feats)
e.record()
torch.cuda.synchronize()
active = False
total = s.elapsed_time(e)
for name, shape, a, b in seq:
    print(f"{name:20s} {str(shape):28s} {a.elapsed_time(b):8.3f} ms")
ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    head.segment_volume(feats)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ms = sorted(ts)[2]
vox = D * 16 * h * 16 * w
print(f"head total {ms:.3f} ms -> {vox / ms / 1e6:.2f} Gvoxel/s, {94864 * vox / ms / 1e9:.1f} TFLOP/s")
