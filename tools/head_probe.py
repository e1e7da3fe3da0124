"""Per-launch timing of the CryoVIT 3-D head at BASELINE config 4 (1536-ch feature volume 128x32x32 -> 128x512x512):
an event pair around every call across the C ABI.  usage: python tools/head_probe.py [D=128] [--nofuse]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200 import _lib, build  # noqa: E402
from cryovit_b200.head import CryoVITHeadB200  # noqa: E402
from oracle import head as ohead  # noqa: E402  (only for the seeded random state dict)

build.build()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
C, D, h, w = 1536, int(args[0]) if args else 128, 32, 32
head = CryoVITHeadB200(C, fuse_groupnorm="--nofuse" not in sys.argv).load_state_dict(ohead.random_state_dict(C, seed=0)).cuda()
feats = (torch.randn(C, D, h, w, device="cuda") * 0.5).half()
orig, active, seq = _lib.call, False, []


def call(name, *a):
    if not active:
        return orig(name, *a)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    r = orig(name, *a)
    e.record()
    seq.append((name, tuple(int(x) for x in a[:-1] if isinstance(x, int) and abs(x) < (1 << 40)), s, e))
    return r


_lib.call = call
for _ in range(2):
    head.segment_volume(feats)
torch.cuda.synchronize()
active = True
head.segment_volume(feats)
torch.cuda.synchronize()
active = False
for name, dims, a, b in seq:
    print(f"{name[5:]:34s} {str(dims):52s} {a.elapsed_time(b):8.3f} ms")
print(f"sum of launches {sum(a.elapsed_time(b) for _, _, a, b in seq):.3f} ms over {len(seq)} calls")
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
print(f"head total ({'fused' if head.fuse_groupnorm else 'two-pass'} GroupNorm) {ms:.3f} ms -> {vox / ms / 1e6:.2f} Gvoxel/s, {94864 * vox / ms / 1e9:.1f} TFLOP/s")
