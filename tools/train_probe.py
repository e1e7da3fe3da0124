"""Per-launch timing of one CryoVIT head training step at BASELINE config 5's crop (features (1536,128,32,32), labels
(128,512,512)): an event pair around every call across the C ABI in one eager forward + backward, then the step time
with eager launches and with the CUDA graph."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200 import _lib, build  # noqa: E402
from cryovit_b200.train import CryoVITHeadTrainerB200  # noqa: E402

build.build()
C, D, h, w = 1536, int(sys.argv[1]) if len(sys.argv) > 1 else 128, 32, 32
g = torch.Generator().manual_seed(0)
feats = (torch.randn(C, D, h, w, generator=g) * 0.5).half().cuda()
labels = (torch.rand(D, 16 * h, 16 * w, generator=g) < 0.1).float()
labels[::5] = -1
labels = labels.cuda()
tr = CryoVITHeadTrainerB200(C)
orig, active, seq = _lib.call, False, []


def call(name, *a):
    if not active:
        return orig(name, *a)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    r = orig(name, *a)
    e.record()
    seq.append((name[5:], tuple(int(x) for x in a[:-1] if isinstance(x, int) and abs(x) < (1 << 40)), s, e))
    return r


_lib.call = call
os.environ["CVIT_TRAIN_GRAPH"] = "0"  # per-kernel events need eager launches
for _ in range(2):
    loss = tr.train_step(feats, labels)
torch.cuda.synchronize()
print("loss", float(loss), "peak memory GB", torch.cuda.max_memory_allocated() / 1e9)
active = True
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
tr.train_step(feats, labels)
e.record()
torch.cuda.synchronize()
active = False
agg = {}
for n, dims, a, b in seq:
    t = agg.setdefault((n, dims), [0, 0.0])
    t[0] += 1
    t[1] += a.elapsed_time(b)
byname = {}
for (n, dims), (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:28s} {str(dims):60s} x{cnt:3d} {ms:9.3f} ms")
    t = byname.setdefault(n, [0, 0.0])
    t[0] += cnt
    t[1] += ms
print("--- by entry point")
for n, (cnt, ms) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:28s} x{cnt:3d} {ms:9.3f} ms")
print(f"instrumented step {s.elapsed_time(e):.2f} ms (our kernels {sum(v[1] for v in agg.values()):.2f} ms over {len(seq)} launches)")
ts = []
for _ in range(3):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    tr.train_step(feats, labels)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print(f"train step, eager launches {sorted(ts)[1]:.2f} ms -> {D * 512 * 512 / sorted(ts)[1] / 1e6:.3f} Gvoxel/s")
os.environ["CVIT_TRAIN_GRAPH"] = "1"  # forward + backward captured into a CUDA graph on the third step of a shape
ts = []
for _ in range(7):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    loss = tr.train_step(feats, labels)
    e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print(f"train step, CUDA graph {sorted(ts[3:])[2]:.2f} ms -> {D * 512 * 512 / sorted(ts[3:])[2] / 1e6:.3f} Gvoxel/s (loss {float(loss):.4f}, "
      f"graph active: {any('graph' in v for v in tr._graphs.values())})")
