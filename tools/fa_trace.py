"""Per-phase cycle accounting of the tcgen05 attention kernel: builds -DCVIT_FA_TRACE copies (extra nvcc flags can
be passed after --, e.g. `-- -DFA_TURNS=0`) and prints, per warp of CTAs 0-3, the average cycles per K/V tile spent
in each phase, plus the kernel time."""
import ctypes
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
extra = sys.argv[sys.argv.index("--") + 1:] if "--" in sys.argv else []
tag = "".join(c if c.isalnum() else "_" for c in "".join(extra))
so = ROOT / "tools" / f"_fa_trace{tag}.so"
if "--build" in sys.argv or not so.exists():
    csrc = ROOT / "cryovit_b200" / "csrc"
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
                           "--expt-relaxed-constexpr", "-DCVIT_FA_TRACE", *extra, "-shared", "-o", str(so),
                           str(csrc / "attention_tcgen05.cu"), str(csrc / "host_common.cu")])
    if "--build" in sys.argv:
        sys.exit(0)
lib = ctypes.CDLL(str(so))
B, T, H = 128, 1029, 24
C = H * 64
qkv = torch.randn(B * T, 3 * C, device="cuda", dtype=torch.bfloat16)
out = torch.empty(B * T, C, device="cuda", dtype=torch.bfloat16)
trace = torch.zeros(4 * 16 * 10, device="cuda", dtype=torch.int64)
lib.cvit_fa_set_trace(ctypes.c_void_p(trace.data_ptr()))
lib.cvit_attention_fwd_bf16.argtypes = [ctypes.c_void_p, ctypes.c_void_p] + [ctypes.c_int64] * 4 + [ctypes.c_void_p]
ts = []
for _ in range(4):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    rc = lib.cvit_attention_fwd_bf16(qkv.data_ptr(), out.data_ptr(), B, T, H, 64, None)
    e.record()
    assert rc == 0
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
print(f"flags {extra}: kernel ms (with accounting overhead): {min(ts):.3f}")
tr = trace.cpu().numpy().reshape(4, 16, 10)
sm_names = ["wait S", "ld", "max", "wait turn", "exp", "pv+st", "publish", "other"]
mma_names = ["w sfree", "w kv", "iss S", "w P", "w Oempty", "iss PV", "observe", "other"]
for cta in range(2):
    for w in list(range(8)) + [13, 14]:
        r = tr[cta, w]
        tiles = max(int(r[8]), 1)
        names = mma_names if w >= 13 else sm_names
        label = "MMA " if w >= 13 else f"{'AB'[w >> 2]}q{w & 3} "
        print(f"CTA{cta} {label} tiles {tiles:4d} per-tile:", "  ".join(f"{n} {r[i] / tiles:6.0f}" for i, n in enumerate(names)),
              f" | total {r[:8].sum() / tiles:6.0f} | {r[:8].sum() / max(int(r[9]), 1):.3f} GHz over {r[9] / 1e6:.3f} ms")
