"""Per-kernel timing probe at the ViT-g / 128-slice shapes (CUDA events, L2-exceeding operands)."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200 import build, ops  # noqa: E402
from cryovit_b200.vit import interleave_w12  # noqa: E402

build.build()
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T, C, Fh, H = 1029, 1536, 4096, 24
M = B * T


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


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

res = {}
def rec(name, ms, flops=None, bytes_=None):
    r = {"ms": round(ms, 4)}
    if flops:
        r["tflops"] = round(flops / ms / 1e9, 1)
    if bytes_:
        r["gbs"] = round(bytes_ / ms / 1e6, 1)
    res[name] = r
    print(name, r, flush=True)

from cryovit_b200 import _lib  # noqa: E402
lib = _lib.load()
gemms = {
    "qkv_gemm": (lambda: ops.linear_bias(ln, qkv_w, qkv_b, qkv), 2 * M * C * 3 * C),
    "proj_gemm_resid": (lambda: ops.linear_scale_residual(attn, proj_w, proj_b, g, x), 2 * M * C * C),
    "w12_swiglu": (lambda: ops.linear_swiglu(ln, w12i, b12i, hidden), 2 * M * C * 2 * Fh),
    "w3_gemm_resid": (lambda: ops.linear_scale_residual(hidden, w3, b3, g, x), 2 * M * Fh * C),
}
for pair in (1, 0):
    lib.cvit_set_gemm_pair(pair)
    for name, (fn, fl) in gemms.items():
        rec(name + ("_pair" if pair else "_single"), timeit(fn), fl)
lib.cvit_set_gemm_pair(1)
rec("cublas_qkv", timeit(lambda: torch.addmm(qkv_b.bfloat16(), ln, qkv_w.t())), 2 * M * C * 3 * C)
rec("attention_tcgen05", timeit(lambda: ops.attention(qkv, attn, B, T, H)), 4 * B * H * T * T * 64)
rec("attention_mma_sync", timeit(lambda: ops.attention(qkv, attn, B, T, H, legacy_mma_sync=True)), 4 * B * H * T * T * 64)
rec("layernorm", timeit(lambda: ops.layernorm(x, n1w, n1b, ln, 1e-6)), None, M * C * 6)
q, k, v = qkv.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
rec("torch_sdpa", timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v)), 4 * B * H * T * T * 64)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/kernel_probe.json").write_text(json.dumps(res, indent=1))
