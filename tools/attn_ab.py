"""A/B of attention builds at the ViT-g shapes: kernel alone and inside a GEMM-attention-GEMM sequence (the power state
of the real step), accuracy vs fp32 SDPA on 4 slices. CRYOVIT_B200_LIB selects the library, CVIT_FA_EXACT=1 the exact pass.

    python tools/attn_ab.py [scale=1.0] [out_dtype=bf16|fp16]
"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from cryovit_b200 import _lib, ops  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
odt = torch.float16 if len(sys.argv) > 2 and sys.argv[2] == "fp16" else torch.bfloat16
B, T, H, C, Fh = 128, 1029, 24, 1536, 4096
M = B * T
qkv = (torch.randn(M, 3 * C, device="cuda") * scale).bfloat16()
out = torch.empty(M, C, device="cuda", dtype=odt)
ln = torch.randn(M, C, device="cuda", dtype=torch.bfloat16)
w = torch.randn(3 * C, C, device="cuda", dtype=torch.bfloat16) * C**-0.5
b = torch.zeros(3 * C, device="cuda")
scratch = torch.empty(M, 3 * C, device="cuda", dtype=torch.bfloat16)


def med(ts):
    return sorted(ts)[len(ts) // 2]


def time_alone(n=9):
    for _ in range(3):
        ops.attention(qkv, out, B, T, H)
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.attention(qkv, out, B, T, H)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return med(ts), min(ts)


def time_in_sequence(n=40, warm=40):
    evs = []
    for i in range(warm + n):
        ops.linear_bias(ln, w, b, scratch)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        ops.attention(qkv, out, B, T, H)
        e.record()
        ops.linear_bias(ln, w, b, scratch)
        if i >= warm:
            evs.append((s, e))
    torch.cuda.synchronize()
    return med([s.elapsed_time(e) for s, e in evs])


alone, best = time_alone()
seq = time_in_sequence()
q, k, v = qkv[: 4 * T].view(4, T, 3, H, 64).permute(2, 0, 3, 1, 4).float()
ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(4 * T, H * 64)
err = (out[: 4 * T].float() - ref).abs().max().item()
tag = f"lib={Path(os.environ.get('CRYOVIT_B200_LIB', 'product')).name} exact={os.environ.get('CVIT_FA_EXACT', '0')} scale={scale} out={odt}"
print(f"[{tag}] alone {alone:.4f} ms (min {best:.4f}) | between GEMMs {seq:.4f} ms | max abs err {err:.3e}", flush=True)
if os.environ.get("ATTN_AB_SDPA"):
    qq, kk, vv = qkv.view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ts = []
    for _ in range(6):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    evs = []
    for i in range(60):
        ops.linear_bias(ln, w, b, scratch)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)
        e.record()
        ops.linear_bias(ln, w, b, scratch)
        if i >= 30:
            evs.append((s, e))
    torch.cuda.synchronize()
    print(f"torch SDPA alone {med(ts):.4f} ms | between GEMMs {med([s.elapsed_time(e) for s, e in evs]):.4f} ms", flush=True)
