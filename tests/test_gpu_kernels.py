"""Kernel-level parity on a B200: every C-ABI operator against a plain PyTorch fp32 restatement of the same
operator evaluated on the same (bf16-rounded) operands. Tolerances are stated per test; they bound the bf16
output rounding (2^-9 relative) plus fp32 accumulation-order differences."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda"


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


def _close(got, ref, atol, rtol, what):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = (err > tol).sum().item()
    assert bad == 0, f"{what}: {bad}/{err.numel()} out of tolerance, max abs err {err.max().item():.4e}, ref max {ref.abs().max().item():.3e}"


@pytest.mark.parametrize("M,N,K", [(300, 512, 256), (128, 256, 64), (1029, 384, 384), (257, 1152, 384), (4116, 4608, 1536)])
@pytest.mark.parametrize("gelu", [False, True])
def test_linear_bias(cuda_lib, M, N, K, gelu):
    from cryovit_b200 import ops
    a = _rand(M, K, seed=1).bfloat16()
    w = _rand(N, K, scale=K ** -0.5, seed=2).bfloat16()
    b = _rand(N, seed=3)
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.linear_bias(a, w, b, out, gelu=gelu)
    ref = a.float() @ w.float().t() + b
    if gelu:
        ref = F.gelu(ref)
    _close(out, ref, atol=2e-2, rtol=1e-2, what=f"linear_bias {M}x{N}x{K} gelu={gelu}")


@pytest.mark.parametrize("M,H,K", [(300, 256, 128), (1029, 4096, 1536)])
def test_linear_swiglu(cuda_lib, M, H, K):
    from cryovit_b200 import ops
    from cryovit_b200.vit import interleave_w12
    a = _rand(M, K, seed=1).bfloat16()
    w12 = _rand(2 * H, K, scale=K ** -0.5, seed=2).bfloat16()
    b12 = _rand(2 * H, seed=3)
    w12i, b12i = interleave_w12(w12, b12)
    out = torch.full((M, H), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.linear_swiglu(a, w12i, b12i, out)
    y = a.float() @ w12.float().t() + b12
    ref = F.silu(y[:, :H]) * y[:, H:]
    _close(out, ref, atol=2e-2, rtol=1e-2, what="linear_swiglu")


@pytest.mark.parametrize("M,N,K", [(300, 256, 128), (1029, 1536, 4096), (789, 384, 1536)])
def test_linear_scale_residual(cuda_lib, M, N, K):
    from cryovit_b200 import ops
    a = _rand(M, K, seed=1).bfloat16()
    w = _rand(N, K, scale=K ** -0.5, seed=2).bfloat16()
    b, g = _rand(N, seed=3), _rand(N, seed=4)
    x0 = _rand(M, N, seed=5)
    x = x0.clone()
    ops.linear_scale_residual(a, w, b, g, x)
    ref = x0 + g * (a.float() @ w.float().t() + b)
    _close(x, ref, atol=1e-3, rtol=1e-4, what="linear_scale_residual")


@pytest.mark.parametrize("M", [129, 300, 1029, 4116, 20000])
def test_gemm_cta_pair_matches_single_cta(cuda_lib, M):
    """The cta_group::2 kernel (256 x 256 tile per CTA pair; the default for the ViT linears) and the single-CTA kernel
    accumulate every output element over K in the same order, so the two modes must agree BIT FOR BIT, for every
    fused epilogue, including row counts whose last pair tile is ragged (129: the second CTA's rows are all out of
    range except one; 1029 = 4*256 + 5: the second CTA of the last pair has no rows at all)."""
    from cryovit_b200 import ops
    from cryovit_b200.vit import interleave_w12
    K, N = 384, 768
    a = _rand(M, K, seed=1).bfloat16()
    w = _rand(N, K, scale=K ** -0.5, seed=2).bfloat16()
    b, g, x0 = _rand(N, seed=3), _rand(N, seed=4), _rand(M, N, seed=5)
    wi, bi = interleave_w12(w, b)

    def run():
        o1 = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
        o2 = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
        o3 = torch.full((M, N // 2), float("nan"), device=DEV, dtype=torch.bfloat16)
        x = x0.clone()
        ops.linear_bias(a, w, b, o1)
        ops.linear_bias(a, w, b, o2, gelu=True)
        ops.linear_swiglu(a, wi, bi, o3)
        ops.linear_scale_residual(a, w, b, g, x)
        torch.cuda.synchronize()
        return o1, o2, o3, x

    assert cuda_lib.cvit_set_gemm_pair(0) == 1  # pairs are the default
    try:
        single = run()
    finally:
        cuda_lib.cvit_set_gemm_pair(1)
    pair = run()
    for s_, p_, what in zip(single, pair, ("bias", "bias_gelu", "swiglu", "scale_residual")):
        assert torch.equal(s_, p_), f"{what}: pair and single-CTA kernels differ at M={M}"
    _close(pair[0], a.float() @ w.float().t() + b, atol=2e-2, rtol=1e-2, what="pair linear_bias vs fp32")


@pytest.mark.parametrize("M,N,K", [(300, 512, 256), (1029, 384, 384), (4116, 4608, 1536)])
def test_linears_fp16_operands(cuda_lib, M, N, K):
    """The fp16-operand format of the three ViT linears (CVIT_FMT_OPERANDS_F16 / CVIT_FMT_OUT_F16): same kernels, IEEE
    fp16 A and W, output fp16 (qkv), bf16 (FFN hidden keeps the fp32 range) or the fp32 residual. The tolerance on the
    fp16 output is 8x tighter than the bf16 one above (2^-11 relative rounding); mixing the two operand types is refused
    on the host (a bf16 x fp16 MMA is an illegal instruction on sm_100a)."""
    from cryovit_b200 import ops
    from cryovit_b200._lib import CryovitB200Error
    from cryovit_b200.vit import interleave_w12
    a = _rand(M, K, seed=1).half()
    w = _rand(N, K, scale=K ** -0.5, seed=2).half()
    b, g, x0 = _rand(N, seed=3), _rand(N, seed=4), _rand(M, N, seed=5)
    y = a.float() @ w.float().t() + b
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.float16)
    ops.linear_bias(a, w, b, out)
    _close(out, y, atol=3e-3, rtol=1.5e-3, what="linear_bias fp16 -> fp16")
    out_g = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.linear_bias(a, w, b, out_g, gelu=True)
    _close(out_g, F.gelu(y), atol=2e-2, rtol=1e-2, what="linear_bias+gelu fp16 -> bf16")
    if N % 256 == 0:  # SwiGLU tiles interleave 128 rows of w1 with 128 of w2
        wi, bi = interleave_w12(w, b)
        out_s = torch.full((M, N // 2), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.linear_swiglu(a, wi, bi, out_s)
        _close(out_s, F.silu(y[:, :N // 2]) * y[:, N // 2:], atol=2e-2, rtol=1e-2, what="linear_swiglu fp16 -> bf16")
    x = x0.clone()
    ops.linear_scale_residual(a, w, b, g, x)
    _close(x, x0 + g * y, atol=1e-3, rtol=1e-4, what="linear_scale_residual fp16")
    with pytest.raises(CryovitB200Error):
        ops.linear_bias(a, w.bfloat16(), b, out)


@pytest.mark.parametrize("K,M,N", [(384, 4104, 1024), (1536, 200, 1024), (1536, 136, 1024), (128, 64, 256), (1536, 16384, 1024), (192, 520, 128)])
@pytest.mark.parametrize("gelu", [True, False])
def test_linear_bias_cfirst(cuda_lib, K, M, N, gelu):
    """The head's projection with A read from its transposed storage (the (C, D*h*w) feature layout) as an MN-major
    tensor-core operand: row counts that are not multiples of the 64-row TMA box or of the 256-row pair tile, K not a
    multiple of the 64-channel stage (192: the last stage is half zero fill)."""
    from cryovit_b200 import ops
    at = _rand(K, M, scale=0.7, seed=1).half()
    w = _rand(N, K, scale=K ** -0.5, seed=2).half()
    b = _rand(N, seed=3)
    out = torch.full((M, N), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.linear_bias_cfirst(at, w, b, out, gelu=gelu)
    ref = at.float().t() @ w.float().t() + b
    if gelu:
        ref = F.gelu(ref)
    _close(out, ref, atol=2e-2, rtol=1e-2, what=f"linear_bias_cfirst K={K} M={M} N={N}")


def test_patch_embed_gemm(cuda_lib):
    from cryovit_b200 import ops
    B, Np, T, C, K = 3, 64, 69, 384, 256
    patches = _rand(B * Np, K, seed=1).bfloat16()
    w = _rand(C, K, scale=K ** -0.5, seed=2).bfloat16()
    table = _rand(Np, C, seed=3)
    x = torch.zeros(B * T, C, device=DEV)
    ops.patch_embed_gemm(patches, w, table, x, B, Np, T, 5)
    ref = torch.zeros(B, T, C, device=DEV)
    ref[:, 5:] = (patches.float() @ w.float().t()).view(B, Np, C) + table
    _close(x.view(B, T, C), ref, atol=1e-3, rtol=1e-4, what="patch_embed")


@pytest.mark.parametrize("C", [384, 1536])
def test_layernorm(cuda_lib, C):
    from cryovit_b200 import ops
    M = 1031
    x = _rand(M, C, scale=3.0, seed=1) + 0.5
    g, b = _rand(C, seed=2), _rand(C, seed=3)
    out = torch.empty(M, C, device=DEV, dtype=torch.bfloat16)
    ops.layernorm(x, g, b, out, 1e-6)
    ref = F.layer_norm(x, (C,), g, b, 1e-6)
    _close(out, ref, atol=1e-2, rtol=1e-2, what="layernorm")
    out16 = torch.empty(M, C, device=DEV, dtype=torch.float16)
    ops.layernorm(x, g, b, out16, 1e-6)
    _close(out16, ref, atol=1e-3, rtol=1e-3, what="layernorm fp16")


@pytest.mark.parametrize("B,T,H", [(2, 1029, 3), (1, 789, 6), (2, 70, 2), (3, 29, 2), (1, 128, 1), (1, 256, 2), (2, 261, 1),
                                   (16, 300, 8), (6, 1029, 6), (40, 120, 8),  # more work items than SMs: persistent path
                                   (2, 133, 2), (2, 264, 2), (3, 385, 1), (1, 517, 2)])  # 1..8 trailing keys: epilogue path
@pytest.mark.parametrize("legacy", [False, True])
@pytest.mark.parametrize("scale", [1.0, 6.0])
def test_attention(cuda_lib, B, T, H, legacy, scale):
    from cryovit_b200 import ops
    C = H * 64
    # scale 6 makes the logits span +-100: exercises the running-max / lazy-rescale path
    qkv = (_rand(B * T, 3 * C, seed=1) * scale).bfloat16()
    out = torch.full((B * T, C), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.attention(qkv, out, B, T, H, legacy_mma_sync=legacy)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, C)
    _close(out, ref, atol=2e-2 * scale, rtol=2e-2, what=f"attention legacy={legacy}")


@pytest.mark.parametrize("B,T,H", [(2, 1029, 3), (3, 261, 2), (1, 1024, 2), (16, 300, 8)])
@pytest.mark.parametrize("scale", [1.0, 6.0, 20.0])
def test_attention_bf16_in_fp16_out(cuda_lib, B, T, H, scale):
    """The default ViT format: bf16 q/k/v and probabilities, output stored as fp16 for the projection GEMM. scale 20
    makes the scores span +-1000 between tiles: lazy rescales (threshold 2^16 with bf16 probabilities) on nearly every
    tile."""
    from cryovit_b200 import ops
    C = H * 64
    qkv = (_rand(B * T, 3 * C, seed=1) * scale).bfloat16()
    out = torch.full((B * T, C), float("nan"), device=DEV, dtype=torch.float16)
    ops.attention(qkv, out, B, T, H)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, C)
    _close(out, ref, atol=2e-2 * scale, rtol=2e-2, what="attention bf16 -> fp16")


def test_attention_extreme_scores(cuda_lib):
    """One key per (slice, head) whose score exceeds every other by ~12800 (q = k = 40 on all 64 dimensions), sitting in
    a middle tile, the ragged tile or the first tile: the running maximum jumps by 2^18000 inside one tile; the lazy
    rescale must follow (exp2 of the old maximum underflows to exactly 0) and the result is that key's v."""
    from cryovit_b200 import ops
    B, T, H = 4, 700, 3
    C = H * 64
    qkv = (_rand(B * T, 3 * C, seed=3) * 0.5).bfloat16().view(B, T, 3, H, 64)
    for T in (700, 645):  # 645 = 5 * 128 + 5: the trailing five keys are merged by the epilogue warps (fp32, CUDA cores)
        qkv = (_rand(B * T, 3 * C, seed=3) * 0.5).bfloat16().view(B, T, 3, H, 64)
        for b, h, key in [(1, 0, 300), (3, 2, T - 1), (2, 1, 5), (0, 1, T - 4)]:
            qkv[b, :, 0, h, :] = 40.0
            qkv[b, key, 1, h, :] = 40.0
        qkv = qkv.view(B * T, 3 * C).contiguous()
        q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
        ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, C)
        for odt in (torch.bfloat16, torch.float16):
            out = torch.full((B * T, C), float("nan"), device=DEV, dtype=odt)
            ops.attention(qkv, out, B, T, H)
            _close(out, ref, atol=2e-2, rtol=2e-2, what=f"attention with one dominant key, T {T}, out {odt}")


@pytest.mark.parametrize("B,T,H", [(2, 1029, 3), (1, 789, 6), (3, 29, 2), (2, 261, 1), (16, 300, 8), (2, 136, 2)])
@pytest.mark.parametrize("scale", [1.0, 6.0])
def test_attention_fp16(cuda_lib, B, T, H, scale):
    """Same kernel with fp16 q/k/v, probabilities and output; the tolerance is 4x tighter than the bf16 one."""
    from cryovit_b200 import ops
    C = H * 64
    qkv = (_rand(B * T, 3 * C, seed=1) * scale).half()
    out = torch.full((B * T, C), float("nan"), device=DEV, dtype=torch.float16)
    ops.attention(qkv, out, B, T, H)
    q, k, v = qkv.float().view(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, C)
    _close(out, ref, atol=5e-3 * scale, rtol=5e-3, what="attention fp16")


@pytest.mark.parametrize("D,H,W,u8", [(3, 64, 96, True), (2, 50, 70, True), (2, 64, 64, False)])
def test_preproc_patchify(cuda_lib, D, H, W, u8):
    from cryovit_b200 import ops
    g = torch.Generator().manual_seed(0)
    if u8:
        src = torch.randint(0, 256, (D, H, W), generator=g, dtype=torch.uint8).to(DEV)
        f = src.float() / 255.0
    else:
        src = torch.rand(D, H, W, generator=g).to(DEV)
        f = src
    OH, OW, ph, pw = ops.patch_grid(H, W)
    out = torch.full((D * ph * pw, 256), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.preproc_patchify(src, out)
    H16, W16 = (H + 15) // 16 * 16, (W + 15) // 16 * 16
    fp = F.pad(f[:, None], (0, W16 - W, 0, H16 - H), mode="replicate")
    r = F.interpolate(fp, scale_factor=(14 / 16, 14 / 16), mode="bicubic")[:, 0]
    assert r.shape[-2:] == (OH, OW)
    ref = r.view(D, ph, 14, pw, 14).permute(0, 1, 3, 2, 4).reshape(D * ph * pw, 196)
    _close(out[:, :196], ref, atol=6e-3, rtol=8e-3, what="preproc_patchify")
    assert (out[:, 196:] == 0).all()


def test_patchify_3ch(cuda_lib):
    from cryovit_b200 import ops
    B, OH, OW = 2, 56, 70
    x = _rand(B, 3, OH, OW, seed=1)
    out = torch.empty(B * 4 * 5, 640, device=DEV, dtype=torch.bfloat16)
    ops.patchify_f32_3ch(x, out)
    ref = x.view(B, 3, 4, 14, 5, 14).permute(0, 2, 4, 1, 3, 5).reshape(B * 20, 588)
    _close(out[:, :588], ref, atol=1e-2, rtol=1e-2, what="patchify3")
    assert (out[:, 588:] == 0).all()


@pytest.mark.parametrize("C,Np", [(384, 49), (1536, 1024), (384, 784)])
def test_final_norm_writeout(cuda_lib, C, Np):
    from cryovit_b200 import ops
    B, T, D_total, d0 = 2, Np + 5, 5, 2
    x = _rand(B * T, C, scale=2.0, seed=1)
    g, b = _rand(C, seed=2), _rand(C, seed=3)
    feats = torch.zeros(C, D_total, Np, device=DEV, dtype=torch.float16)
    ops.final_norm_writeout(x, g, b, feats, B, T, 5, Np, d0, 1e-6)
    ref = F.layer_norm(x, (C,), g, b, 1e-6).view(B, T, C)[:, 5:]
    ref = ref.permute(2, 0, 1).half()
    _close(feats[:, d0:d0 + B], ref, atol=2e-3, rtol=2e-3, what="final_norm_writeout")
    assert (feats[:, :d0] == 0).all() and (feats[:, d0 + B:] == 0).all()


# ------------------------------------------------------------------------------------------------- head ops
@pytest.mark.parametrize("C,D,h,w", [(100, 3, 5, 7), (384, 4, 8, 8), (1536, 2, 16, 16), (128, 1, 8, 9)])
def test_features_to_ndhwc(cuda_lib, C, D, h, w):
    """(C, DHW) fp16 -> (DHW, C) bf16: the scalar tile kernel (ragged shapes) and the 16-byte vectorised one
    (C and DHW multiples of 64) must both be the exact cast + transpose."""
    from cryovit_b200 import ops
    f = _rand(C, D, h, w, seed=1).half()
    out = torch.empty(D, h, w, C, device=DEV, dtype=torch.bfloat16)
    ops.features_to_ndhwc(f, out)
    assert torch.equal(out, f.permute(1, 2, 3, 0).bfloat16())


@pytest.mark.parametrize("C,G,D,H,W", [(1024, 128, 3, 8, 16), (128, 16, 3, 8, 16), (32, 8, 3, 8, 16), (32, 8, 5, 96, 112),
                                       (1024, 128, 9, 16, 16)])
def test_groupnorm(cuda_lib, C, G, D, H, W):
    from cryovit_b200 import ops
    x = (_rand(D, H, W, C, seed=1) * 1.5 + 0.3).bfloat16()
    g, b = _rand(C, seed=2), _rand(C, seed=3)
    out = torch.empty_like(x)
    stats = torch.empty(2 * G, device=DEV)
    ops.groupnorm_ndhwc(x, out, g, b, stats, G, 1e-3)
    ref = F.group_norm(x.float().permute(3, 0, 1, 2)[None], G, g, b, 1e-3)[0].permute(1, 2, 3, 0)
    _close(out, ref, atol=2e-2, rtol=1e-2, what="groupnorm")


def _conv_w_taps(w, cout_pad):
    # torch Conv3d weight [Cout, Cin, 3, 3, 3] -> [27 * cout_pad, Cin], tap = (kd*3+kh)*3+kw
    Cout, Cin = w.shape[:2]
    wt = torch.zeros(27, cout_pad, Cin, device=w.device, dtype=w.dtype)
    wt[:, :Cout] = w.permute(2, 3, 4, 0, 1).reshape(27, Cout, Cin)
    return wt.reshape(27 * cout_pad, Cin).contiguous()


@pytest.mark.parametrize("D,H,W,Cin,Cout,dil", [
    (8, 8, 16, 64, 64, 2),      # generic, Cin one chunk
    (6, 32, 32, 128, 64, 4),    # SB2a-like, two channel chunks
    (40, 4, 32, 1024, 192, 32), # SB1a geometry class: D > dil
    (8, 4, 32, 1024, 192, 32),  # D <= dil: depth taps are pure padding
    (5, 8, 64, 32, 32, 1),      # SB3-like, 64B swizzle
    (4, 4, 256, 32, 16, 2),     # SB4a: Cout padded 16 -> 32, W > 128
    (4, 2, 128, 16, 16, 1),     # SB4b: 32B swizzle
    (10, 8, 32, 128, 192, 3),   # CTA-pair path (even tile count per plane): two row tiles of a plane share each B tile
    (6, 12, 16, 64, 192, 2),    # CTA pair, the second tile of every plane is partial (rows 8..11 of 16)
    (6, 16, 16, 64, 256, 2),    # CTA pair, 256-wide tile (input-gradient shapes of the training path)
    (4, 16, 16, 64, 512, 1),    # CTA pair, two N tiles
    (3, 32, 32, 1024, 192, 32), # CTA pair at the head's own plane size, D <= dil
])
def test_conv3d_dilated(cuda_lib, D, H, W, Cin, Cout, dil):
    from cryovit_b200 import ops
    x = _rand(D, H, W, Cin, seed=1).bfloat16()
    w = _rand(Cout, Cin, 3, 3, 3, scale=(27 * Cin) ** -0.5, seed=2).bfloat16()
    b = _rand(Cout, seed=3)
    cout_pad = max(32, Cout)
    bp = torch.zeros(cout_pad, device=DEV)
    bp[:Cout] = b
    out = torch.full((D, H, W, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.conv3d_dilated(x, _conv_w_taps(w, cout_pad), bp, out, dil)
    ref = F.conv3d(x.float().permute(3, 0, 1, 2)[None], w.float(), b, padding="same", dilation=(dil, 1, 1))
    ref = F.gelu(ref)[0].permute(1, 2, 3, 0)
    _close(out, ref, atol=2e-2, rtol=1e-2, what="conv3d_dilated")


@pytest.mark.parametrize("D,H,W,Cin,Cout,dil", [
    (5, 32, 16, 32, 32, 1),     # SB3-like, whole tiles
    (9, 20, 13, 32, 32, 8),     # ragged H and W (partial border tiles), D > dil and D-range taps skipped at the ends
    (3, 16, 24, 32, 32, 4),     # D <= dil: both outer depth taps are pure padding
    (4, 17, 260, 32, 16, 2),    # SB4a: Cout 16
    (4, 33, 40, 16, 16, 1),     # SB4b: Cin 16
    (4, 16, 8, 8, 8, 1),        # output_layer.0: Cin 8 (two taps per MMA), Cout 8 of N = 16
    (6, 50, 70, 8, 8, 1),
    (130, 16, 8, 32, 32, 1),    # more tiles than SMs: persistent loop, ring wrap-around
])
def test_conv3d_halo(cuda_lib, D, H, W, Cin, Cout, dil):
    from cryovit_b200 import ops
    from cryovit_b200.head import halo_weight_image
    x = _rand(D, H, W, Cin, seed=1).bfloat16()
    w = _rand(Cout, Cin, 3, 3, 3, scale=(27 * Cin) ** -0.5, seed=2).bfloat16()
    b = _rand(Cout, seed=3)
    cout_pad = 32 if Cout > 16 else 16
    bp = torch.zeros(cout_pad, device=DEV)
    bp[:Cout] = b
    out = torch.full((D, H, W, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.conv3d_halo(x, halo_weight_image(w.cpu(), cout_pad).bfloat16().to(DEV), bp, out, dil, cout_pad)
    ref = F.conv3d(x.float().permute(3, 0, 1, 2)[None], w.float(), b, padding="same", dilation=(dil, 1, 1))
    ref = F.gelu(ref)[0].permute(1, 2, 3, 0)
    _close(out, ref, atol=2e-2, rtol=1e-2, what="conv3d_halo")


@pytest.mark.parametrize("D,H,W,Cin,Cout", [(3, 8, 16, 192, 128), (2, 8, 16, 64, 32), (2, 8, 16, 32, 32), (2, 8, 16, 16, 8),
                                            # tall-tile path (8 / 2 row blocks per stage), incl. ragged last tiles
                                            (2, 64, 64, 16, 8), (3, 50, 70, 16, 8), (1, 50, 70, 32, 32), (2, 40, 52, 64, 32)])
def test_convT(cuda_lib, D, H, W, Cin, Cout):
    from cryovit_b200 import ops
    x = _rand(D, H, W, Cin, seed=1).bfloat16()
    w = _rand(Cin, Cout, 1, 2, 2, scale=Cin ** -0.5, seed=2).bfloat16()
    b = _rand(Cout, seed=3)
    w_sub = w[:, :, 0].permute(2, 3, 1, 0).reshape(4 * Cout, Cin).contiguous()  # row (i*2+j)*Cout + co
    out = torch.full((D, 2 * H, 2 * W, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.convT_1x2x2(x, w_sub, b.repeat(4).contiguous(), out)
    ref = F.conv_transpose3d(x.float().permute(3, 0, 1, 2)[None], w.float(), b, stride=(1, 2, 2))
    ref = F.gelu(ref)[0].permute(1, 2, 3, 0)
    _close(out, ref, atol=2e-2, rtol=1e-2, what="convT")


@pytest.mark.parametrize("D,H,W,Cin,Cout,dil", [(3, 16, 16, 32, 16, 1), (5, 20, 40, 32, 16, 2), (4, 33, 64, 16, 16, 1), (9, 16, 32, 32, 32, 4),
                                                 (2, 50, 36, 16, 16, 1), (7, 8, 6, 32, 16, 8), (3, 48, 256, 32, 16, 2)])
def test_conv3d_wpackn(cuda_lib, D, H, W, Cin, Cout, dil):
    """Narrow-layer convolution with P output voxels of a row per tensor-core row (banded weights, csrc/conv_wpackn.cu)
    against torch on the same bf16 operands: ragged tiles in H and W, dilated depth taps leaving the volume (D <= dil),
    plain bias (64 equal table rows), with and without the GELU."""
    from cryovit_b200 import ops
    from cryovit_b200.head import wpackn_weight_image
    cp = 32 if Cout > 16 else 16
    P = ops.wpackn_group(Cin, cp)
    assert P in (2, 4) and W % P == 0
    x = _rand(D, H, W, Cin, seed=1).bfloat16()
    w = _rand(Cout, Cin, 3, 3, 3, scale=(27 * Cin) ** -0.5, seed=2).bfloat16()
    b = _rand(Cout, seed=3)
    bias = torch.zeros(cp, device=DEV)
    bias[:Cout] = b
    img = wpackn_weight_image(w.float().cpu(), cp, P).to(DEV).bfloat16()
    ref = F.conv3d(x.float().permute(3, 0, 1, 2)[None], w.float(), b, padding="same", dilation=(dil, 1, 1))[0].permute(1, 2, 3, 0)
    for act in (True, False):
        out = torch.full((D, H, W, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.conv3d_wpackn(x, img, bias.repeat(64).contiguous(), out, dil, cp, act=act)
        _close(out, F.gelu(ref) if act else ref, atol=2e-2, rtol=1e-2, what=f"conv3d_wpackn act={act}")


def _gn_partial_sums(partials, rows, n_cols, cpg):
    """[32-row block][n_cols / cpg][2] statistics buffer -> (sum, sum of squares) per column group over the valid rows."""
    pc = n_cols // cpg
    p = partials[: (rows + 31) // 32 * pc * 2].view(-1, pc, 2).double().sum(0)
    return p[:, 0], p[:, 1]


@pytest.mark.parametrize("M,N,K,cpg", [(300, 1024, 384, 8), (4104, 1024, 1536, 8), (136, 1024, 1536, 8), (70, 256, 64, 4)])
def test_groupnorm_statistics_from_the_projection_epilogue(cuda_lib, M, N, K, cpg):
    """layers[0] + GELU that also emits the GroupNorm statistics of what it stores (both A layouts): per group of cpg
    columns, the sums over all rows -- tiles overhanging M, CTA pairs and single CTAs -- against torch."""
    from cryovit_b200 import ops
    at = _rand(K, M, seed=1).half()
    w = _rand(N, K, scale=K ** -0.5, seed=2)
    b = _rand(N, seed=3)
    ref = F.gelu(at.float().t() @ w.half().float().t() + b).double()
    rs, rq = ref.view(M, N // cpg, cpg).sum((0, 2)), (ref * ref).view(M, N // cpg, cpg).sum((0, 2))
    for cfirst in (True, False):
        if cfirst and (M % 8 or N % 128):
            continue
        out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        partials = torch.full((ops.gn_partials_numel(M, N, cpg),), float("nan"), device=DEV)
        if cfirst:
            ops.linear_bias_cfirst_gn(at, w.half(), b, out, partials, cpg)
        else:
            ops.linear_bias_gelu_gn(at.t().contiguous().bfloat16(), w.bfloat16(), b, out, partials, cpg)
        _close(out, ref, atol=3e-2, rtol=2e-2, what="projection")
        s, q = _gn_partial_sums(partials, M, N, cpg)
        tol = 3e-2 if not cfirst else 1e-2  # the bf16 variant rounds A and W to bf16: compare loosely
        assert torch.allclose(s, rs, rtol=tol, atol=tol * M ** 0.5 * cpg), (s - rs).abs().max()
        assert torch.allclose(q, rq, rtol=tol, atol=tol * M ** 0.5), (q - rq).abs().max()


@pytest.mark.parametrize("D,H,W,Cin,Cout,cpg,Cn,dil,halo", [(3, 8, 16, 192, 128, 8, 64, 2, False), (5, 16, 24, 64, 32, 4, 32, 3, True),
                                                        (4, 16, 16, 32, 32, 4, 16, 1, True), (9, 8, 8, 192, 128, 8, 64, 4, False)])
def test_groupnorm_folded_between_convT_and_conv(cuda_lib, D, H, W, Cin, Cout, cpg, Cn, dil, halo):
    """The fused GroupNorm chain of a SynthesisBlock boundary against torch: ConvTranspose3d + GELU (statistics from its
    epilogue) -> GroupNorm(Cout / cpg groups, eps 1e-3, random affine) -> dilated Conv3d(Cout -> Cn) + GELU, where the
    normalisation is folded into the convolution's weights and its 64-row border-aware bias table. Every border voxel
    (all three axes, dilated depth taps) must match: the shift may only arrive through taps inside the volume."""
    from cryovit_b200 import ops
    from cryovit_b200.head import _conv_taps, halo_weight_image
    x = _rand(D, H, W, Cin, seed=1).bfloat16()
    wt = _rand(Cin, Cout, 1, 2, 2, scale=Cin ** -0.5, seed=2)
    bt = _rand(Cout, seed=3)
    gamma, beta = 1.0 + 0.3 * _rand(Cout, seed=4), 0.5 * _rand(Cout, seed=5)
    wc = _rand(Cn, Cout, 3, 3, 3, scale=(27 * Cout) ** -0.5, seed=6)
    bc = _rand(Cn, seed=7)
    G = Cout // cpg
    # producer: transposed convolution + GELU with statistics
    y = torch.empty(D, 2 * H, 2 * W, Cout, device=DEV, dtype=torch.bfloat16)
    partials = torch.full((max(ops.gn_partials_numel(D * H * W, 4 * Cout, cpg), 16 * 256 * 1024),), float("nan"), device=DEV)
    w_sub = wt[:, :, 0].permute(2, 3, 1, 0).reshape(4 * Cout, Cin).bfloat16().contiguous()
    prows, pcols = ops.convT_1x2x2_gn(x, w_sub, bt.repeat(4).contiguous(), y, partials, cpg)
    yr = F.gelu(F.conv_transpose3d(x.float().permute(3, 0, 1, 2)[None], wt.bfloat16().float(), bt, stride=(1, 2, 2)))  # [1,Cout,D,2H,2W]
    _close(y, yr[0].permute(1, 2, 3, 0), atol=3e-2, rtol=2e-2, what="convT")
    # fold + consumer
    cp = (32 if Cn > 16 else 16) if halo else max(32, Cn)
    bias = torch.zeros(cp, device=DEV)
    bias[:Cn] = bc
    if halo:
        w32, layout = halo_weight_image(wc.cpu(), cp).to(DEV), ops.LAYOUT_HALO
    else:
        w32, layout = _conv_taps(wc.cpu(), cp).reshape(-1).to(DEV), ops.LAYOUT_TAPS
    ab = ops.groupnorm_fold_ab(Cout, G, DEV)
    w_fold = torch.empty(w32.numel(), device=DEV, dtype=torch.bfloat16)
    table = torch.empty(64 * cp, device=DEV)
    vox = D * 2 * H * 2 * W
    ops.groupnorm_fold(partials, prows, pcols // cpg, G, vox * cpg, gamma, beta, 1e-3, ab, w32, w_fold, Cout, cp, layout, bias, table)
    # scale / shift against torch's statistics of the STORED tensor
    ys = y.float().permute(3, 0, 1, 2).reshape(G, -1)
    mean, var = ys.mean(1), ys.var(1, unbiased=False)
    a_ref = gamma * (var + 1e-3).rsqrt().repeat_interleave(cpg)
    b_ref = beta - mean.repeat_interleave(cpg) * a_ref
    assert torch.allclose(ab[:Cout], a_ref, rtol=2e-3, atol=1e-4) and torch.allclose(ab[Cout:2 * Cout], b_ref, rtol=2e-3, atol=2e-3)
    out = torch.full((D, 2 * H, 2 * W, Cn), float("nan"), device=DEV, dtype=torch.bfloat16)
    if halo:
        ops.conv3d_halo_tab(y, w_fold, table, out, dil, cp)
    else:
        ops.conv3d_dilated_tab(y, w_fold.view(27 * cp, Cout), table, out, dil)
    ref = F.gelu(F.conv3d(F.group_norm(y.float().permute(3, 0, 1, 2)[None], G, gamma, beta, eps=1e-3), wc, bc, padding="same",
                          dilation=(dil, 1, 1)))[0].permute(1, 2, 3, 0)
    _close(out, ref, atol=4e-2, rtol=3e-2, what=f"GroupNorm folded into conv (halo={halo})")
    if halo and ops.wpackn_group(Cout, cp):  # the same fold into the W-packed kernel's banded weight image
        from cryovit_b200.head import wpackn_weight_image
        w32p = wpackn_weight_image(wc.cpu(), cp, ops.wpackn_group(Cout, cp)).to(DEV)
        w_foldp = torch.empty(w32p.numel(), device=DEV, dtype=torch.bfloat16)
        table2 = torch.empty(64 * cp, device=DEV)
        ops.groupnorm_fold(partials, prows, pcols // cpg, G, vox * cpg, gamma, beta, 1e-3, ab, w32p, w_foldp, Cout, cp,
                           ops.LAYOUT_WPACKN, bias, table2)
        assert torch.allclose(table2, table, rtol=1e-4, atol=1e-5)
        out2 = torch.full((D, 2 * H, 2 * W, Cn), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.conv3d_wpackn(y, w_foldp, table2, out2, dil, cp)
        _close(out2, ref, atol=4e-2, rtol=3e-2, what="GroupNorm folded into the W-packed conv")
    if halo and ops.rows_supported(Cout, Cn):  # ... and into the one-voxel-per-row kernel's rotated chunk images
        from cryovit_b200.head import rowsn_weight_image
        w32r = rowsn_weight_image(wc).to(DEV)
        w_foldr = torch.empty(w32r.numel(), device=DEV, dtype=torch.bfloat16)
        table3 = torch.empty(64 * Cn, device=DEV)
        ops.groupnorm_fold(partials, prows, pcols // cpg, G, vox * cpg, gamma, beta, 1e-3, ab, w32r, w_foldr, Cout, Cn,
                           ops.LAYOUT_ROWS, bc, table3)
        assert torch.allclose(table3.view(64, Cn), table.view(64, cp)[:, :Cn], rtol=1e-4, atol=1e-5)
        out3 = torch.full((D, 2 * H, 2 * W, Cn), float("nan"), device=DEV, dtype=torch.bfloat16)
        ops.conv3d_rows(y, w_foldr, table3, out3, dil)
        _close(out3, ref, atol=4e-2, rtol=3e-2, what="GroupNorm folded into the rows conv")


@pytest.mark.parametrize("D,H,W", [(3, 16, 128), (2, 10, 200)])
def test_head_tail(cuda_lib, D, H, W):
    from cryovit_b200 import ops
    x = _rand(D, H, W, 8, seed=1).bfloat16()
    w1 = _rand(8, 8, 3, 3, 3, scale=0.1, seed=2)
    b1 = _rand(8, seed=3)
    w2 = _rand(1, 8, 3, 3, 3, scale=0.3, seed=4)
    b2 = _rand(1, seed=5)
    w1t = w1.permute(2, 3, 4, 0, 1).reshape(27, 8, 8).contiguous()
    w2t = w2.permute(2, 3, 4, 0, 1).reshape(27, 8).contiguous()
    scratch = torch.empty(D, H, W, 8, device=DEV, dtype=torch.bfloat16)
    logits = torch.empty(D, H, W, device=DEV)
    probs = torch.empty(D, H, W, device=DEV)
    ops.head_tail(x, w1t, b1, w2t, b2, scratch, logits, probs)
    y = F.gelu(F.conv3d(x.float().permute(3, 0, 1, 2)[None], w1, b1, padding="same"))
    y = y.bfloat16().float()  # the 8-channel intermediate is stored in bf16
    z = F.conv3d(y, w2, b2, padding="same").clamp(-5, 5)[0, 0]
    _close(logits, z, atol=5e-2, rtol=1e-2, what="tail logits")  # bf16 intermediate, 216-term sums
    _close(probs, torch.sigmoid(z), atol=5e-3, rtol=1e-2, what="tail probs")


@pytest.mark.parametrize("D,H,W", [(3, 16, 64), (2, 20, 80), (1, 33, 144), (4, 48, 256)])
def test_conv3d_wpack_output_layers(cuda_lib, D, H, W):
    """Both 8-channel output convolutions with the output voxels of a row packed into the MMA N dimension (banded
    weights, csrc/conv_wpack.cu) against fp32 torch on the same bf16 operands: depth taps outside the volume (D = 1, 2),
    ragged tiles in H (20, 33) and in W (80 and 144 are not multiples of the 64 / 128-voxel tile width)."""
    from cryovit_b200 import ops
    from cryovit_b200.head import wpack_weight_image
    x = _rand(D, H, W, 8, seed=1).bfloat16()
    w1 = _rand(8, 8, 3, 3, 3, scale=0.1, seed=2).bfloat16()
    b1 = _rand(8, seed=3)
    w2 = _rand(1, 8, 3, 3, 3, scale=0.3, seed=4).bfloat16()
    b2 = _rand(1, seed=5)
    mid = torch.full((D, H, W, 8), float("nan"), device=DEV, dtype=torch.bfloat16)
    ops.conv3d_wpack8_gelu(x, wpack_weight_image(w1.cpu(), 8).to(DEV).bfloat16(), b1.repeat(8).contiguous(), mid)
    y = F.gelu(F.conv3d(x.float().permute(3, 0, 1, 2)[None], w1.float(), b1, padding="same"))
    _close(mid, y[0].permute(1, 2, 3, 0), atol=1e-2, rtol=1e-2, what="wpack 8->8 + GELU")
    logits = torch.full((D, H, W), float("nan"), device=DEV)
    probs = torch.full((D, H, W), float("nan"), device=DEV)
    ops.conv3d_wpack8_final(mid, wpack_weight_image(w2.cpu(), 16).to(DEV).bfloat16(), b2.repeat(16).contiguous(), logits, probs)
    z = F.conv3d(mid.float().permute(3, 0, 1, 2)[None], w2.float(), b2, padding="same").clamp(-5, 5)[0, 0]
    _close(logits, z, atol=2e-3, rtol=1e-3, what="wpack 8->1 logits")  # same bf16 operands, fp32 accumulation both sides
    _close(probs, torch.sigmoid(z), atol=1e-3, rtol=1e-3, what="wpack probs")
    only = torch.full((D, H, W), float("nan"), device=DEV)
    ops.conv3d_wpack8_final(mid, wpack_weight_image(w2.cpu(), 16).to(DEV).bfloat16(), b2.repeat(16).contiguous(), None, only)
    assert torch.equal(only, probs)


def test_conv3d_wpack_rejects_unpackable_width(cuda_lib):
    from cryovit_b200 import ops
    from cryovit_b200._lib import CryovitB200Error
    from cryovit_b200.head import wpack_weight_image
    x = torch.zeros(1, 16, 60, 8, device=DEV, dtype=torch.bfloat16)  # 60 is not a multiple of 8
    w = wpack_weight_image(torch.zeros(8, 8, 3, 3, 3), 8).to(DEV).bfloat16()
    with pytest.raises(CryovitB200Error, match="multiple of 8"):
        ops.conv3d_wpack8_gelu(x, w, torch.zeros(64, device=DEV), torch.empty_like(x))


@pytest.mark.parametrize("D,H,W", [(3, 8, 128), (5, 20, 136), (20, 17, 264), (35, 9, 8), (2, 3, 520)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_conv3d_rows8(cuda_lib, D, H, W, act):
    """8 -> 8 convolution with one voxel per MMA row (sliding accumulator windows in tensor memory) against F.conv3d:
    ragged row tiles / segments / plane chunks, all three activation modes."""
    from cryovit_b200 import ops
    from cryovit_b200.head import rows8_weight_image
    x = _rand(D, H, W, 8, seed=1).bfloat16()
    w = (_rand(8, 8, 3, 3, 3, seed=2) * (27 * 8) ** -0.5).bfloat16()
    b = _rand(8, seed=3)
    out = torch.full((D, H, W, 8), float("nan"), device=DEV, dtype=torch.bfloat16)
    aux = torch.full_like(out, float("nan")) if act == 2 else None
    ops.conv3d_rows8(x, rows8_weight_image(w).bfloat16(), b, out, act=act, aux=aux)
    z = F.conv3d(x.float().permute(3, 0, 1, 2)[None], w.float(), b, padding=1)[0].permute(1, 2, 3, 0)
    if act == 1:
        _close(out, F.gelu(z), atol=2e-2, rtol=2e-2, what="rows8 gelu")
    else:
        _close(out, z, atol=2e-2, rtol=2e-2, what="rows8 pre-activation")
        if act == 2:
            _close(aux, F.gelu(z), atol=2e-2, rtol=2e-2, what="rows8 aux")


@pytest.mark.parametrize("D,H,W", [(3, 8, 128), (5, 20, 136), (18, 17, 264)])
def test_conv3d_rows8_final(cuda_lib, D, H, W):
    from cryovit_b200 import ops
    from cryovit_b200.head import rows8_weight_image
    x = _rand(D, H, W, 8, seed=1).bfloat16()
    w = (_rand(1, 8, 3, 3, 3, seed=2) * 0.3).bfloat16()
    b = _rand(1, seed=3)
    logits, probs = torch.full((D, H, W), float("nan"), device=DEV), torch.full((D, H, W), float("nan"), device=DEV)
    ops.conv3d_rows8_final(x, rows8_weight_image(w).bfloat16(), b, logits, probs)
    ref = F.conv3d(x.float().permute(3, 0, 1, 2)[None], w.float(), b, padding=1)[0, 0].clamp(-5, 5)
    _close(logits, ref, atol=2e-2, rtol=2e-2, what="rows8 final logits")
    _close(probs, torch.sigmoid(ref), atol=5e-3, rtol=5e-3, what="rows8 final probs")


@pytest.mark.parametrize("D,H,W,Cin,Cout,dil", [(5, 11, 128, 16, 16, 1), (9, 20, 136, 32, 16, 2), (20, 9, 264, 32, 32, 4),
                                                 (7, 6, 40, 16, 32, 3), (40, 5, 130, 16, 16, 1), (3, 4, 12, 32, 32, 8)])
@pytest.mark.parametrize("act", [0, 1, 2])
def test_conv3d_rows(cuda_lib, D, H, W, Cin, Cout, dil, act):
    """16 / 32-channel dilated convolution with one voxel per MMA row against F.conv3d: chunked operands, 144-column windows,
    dilation residue classes, ragged tiles / odd tile counts, a per-mask bias table, the three activation modes."""
    from cryovit_b200 import ops
    from cryovit_b200.head import rowsn_weight_image
    x = _rand(D, H, W, Cin, seed=1).bfloat16()
    w = (_rand(Cout, Cin, 3, 3, 3, seed=2) * (27 * Cin) ** -0.5).bfloat16()
    b = _rand(Cout, seed=3)
    table = b.repeat(64).contiguous()
    out = torch.full((D, H, W, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    aux = torch.full_like(out, float("nan")) if act == 2 else None
    ops.conv3d_rows(x, rowsn_weight_image(w).bfloat16(), table, out, dil, act=act, aux=aux)
    z = F.conv3d(x.float().permute(3, 0, 1, 2)[None], w.float(), b, padding="same", dilation=(dil, 1, 1))[0].permute(1, 2, 3, 0)
    if act == 1:
        _close(out, F.gelu(z), atol=2e-2, rtol=2e-2, what="rows gelu")
    else:
        _close(out, z, atol=2e-2, rtol=2e-2, what="rows pre-activation")
        if act == 2:
            _close(aux, F.gelu(z), atol=2e-2, rtol=2e-2, what="rows aux")


@pytest.mark.parametrize("kind,D,H,W,Cin,Cout,dil", [("rows8", 5, 20, 136, 8, 8, 1), ("rows", 9, 10, 136, 32, 16, 2), ("rows", 6, 9, 40, 16, 32, 3),
                                                     ("rows", 20, 6, 128, 32, 32, 4)])
def test_conv3d_rows_gelu_grad(cuda_lib, kind, D, H, W, Cin, Cout, dil):
    """act 3 of the one-voxel-per-row kernels: out = conv(x) * gelu'(z) with z prefetched before the accumulator wait, and the
    column sums of what is stored (the bias gradient of the layer below) from the same kernel."""
    from cryovit_b200 import ops, train_ops as T
    from cryovit_b200.head import rows8_weight_image, rowsn_weight_image
    x = _rand(D, H, W, Cin, seed=1).bfloat16()
    w = (_rand(Cout, Cin, 3, 3, 3, seed=2) * (27 * Cin) ** -0.5).bfloat16()
    z = (_rand(D, H, W, Cout, seed=4) * 1.5).bfloat16()
    out = torch.full((D, H, W, Cout), float("nan"), device=DEV, dtype=torch.bfloat16)
    db = torch.zeros(Cout, device=DEV)
    zero = torch.zeros(Cout, device=DEV)
    if kind == "rows8":
        ops.conv3d_rows8(x, rows8_weight_image(w).bfloat16(), zero, out, act=3, aux=z, db=db)
    else:
        ops.conv3d_rows(x, rowsn_weight_image(w).bfloat16(), zero.repeat(64).contiguous(), out, dil, act=3, aux=z, db=db)
    y = F.conv3d(x.float().permute(3, 0, 1, 2)[None], w.float(), None, padding="same", dilation=(dil, 1, 1))[0].permute(1, 2, 3, 0)
    ref, dbr = torch.empty_like(out), torch.zeros(Cout, device=DEV)
    T.gelu_bwd(y.bfloat16().contiguous(), z, ref, dbr)
    _close(out, ref, atol=2e-2, rtol=2e-2, what="rows gelu-grad")
    assert (db - dbr).abs().max() <= 2e-2 * dbr.abs().max() + 5e-2, (db, dbr)
