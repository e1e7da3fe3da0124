"""CPU suite: the C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol that
include/cryovit_b200.h declares (no compute calls here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "cryovit_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cvit_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    names = _declared()
    assert len(names) >= 18
    for n in ("cvit_linear_bias_bf16", "cvit_attention_fwd_bf16", "cvit_final_norm_writeout_f16",
              "cvit_conv3d_dilated_ndhwc", "cvit_last_error"):
        assert n in names


def test_library_builds_and_exports_every_declared_symbol():
    from cryovit_b200 import _lib, build

    so = build.build()
    lib = ctypes.CDLL(str(so))
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in include/cryovit_b200.h but not exported"
    bound = _lib.load()
    assert bound.cvit_abi_version() == 1
    assert set(_lib.SIGNATURES) | {"cvit_last_error", "cvit_abi_version", "cvit_conv3d_halo_weight_bytes", "cvit_conv3d_wpack_weight_bytes",
                                   "cvit_groupnorm_fold_ab_elems", "cvit_conv3d_wpackn_group", "cvit_conv3d_wpackn_weight_bytes",
                                   "cvit_conv3d_rows8_weight_bytes", "cvit_conv3d_rows_weight_bytes", "cvit_convT_gn_partial_rows"} == set(_declared())
    assert bound.cvit_conv3d_wpack_weight_bytes(8, 8) == 9 * 5 * 2 * 64 * 16 and bound.cvit_conv3d_wpack_weight_bytes(16, 1) == 9 * 9 * 2 * 16 * 16
    assert bound.cvit_conv3d_halo_weight_bytes(32, 32) == 9 * 6 * 2 * 32 * 16 and bound.cvit_conv3d_halo_weight_bytes(8, 16) == 9 * 2 * 2 * 16 * 16


def test_product_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cryovit_b200._lib import CryovitB200Error
    from cryovit_b200.vit import build_model

    m = build_model("dinov2_vits14_reg")
    with pytest.raises(CryovitB200Error):
        m.cuda()
    with pytest.raises(CryovitB200Error):
        m.forward_features(torch.zeros(1, 3, 28, 28))


def test_nvtx_spans_are_free_when_disabled(monkeypatch):
    """CRYOVIT_B200_NVTX unset: every span is the same no-op context manager (nothing on the measured path)."""
    import importlib

    monkeypatch.delenv("CRYOVIT_B200_NVTX", raising=False)
    from cryovit_b200 import nvtx
    nvtx = importlib.reload(nvtx)
    assert not nvtx.ENABLED and nvtx.span("a") is nvtx.span("b")
    with nvtx.span("x"):
        pass
