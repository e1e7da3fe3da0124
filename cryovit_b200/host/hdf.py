"""Tomogram file layout (reference seam B5; run/dino_features.py:109-153, datasets/vit_dataset.py:71-88,
datasets/tomo_dataset.py:89-146):

    data                 (D, H, W) uint8 | float32 in [0, 1]              gzip
    labels/<name>        (D, H, W) int8, -1 = ignore                      gzip
    dino_features        (C, D, H/16, W/16) float16                       uncompressed, contiguous

The container is HDF5 through ``h5py`` whenever that package is importable (production). This image ships no h5py,
so a same-key ``.npz`` container stands in (one zip member per dataset, ``labels/<name>`` keys kept verbatim, gzip
= zip deflate): the file NAME and every key are unchanged, only the byte container differs. Which one is in use is
reported by :func:`backend`.
"""
from __future__ import annotations

import io
import os
import zipfile
from pathlib import Path

import numpy as np

try:  # pragma: no cover - depends on the deployment image
    import h5py  # type: ignore

    _H5 = True
except Exception:  # noqa: BLE001
    h5py = None
    _H5 = False


def backend() -> str:
    return "h5py" if _H5 else "npz"


def read_tomogram(path: Path | str, keys: list[str] | None = None) -> dict[str, np.ndarray]:
    """All datasets of a tomogram file as {key: array}; group members come back as ``group/member``."""
    path = Path(path)
    out: dict[str, np.ndarray] = {}
    if _H5:
        with h5py.File(path, "r") as fh:
            def visit(name, obj):
                if isinstance(obj, h5py.Dataset) and (keys is None or name in keys):
                    out[name] = obj[()]
            fh.visititems(visit)
        return out
    with zipfile.ZipFile(path, "r") as zf:
        for member in zf.namelist():
            name = member[:-4] if member.endswith(".npy") else member
            if keys is None or name in keys:
                out[name] = np.load(io.BytesIO(zf.read(member)), allow_pickle=False)
    return out


def list_keys(path: Path | str) -> list[str]:
    path = Path(path)
    if _H5:
        names: list[str] = []
        with h5py.File(path, "r") as fh:
            fh.visititems(lambda n, o: names.append(n) if isinstance(o, h5py.Dataset) else None)
        return names
    with zipfile.ZipFile(path, "r") as zf:
        return [m[:-4] if m.endswith(".npy") else m for m in zf.namelist()]


def write_tomogram(path: Path | str, datasets: dict[str, np.ndarray], uncompressed: tuple[str, ...] = ("dino_features",)) -> None:
    """Write (overwrite: the reference opens with "w", dino_features.py:119) a tomogram file. Every dataset is
    gzip-compressed except those named in ``uncompressed`` (the reference stores dino_features raw, :148-153)."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    # written next to the target and renamed over it: an interrupted run never leaves a truncated result file
    # behind (``skip_existing`` would have to trust it), and a reader never sees a half-written one
    tmp = path.with_name(path.name + f".tmp{os.getpid()}")
    try:
        if _H5:
            with h5py.File(tmp, "w") as fh:
                for key, arr in datasets.items():
                    kw = {} if key in uncompressed else {"compression": "gzip"}  # h5py default: level 4
                    fh.create_dataset(key, data=arr, shape=arr.shape, dtype=arr.dtype, **kw)
        else:
            with zipfile.ZipFile(tmp, "w", compresslevel=4) as zf:  # h5py's gzip default is level 4 too
                for key, arr in datasets.items():
                    info = zipfile.ZipInfo(key + ".npy")
                    info.compress_type = zipfile.ZIP_STORED if key in uncompressed else zipfile.ZIP_DEFLATED
                    with zf.open(info, "w", force_zip64=True) as member:  # streamed: no second copy of 403 MB
                        np.save(member, np.ascontiguousarray(arr), allow_pickle=False)
        os.replace(tmp, path)
    finally:
        if tmp.exists():
            tmp.unlink()
