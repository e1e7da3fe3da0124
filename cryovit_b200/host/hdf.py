"""Tomogram file layout (reference seam B5; run/dino_features.py:109-153, datasets/vit_dataset.py:71-88,
datasets/tomo_dataset.py:89-146):

    data                 (D, H, W) uint8 | float32 in [0, 1]              gzip
    labels/<name>        (D, H, W) int8, -1 = ignore                      gzip
    dino_features        (C, D, H/16, W/16) float16                       uncompressed, contiguous

The files are HDF5, as the reference's are. Three byte-level back ends, reported by :func:`backend`:

* ``"h5py"``          -- whenever that package is importable (production);
* ``"hdf5-classic"``  -- the self-contained writer / reader of the classic HDF5 format in :mod:`.hdf5_classic` (what
                         h5py's default ``libver="earliest"`` emits). This image ships no h5py, so this is what runs
                         here and on the GPU box: the files are real HDF5 (signature, superblock, symbol-table groups,
                         gzip'd chunk B-trees, contiguous ``dino_features``), not a stand-in container;
* ``"npz"``           -- the same-key zip container of the first round, only with ``CRYOVIT_HDF_BACKEND=npz``.

Reading sniffs the file: HDF5 signature -> h5py / hdf5-classic, ``PK`` -> npz, so files of either kind stay readable.
"""
from __future__ import annotations

import io
import os
import zipfile
from pathlib import Path

import numpy as np

from . import hdf5_classic

try:  # pragma: no cover - depends on the deployment image
    import h5py  # type: ignore

    _H5 = True
except Exception:  # noqa: BLE001
    h5py = None
    _H5 = False

GZIP_LEVEL = 4  # h5py's ``compression="gzip"`` default (``compression_opts`` unset)


def backend() -> str:
    want = os.environ.get("CRYOVIT_HDF_BACKEND", "").strip().lower()
    if want in ("npz", "hdf5-classic"):
        return want
    if want == "h5py" and not _H5:
        raise ImportError("CRYOVIT_HDF_BACKEND=h5py, but h5py is not importable")
    return "h5py" if _H5 else "hdf5-classic"


def _kind(path: Path) -> str:
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic[:2] == b"PK":
        return "npz"
    if magic == hdf5_classic.SIGNATURE or hdf5_classic.is_hdf5(path):
        return "h5py" if _H5 and backend() == "h5py" else "hdf5-classic"
    raise ValueError(f"{path}: neither an HDF5 file nor the zip container")


def read_tomogram(path: Path | str, keys: list[str] | None = None) -> dict[str, np.ndarray]:
    """All datasets of a tomogram file as {key: array}; group members come back as ``group/member``."""
    path = Path(path)
    out: dict[str, np.ndarray] = {}
    kind = _kind(path)
    if kind == "h5py":  # pragma: no cover - needs h5py
        with h5py.File(path, "r") as fh:
            def visit(name, obj):
                if isinstance(obj, h5py.Dataset) and (keys is None or name in keys):
                    out[name] = obj[()]
            fh.visititems(visit)
        return out
    if kind == "hdf5-classic":
        return hdf5_classic.read_file(path, keys)
    with zipfile.ZipFile(path, "r") as zf:
        for member in zf.namelist():
            name = member[:-4] if member.endswith(".npy") else member
            if keys is None or name in keys:
                out[name] = np.load(io.BytesIO(zf.read(member)), allow_pickle=False)
    return out


def list_keys(path: Path | str) -> list[str]:
    path = Path(path)
    kind = _kind(path)
    if kind == "h5py":  # pragma: no cover - needs h5py
        names: list[str] = []
        with h5py.File(path, "r") as fh:
            fh.visititems(lambda n, o: names.append(n) if isinstance(o, h5py.Dataset) else None)
        return names
    if kind == "hdf5-classic":
        return hdf5_classic.list_keys(path)
    with zipfile.ZipFile(path, "r") as zf:
        return [m[:-4] if m.endswith(".npy") else m for m in zf.namelist()]


def write_tomogram(path: Path | str, datasets: dict[str, np.ndarray], uncompressed: tuple[str, ...] = ("dino_features",),
                   chunks: dict[str, tuple[int, ...]] | None = None) -> None:
    """Write (overwrite: the reference opens with "w", dino_features.py:119) a tomogram file. Every dataset is
    gzip-compressed except those named in ``uncompressed`` (the reference stores dino_features raw and contiguous,
    :148-153); ``chunks[key]`` stores an uncompressed dataset in chunks instead (HDF5 back ends only)."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    # written next to the target and renamed over it: an interrupted run never leaves a truncated result file
    # behind (``skip_existing`` would have to trust it), and a reader never sees a half-written one
    tmp = path.with_name(path.name + f".tmp{os.getpid()}")
    kind = backend()
    try:
        if kind == "h5py":  # pragma: no cover - needs h5py
            with h5py.File(tmp, "w") as fh:
                for key, arr in datasets.items():
                    kw = {} if key in uncompressed else {"compression": "gzip"}
                    if chunks and key in chunks:
                        kw["chunks"] = tuple(min(c, n) for c, n in zip(chunks[key], arr.shape))
                    fh.create_dataset(key, data=arr, shape=arr.shape, dtype=arr.dtype, **kw)
        elif kind == "hdf5-classic":
            hdf5_classic.write_file(tmp, datasets, gzip={k: GZIP_LEVEL for k, a in datasets.items()
                                                         if k not in uncompressed and np.ndim(a) > 0},
                                    chunks={k: c for k, c in (chunks or {}).items() if k in uncompressed})
        else:
            with zipfile.ZipFile(tmp, "w", compresslevel=GZIP_LEVEL) as zf:
                for key, arr in datasets.items():
                    info = zipfile.ZipInfo(key + ".npy")
                    info.compress_type = zipfile.ZIP_STORED if key in uncompressed else zipfile.ZIP_DEFLATED
                    with zf.open(info, "w", force_zip64=True) as member:  # streamed: no second copy of 403 MB
                        np.save(member, np.ascontiguousarray(arr), allow_pickle=False)
        os.replace(tmp, path)
    finally:
        if tmp.exists():
            tmp.unlink()
