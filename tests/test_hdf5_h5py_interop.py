"""Cross-check of host/hdf5_classic.py against libhdf5 itself. Runs wherever ``h5py`` is importable (it is in neither
the build image nor the GPU box of this project, where these tests skip): files this writer produces must open in
h5py with the reference's storage properties, and files h5py writes the way the reference does
(run/dino_features.py:119-153: default ``libver``, ``compression="gzip"`` for data / labels, raw ``dino_features``)
must read back through the classic reader."""
import numpy as np
import pytest

h5py = pytest.importorskip("h5py")

from cryovit_b200.host import hdf5_classic as h5c  # noqa: E402


def _sets():
    rng = np.random.default_rng(0)
    return {
        "data": rng.integers(0, 255, (20, 70, 50), dtype=np.uint8),
        "labels/mito": rng.integers(-1, 2, (20, 70, 50)).astype(np.int8),
        "labels/granule": rng.integers(-1, 2, (20, 70, 50)).astype(np.int8),
        "dino_features": rng.standard_normal((16, 20, 5, 4)).astype(np.float16),
    }


def test_libhdf5_opens_what_the_classic_writer_wrote(tmp_path):
    sets = _sets()
    path = tmp_path / "w.hdf"
    h5c.write_file(path, sets, gzip={k: 4 for k in sets if k != "dino_features"}, chunks={"data": (3, 32, 32)})
    with h5py.File(path, "r") as fh:
        assert sorted(fh.keys()) == ["data", "dino_features", "labels"] and sorted(fh["labels"].keys()) == ["granule", "mito"]
        for k, v in sets.items():
            assert fh[k].dtype == v.dtype and fh[k].shape == v.shape and np.array_equal(fh[k][()], v), k
        assert fh["data"].compression == "gzip" and fh["data"].compression_opts == 4 and fh["data"].chunks == (3, 32, 32)
        assert fh["dino_features"].compression is None and fh["dino_features"].chunks is None


def test_classic_reader_reads_what_h5py_wrote_like_the_reference(tmp_path):
    sets = _sets()
    path = tmp_path / "r.hdf"
    with h5py.File(path, "w") as fh:  # run/dino_features.py:119-153
        for k, v in sets.items():
            if k == "dino_features":
                fh.create_dataset(k, data=v)
            else:
                fh.create_dataset(k, data=v, shape=v.shape, dtype=v.dtype, compression="gzip")
    got = h5c.read_file(path)
    assert sorted(got) == sorted(sets)
    for k, v in sets.items():
        assert got[k].dtype == v.dtype and np.array_equal(got[k], v), k
