"""The h5py-free statistics reader (jatts_b200/_h5lite.py) against files from the independent spec-level writer."""
import numpy as np
import pytest

from h5_writer import write_h5
from jatts_b200 import _h5lite


def test_reads_the_recipe_statistics_layout(tmp_path):
    g = np.random.default_rng(0)
    data = {"mel_mean": g.standard_normal(80).astype(np.float32), "mel_scale": (g.random(80) + 0.5).astype(np.float32),
            "mean": g.standard_normal(80).astype(np.float64), "scale": g.random((2, 80)).astype(np.float64),
            "count": np.arange(5, dtype=np.int64)}
    p = tmp_path / "stats.h5"
    write_h5(p, data)
    f = _h5lite.H5LiteFile(p)
    assert sorted(f.keys()) == sorted(data)
    for k, v in data.items():
        got = f[k]
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
    assert np.array_equal(_h5lite.read_hdf5(p, "mean"), data["mean"])
    with pytest.raises(KeyError):
        _h5lite.read_hdf5(p, "nope")


def test_rejects_other_files(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"not hdf5 at all" * 10)
    with pytest.raises(ValueError):
        _h5lite.H5LiteFile(p)
    q = tmp_path / "v2.h5"
    q.write_bytes(b"\x89HDF\r\n\x1a\n" + bytes([2]) + bytes(100))      # superblock version 2 (libver='latest')
    with pytest.raises(NotImplementedError):
        _h5lite.H5LiteFile(q)
