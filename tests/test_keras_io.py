"""CPU tests of the HDF5 / Keras-v3 archive reader (recsys_b200.keras_io) that lets the service load the reference's
`encoder.keras` (src/trainer.py:230) without h5py.  The fixtures are produced by the module's own writer (h5py is not
available here, so conformance with files written by h5py / Keras themselves is NOT pinned -- see the module docstring); a
few hand-assembled variants cover structures the writer does not emit (compact layout, version-2 dataspace), and ONE file
written by the genuine HDF5 library is read: tests/golden/libhdf5_written_matlab73.mat, a MATLAB 7.4 "-v7.3" file (HDF5
behind a 512-byte user block, superblock v0, symbol-table groups -- the structures h5py's defaults also produce) copied
from scipy's test data (scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, BSD-3-Clause)."""

import struct
import zipfile
from pathlib import Path

import numpy as np
import pytest

import recsys_b200 as rs
from recsys_b200 import keras_io


def test_hdf5_roundtrip_nested_groups_and_dtypes(tmp_path):
    rng = np.random.default_rng(0)
    ds = {
        "a/vars/0": rng.standard_normal((7, 5)).astype(np.float32),
        "a/vars/1": rng.standard_normal(5).astype(np.float32),
        "a/b/c/deep": rng.standard_normal((2, 3, 4)),
        "ints/i64": rng.integers(-9, 9, (11,)),
        "ints/u8": rng.integers(0, 255, (3, 3)).astype(np.uint8),
        "ints/i16": rng.integers(-300, 300, (4,)).astype(np.int16),
        "half": rng.standard_normal(6).astype(np.float16),
        "scalar": np.float32(3.5),
        "empty": np.zeros((0, 4), np.float32),
    }
    for i in range(21):                                   # more links than one symbol-table node holds (8)
        ds[f"many/d{i:02d}"] = np.full((2,), i, np.int32)
    path = tmp_path / "x.h5"
    keras_io.write_hdf5(path, ds)
    got = keras_io.read_hdf5(path)
    assert set(got) == set(ds)
    for k, v in ds.items():
        assert got[k].dtype == np.asarray(v).dtype and got[k].shape == np.asarray(v).shape, k
        np.testing.assert_array_equal(got[k], v)
    assert keras_io.read_hdf5(path.read_bytes()).keys() == got.keys()          # bytes in, same result
    with pytest.raises(ValueError):
        keras_io.read_hdf5(b"not an hdf5 file at all" * 100)


def test_reader_accepts_compact_layout_and_v2_dataspace(tmp_path):
    """Patch a written file: turn the contiguous layout message of a small dataset into a compact one and its
    dataspace into version 2 -- both legal encodings the writer never produces."""
    path = tmp_path / "y.h5"
    arr = np.arange(4, dtype=np.float32)
    keras_io.write_hdf5(path, {"v": arr})
    blob = bytearray(path.read_bytes())
    i = blob.find(struct.pack("<HH", 0x08, 24) + b"\0\0\0\0" + b"\x03\x01")    # the layout message (type 8, size 24)
    assert i > 0
    body = i + 8
    blob[body:body + 24] = (struct.pack("<BBH", 3, 0, 16) + arr.tobytes()).ljust(24, b"\0")
    j = blob.find(struct.pack("<HH", 0x01, 16) + b"\0\0\0\0" + b"\x01\x01")    # dataspace v1, rank 1
    assert j > 0
    blob[j + 8:j + 8 + 16] = (struct.pack("<BBBB", 2, 1, 0, 1) + struct.pack("<Q", 4)).ljust(16, b"\0")
    got = keras_io.read_hdf5(bytes(blob))
    np.testing.assert_array_equal(got["v"], arr)


def test_keras_encoder_archive_roundtrip(tmp_path):
    cfg = rs.ModelConfig()
    w = rs.synth.make_weights(cfg, 60, 40, seed=3)
    path = tmp_path / "encoder.keras"
    keras_io.save_keras_encoder(path, w)
    with zipfile.ZipFile(path) as z:
        assert "model.weights.h5" in z.namelist()
        names = keras_io.read_hdf5(z.read("model.weights.h5")).keys()
    assert "user_tower/layers/dense_3/vars/1" in names and "item_embedding/vars/0" in names
    got = keras_io.load_keras_encoder(path)
    for side in ("user", "item"):
        np.testing.assert_array_equal(got[f"{side}_embedding"], w[f"{side}_embedding"])
        assert len(got[f"{side}_tower"]) == len(w[f"{side}_tower"]) == 4
        for (k1, b1), (k2, b2) in zip(got[f"{side}_tower"], w[f"{side}_tower"]):
            np.testing.assert_array_equal(k1, k2)
            np.testing.assert_array_equal(b1, b2)


def test_encoder_mapping_tolerates_prefixes_and_orders_layers_numerically():
    rng = np.random.default_rng(1)
    ds = {}
    dims = [8] + [5] * 12
    for side in ("user", "item"):
        ds[f"encoder/{side}_embedding/vars/0"] = rng.standard_normal((9, 8)).astype(np.float32)
        for i in range(12):                               # dense, dense_1, ..., dense_11: "dense_10" must come after "dense_9"
            name = "dense" if i == 0 else f"dense_{i}"
            ds[f"encoder/{side}_tower/layers/{name}/vars/0"] = np.full((dims[i], dims[i + 1]), i, np.float32)
            ds[f"encoder/{side}_tower/layers/{name}/vars/1"] = np.full((dims[i + 1],), i, np.float32)
    w = keras_io.encoder_weights_from_datasets(ds)
    assert [int(k[0, 0]) for k, _ in w["user_tower"]] == list(range(12))
    del ds["encoder/item_tower/layers/dense_4/vars/1"]
    with pytest.raises(KeyError):
        keras_io.encoder_weights_from_datasets(ds)


@pytest.mark.parametrize("deflate", [False, True])
def test_reader_handles_chunked_and_compressed_datasets(tmp_path, deflate):
    """Chunked storage behind a B-tree v1 index, optionally shuffle + deflate, with ragged edge chunks."""
    rng = np.random.default_rng(4)
    ds = {"w/kernel": rng.standard_normal((37, 21)).astype(np.float32), "w/ids": rng.integers(0, 1 << 40, (50,)),
          "w/plain": rng.standard_normal(9).astype(np.float32)}
    path = tmp_path / "c.h5"
    keras_io.write_hdf5(path, ds, chunks={"w/kernel": (16, 8), "w/ids": (16,)}, deflate=deflate)
    got = keras_io.read_hdf5(path)
    for k, v in ds.items():
        np.testing.assert_array_equal(got[k], v)


def test_hdf5_roundtrip_random_trees():
    """Property test: any tree of groups / datasets the writer accepts comes back identical through the reader."""
    from hypothesis import given, settings, strategies as st

    names = st.text(alphabet="abcdefghijklmnopqrstuvwxyz_0123456789", min_size=1, max_size=12)
    dtypes = st.sampled_from([np.float32, np.float64, np.float16, np.int64, np.int32, np.uint8])
    shapes = st.lists(st.integers(0, 9), min_size=0, max_size=3).map(tuple)

    @st.composite
    def datasets(draw):
        out = {}
        for _ in range(draw(st.integers(1, 12))):
            path = "/".join(draw(st.lists(names, min_size=1, max_size=4)))
            if any(path == p or path.startswith(p + "/") or p.startswith(path + "/") for p in out):
                continue                                   # a name cannot be both a group and a dataset
            dt, shp = draw(dtypes), draw(shapes)
            seed = draw(st.integers(0, 2 ** 31))
            a = np.random.default_rng(seed).integers(0, 200, size=shp)
            out[path] = a.astype(dt)
        return out

    @settings(max_examples=40, deadline=None)
    @given(datasets())
    def check(ds):
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            path = Path(tmp) / "t.h5"
            keras_io.write_hdf5(path, ds)
            got = keras_io.read_hdf5(path)
        assert set(got) == set(ds)
        for k, v in ds.items():
            assert got[k].dtype == v.dtype and got[k].shape == v.shape
            np.testing.assert_array_equal(got[k], v)

    check()


def test_reader_parses_a_file_written_by_the_real_hdf5_library():
    """The one independent check of the reader: a file libhdf5 itself wrote (through MATLAB 7.4).  Its single variable is
    scipy's `testdouble` case, the row 0 : pi/4 : 2*pi stored as a [9, 1] float64 dataset."""
    path = Path(__file__).resolve().parent / "golden" / "libhdf5_written_matlab73.mat"
    data = path.read_bytes()
    assert data[:19] == b"MATLAB 7.0 MAT-file" and data[512:520] == b"\x89HDF\r\n\x1a\n"     # user block, then the superblock
    ds = keras_io.read_hdf5(data)
    assert list(ds) == ["testdouble"]
    x = ds["testdouble"]
    assert x.dtype == np.float64 and x.shape == (9, 1)
    np.testing.assert_allclose(x[:, 0], np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
