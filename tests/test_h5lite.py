"""mhdflows_jl_b200/h5lite.py: the reader against a file written by the real HDF5 library (the MATLAB-7.3 fixture that
ships with SciPy, checked against its MATLAB-5 twin read by scipy.io.loadmat), the writer through the reader plus the
byte-level structure libhdf5 expects (superblock v0, TREE / HEAP / SNOD blocks, version-1 object headers)."""
import os
import struct

import numpy as np
import pytest
import scipy.io

from mhdflows_jl_b200 import h5lite as H

DATA = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data")


def test_reader_against_a_file_written_by_libhdf5():
    path = os.path.join(DATA, "testhdf5_7.4_GLNX86.mat")
    twin = os.path.join(DATA, "testdouble_7.4_GLNX86.mat")
    if not (os.path.exists(path) and os.path.exists(twin)):
        pytest.skip("SciPy's MATLAB fixtures are not installed")
    f = H.File(path)                                     # 512-byte user block, superblock version 0
    assert f.base == 512 and f.names() == ["testdouble"] and not f.is_group("testdouble")
    v = f.read("testdouble")
    ref = scipy.io.loadmat(twin)["testdouble"]
    assert v.dtype == np.float64 and v.shape == (9, 1)
    assert np.array_equal(v.ravel(), ref.ravel())
    # the heap free list of the genuine file ends with libhdf5's marker 1 -- the value the writer emits
    heap = f.base + f.root["heap"]
    seg_size, free_off, seg_addr = struct.unpack_from("<QQQ", f.buf, heap + 8)
    nxt, size = struct.unpack_from("<QQ", f.buf, f.base + seg_addr + free_off)
    assert nxt == H.HEAP_FREE_NULL and free_off + size == seg_size


def test_write_read_round_trip_and_structure(tmp_path):
    rng = np.random.default_rng(0)
    ds = {"i_velocity": rng.standard_normal((6, 5, 4)).astype(np.float32), "j_velocity": rng.standard_normal((6, 5, 4)).astype(np.float32),
          "k_velocity": rng.standard_normal((6, 5, 4)).astype(np.float32), "i_mag_field": rng.standard_normal((6, 5, 4)),
          "j_mag_field": rng.standard_normal((6, 5, 4)), "k_mag_field": rng.standard_normal((6, 5, 4)),
          "gas_density": rng.integers(0, 9, (3, 2)).astype(np.int32), "time": np.float32(0.75)}
    path = H.write(str(tmp_path / "dump_t_0001.h5"), ds)
    f = H.File(path)
    assert f.names() == sorted(ds)
    for k, a in ds.items():
        got = f.read(k)
        assert got.dtype == np.asarray(a).dtype and np.array_equal(got, a)
    assert f.read("time").shape == ()                    # scalar dataspace, like write(fw, "time", prob.clock.t)
    b = f.buf
    assert b[:8] == H.SIG and b[8] == 0 and (b[13], b[14]) == (8, 8)
    assert struct.unpack_from("<Q", b, 40)[0] == len(b)  # end-of-file address
    assert b[f.root["btree"]:f.root["btree"] + 4] == b"TREE" and b[f.root["heap"]:f.root["heap"] + 4] == b"HEAP"
    # B-tree: one child, keys = ("" , largest name); symbol-table node: entries sorted by name
    bt = f.root["btree"]
    assert struct.unpack_from("<BBH", b, bt + 4) == (0, 0, 1)
    k0, child, k1 = struct.unpack_from("<QQQ", b, bt + 24)
    seg = struct.unpack_from("<Q", b, f.root["heap"] + 24)[0]
    assert k0 == 0 and b[child:child + 4] == b"SNOD" and b[seg + k1:seg + k1 + 5] == b"time\0"
    assert struct.unpack_from("<H", b, child + 6)[0] == len(ds)
    offs = [struct.unpack_from("<Q", b, child + 8 + 40 * i)[0] for i in range(len(ds))]
    names = [b[seg + o:b.find(b"\0", seg + o)].decode() for o in offs]
    assert names == sorted(ds) and all(o % 8 == 0 for o in offs)
    # every object header: version 1, message sizes multiples of 8, data 8-byte aligned inside the file
    for k in ds:
        h = f._resolve(k)["header"]
        assert h % 8 == 0 and b[h] == 1
        nmsg, refs, size = struct.unpack_from("<HII", b, h + 2)
        assert nmsg == 4 and refs == 1 and size % 8 == 0
        types = [t for t, _ in f._messages(h)]
        assert types == [0x0001, 0x0003, 0x0005, 0x0008]
    # the Float64 datatype message equals the one libhdf5 wrote into the genuine fixture
    dt = dict(f._messages(f._resolve("i_mag_field")["header"]))[0x0003]
    assert dt.hex() == "11203f000800000000004000340b0034ff03000000000000"


def test_writer_refuses_what_it_cannot_represent(tmp_path):
    with pytest.raises(H.H5Error):
        H.write(str(tmp_path / "a.h5"), {})
    with pytest.raises(H.H5Error):
        H.write(str(tmp_path / "a.h5"), {f"d{i}": np.zeros(2) for i in range(9)})
    with pytest.raises(H.H5Error):
        H.write(str(tmp_path / "a.h5"), {"c": np.zeros(2, np.complex64)})
    with pytest.raises(H.H5Error):
        H.write(str(tmp_path / "a.h5"), {"a/b": np.zeros(2)})
    p = tmp_path / "junk.h5"
    p.write_bytes(b"not hdf5" * 100)
    with pytest.raises(H.H5Error):
        H.File(str(p))
