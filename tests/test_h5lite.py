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


def _structure(f):
    """Version / flag / size conventions of every structure kind in an HDF5 file, decoded from the raw bytes: what libhdf5
    checks when it opens the file and walks the root group (superblock, root symbol entry, B-tree node, local heap,
    symbol-table node, object headers and their dataspace / datatype / layout messages)."""
    b, base = f.buf, f.base
    sb = base   # superblock starts at the base address
    s = {"sig": bytes(b[sb:sb + 8]), "sb_version": b[sb + 8], "freespace_version": b[sb + 9], "root_group_version": b[sb + 10],
         "reserved0": b[sb + 11], "shared_header_version": b[sb + 12], "sizes": (b[sb + 13], b[sb + 14]), "reserved1": b[sb + 15],
         "k_leaf_k_int": struct.unpack_from("<HH", b, sb + 16), "consistency": struct.unpack_from("<I", b, sb + 20)[0],
         "free_space_addr": struct.unpack_from("<Q", b, sb + 24 + 8)[0], "driver_addr": struct.unpack_from("<Q", b, sb + 24 + 24)[0],
         "eof_is_file_end": base + struct.unpack_from("<Q", b, sb + 24 + 16)[0] == len(b),
         "root_cache_type": f.root["cache"]}
    bt, hp = base + f.root["btree"], base + f.root["heap"]
    ntype, level, used = struct.unpack_from("<BBH", b, bt + 4)
    s["btree"] = {"sig": bytes(b[bt:bt + 4]), "type": ntype, "level": level, "siblings": struct.unpack_from("<QQ", b, bt + 8)}
    s["heap"] = {"sig": bytes(b[hp:hp + 4]), "version": b[hp + 4], "reserved": bytes(b[hp + 5:hp + 8])}
    seg_size, free_off, seg_addr = struct.unpack_from("<QQQ", b, hp + 8)
    s["heap"]["segment_in_file"] = base + seg_addr + seg_size <= len(b)
    s["heap"]["segment_aligned"] = seg_size % 8 == 0
    child = base + struct.unpack_from("<Q", b, bt + 24 + 8)[0]
    s["snod"] = {"sig": bytes(b[child:child + 4]), "version": b[child + 4], "reserved": b[child + 5]}
    links = f._group_links(f.root["btree"], f.root["heap"])
    names = list(links)
    s["names_sorted"] = names == sorted(names)
    objs = []
    for name, e in links.items():
        h = base + e["header"]
        nmsg, refs, size = struct.unpack_from("<HII", b, h + 2)
        o = {"version": b[h], "reserved": b[h + 1], "refcount": refs, "header_aligned": h % 8 == 0, "entry_cache_type": e["cache"]}
        for mtype, data in f._messages(e["header"]):
            if mtype == 0x0001:
                o["dataspace"] = {"version": data[0], "rank": data[1], "flags": data[2]}
            elif mtype == 0x0003:
                o["datatype"] = {"class_version": data[0], "bits": bytes(data[1:4]), "size": struct.unpack_from("<I", data, 4)[0], "props": bytes(data[8:])}
            elif mtype == 0x0008:
                # version 3: (version, class, ...); versions 1 / 2: (version, dimensionality, class, ...)
                o["layout"] = {"version": data[0], "class": data[1] if data[0] >= 3 else data[2]}
                if data[0] >= 3 and data[1] == 1:
                    addr, nbytes = struct.unpack_from("<QQ", data, 2)
                    o["layout"]["data_in_file"] = base + addr + nbytes <= len(b)
        objs.append(o)
    s["objects"] = objs
    return s


def test_written_file_follows_the_conventions_of_a_libhdf5_written_file(tmp_path):
    """No HDF5 library can be installed here, so the writer is validated structurally: every structure kind our file contains
    (superblock, root symbol entry, B-tree node, local heap, symbol-table node, version-1 object headers, dataspace /
    datatype / contiguous-layout messages) must carry the same versions, reserved bytes, flag and size conventions as the
    same structure in a file written by the real library -- the fields libhdf5 validates when HDF5.jl opens a dump."""
    path = os.path.join(DATA, "testhdf5_7.4_GLNX86.mat")
    if not os.path.exists(path):
        pytest.skip("SciPy's MATLAB fixtures are not installed")
    real = _structure(H.File(path))
    rng = np.random.default_rng(1)
    ds = {n: rng.standard_normal((4, 6, 8)) for n in ("i_velocity", "j_velocity", "k_velocity", "i_mag_field", "j_mag_field", "k_mag_field")}
    ds["time"] = np.float64(1.5)
    ours = _structure(H.File(H.write(str(tmp_path / "run_t_0000.h5"), ds)))
    for key in ("sig", "sb_version", "freespace_version", "root_group_version", "reserved0", "shared_header_version", "sizes", "reserved1",
                "k_leaf_k_int", "free_space_addr", "driver_addr", "root_cache_type", "names_sorted"):
        assert ours[key] == real[key], (key, ours[key], real[key])
    # file consistency flags: "unused" in superblock versions 0 / 1 (spec II.A); MATLAB's old library left 3 behind, a cleanly
    # closed file has 0
    assert ours["consistency"] == 0 and ours["eof_is_file_end"]
    for blk in ("btree", "heap", "snod"):
        assert ours[blk] == real[blk], (blk, ours[blk], real[blk])
    ref = real["objects"][0]                              # the genuine Float64 dataset
    assert len(ours["objects"]) == 7
    for o in ours["objects"]:
        for key in ("version", "reserved", "refcount", "header_aligned", "entry_cache_type"):
            assert o[key] == ref[key], (key, o[key], ref[key])
        assert o["datatype"] == ref["datatype"]           # IEEE Float64 LE, byte for byte as libhdf5 encodes it
        # layout message: version 3 (what libhdf5 >= 1.8, i.e. HDF5.jl, writes for contiguous data; MATLAB's older library wrote 2)
        assert o["layout"]["version"] == 3 and ref["layout"]["version"] in (2, 3) and o["layout"]["class"] == ref["layout"]["class"] == 1
        assert o["layout"].get("data_in_file", True)
        assert o["dataspace"]["version"] == ref["dataspace"]["version"]
    ranks = sorted(o["dataspace"]["rank"] for o in ours["objects"])
    assert ranks == [0, 3, 3, 3, 3, 3, 3]                 # `time` is a scalar dataspace, the fields are (nz, ny, nx)
