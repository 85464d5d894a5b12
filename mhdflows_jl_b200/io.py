"""Checkpoint / restart of the reference (`savefile`, `Restart!`, `readMHDFlows`: src/integrator.jl:208-288,
src/utils/IC.jl:245-257) over the C ABI.

Same files as the reference: HDF5, `<path>_t_NNNN.h5`, datasets `i_velocity ... k_mag_field` = the real-space `vars`
fields (i.e. the STALE fields) in the element type of the problem and the scalar `time`.  HDF5.jl stores a
Julia `(nx, ny, nz)` array with dataspace dimensions `(nz, ny, nx)`, which is exactly our C-order array.  No HDF5 library
exists in this environment (h5py / HDF5.jl are not installable offline): the files are written and read by the small
format implementation in h5lite.py.  `.npz` dumps of earlier builds are still readable.
Host-side only; nothing here touches the hot path.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import h5lite

_U = (("i_velocity", "ux"), ("j_velocity", "uy"), ("k_velocity", "uz"))
_B = (("i_mag_field", "bx"), ("j_mag_field", "by"), ("k_mag_field", "bz"))


def _filename(file_path_and_name: str, file_number: int) -> str:
    return f"{file_path_and_name}_t_{int(file_number):04d}.h5"        # integrator.jl:260-262


def _dist(prob):
    """torch.distributed of a slab-decomposed problem (one process per GPU): the dump / restart of such a problem is a
    collective over the ranks."""
    try:
        import torch.distributed as dist
        ok = dist.is_available() and dist.is_initialized() and dist.get_world_size() == prob.nranks and dist.get_rank() == prob.rank
    except Exception:
        ok = False
    if not ok:
        raise NotImplementedError("savefile / Restart of a slab-decomposed problem (nranks > 1) gathers / scatters the z slabs through "
                                  "torch.distributed: initialise a process group whose ranks are the problem's ranks")
    return dist


def _gather_slabs(prob, slab):
    """The z slabs (nz / P, ny, nx) of all ranks -> the full (nz, ny, nx) field on rank 0 (None on the other ranks)."""
    import torch
    dist = _dist(prob)
    t = torch.from_numpy(np.ascontiguousarray(slab))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    bufs = [torch.empty_like(t) for _ in range(prob.nranks)] if prob.rank == 0 else None
    dist.gather(t, bufs, dst=0)
    if prob.rank != 0:
        return None
    return np.concatenate([b.cpu().numpy() for b in bufs], axis=0)


def savefile(prob, file_number, file_path_and_name=""):
    """savefile(prob, file_number; file_path_and_name) (integrator.jl:259-288).  Slab-decomposed problems (ours: the reference
    is single-device) write ONE file with the full (nx, ny, nz) datasets: the z slabs are gathered field by field to rank 0,
    which writes; every rank returns the path."""
    data = {}
    T = np.float64
    fields = (() if prob.flag.e else _U) + (_B if prob.flag.b else ())
    for ds, f in fields:
        a = prob.get_real(f, L.STALE)
        T = a.dtype.type
        data[ds] = _gather_slabs(prob, a) if prob.nranks > 1 else a
    path = _filename(file_path_and_name, file_number)
    if prob.rank == 0:
        data["time"] = T(prob.clock.t)                                # write(fw, "time", prob.clock.t): a scalar of type T
        h5lite.write(path, data)
    if prob.nranks > 1:
        _dist(prob).barrier()                                         # the file is complete when any rank returns
    return path


def readMHDFlows(path, as_tuple=False):
    """readMHDFlows(FileName) (utils/IC.jl:245-257).  Default: the datasets of one dump as a dict (works for HD and EMHD
    dumps too); `as_tuple=True` gives the reference's `(iv, jv, kv, ib, jb, kb, t)` with Float32 fields."""
    if str(path).endswith(".npz"):
        with np.load(path) as f:
            d = {k: f[k] for k in f.files}
    else:
        f = h5lite.File(path)
        d = {k: f.read(k) for k in f.names()}
    if not as_tuple:
        return d
    return tuple(d[ds].astype(np.float32) for ds, _ in _U + _B) + (d["time"],)


def Restart(prob, file_path_and_name):
    """Restart!(prob, file_path_and_name) (integrator.jl:208-257): load the fields into vars, r2c them into sol and
    restore clock.t (the step counter and dt are not restored, like the reference)."""
    d = readMHDFlows(file_path_and_name)
    nzl = prob._real_shape[0]

    def mine(a):   # a dump holds the whole grid: a slab-decomposed problem takes its own z planes
        if prob.nranks > 1:
            if a.shape != (prob.grid.nz, prob.grid.ny, prob.grid.nx):
                raise ValueError(f"restart file holds fields of shape {a.shape}, the problem's grid is {(prob.grid.nz, prob.grid.ny, prob.grid.nx)}")
            return a[prob.rank * nzl:(prob.rank + 1) * nzl]
        return a
    if not prob.flag.e:
        for ds, f in _U:
            prob.set_real(f, mine(d[ds]))
    if prob.flag.b:
        for ds, f in _B:
            prob.set_real(f, mine(d[ds]))
    prob.clock.t = float(d["time"])
    return None
