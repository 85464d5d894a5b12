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


def savefile(prob, file_number, file_path_and_name=""):
    """savefile(prob, file_number; file_path_and_name) (integrator.jl:259-288)."""
    data = {}
    if not prob.flag.e:
        for ds, f in _U:
            data[ds] = prob.get_real(f, L.STALE)
    if prob.flag.b:
        for ds, f in _B:
            data[ds] = prob.get_real(f, L.STALE)
    T = next((a.dtype.type for a in data.values()), np.float64)
    data["time"] = T(prob.clock.t)                                    # write(fw, "time", prob.clock.t): a scalar of type T
    path = _filename(file_path_and_name, file_number)
    h5lite.write(path, data)
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
    if not prob.flag.e:
        for ds, f in _U:
            prob.set_real(f, d[ds])
    if prob.flag.b:
        for ds, f in _B:
            prob.set_real(f, d[ds])
    prob.clock.t = float(d["time"])
    return None
