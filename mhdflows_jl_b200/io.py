"""Checkpoint / restart of the reference (`savefile`, `Restart!`, `readMHDFlows`: src/integrator.jl:208-288,
src/utils/IC.jl:245-257) over the C ABI.

Same dataset names and contents as the reference's HDF5 files -- the real-space `vars` fields (i.e. the STALE
fields) and `time` -- but stored as NumPy `.npz` archives: no HDF5 library exists in this environment
(h5py / HDF5.jl are not installable offline), so the container format is the one documented deviation.
Array layout inside the archive is the Julia one transposed to C order, `(nz, ny, nx)`.
Host-side only; nothing here touches the hot path.
"""
from __future__ import annotations

import numpy as np

from . import _lib as L

_U = (("i_velocity", "ux"), ("j_velocity", "uy"), ("k_velocity", "uz"))
_B = (("i_mag_field", "bx"), ("j_mag_field", "by"), ("k_mag_field", "bz"))


def _filename(file_path_and_name: str, file_number: int) -> str:
    return f"{file_path_and_name}_t_{int(file_number):04d}.npz"       # "<path>_t_NNNN" like integrator.jl:260-262


def savefile(prob, file_number, file_path_and_name=""):
    """savefile(prob, file_number; file_path_and_name) (integrator.jl:259-288)."""
    data = {}
    if not prob.flag.e:
        for ds, f in _U:
            data[ds] = prob.get_real(f, L.STALE)
    if prob.flag.b:
        for ds, f in _B:
            data[ds] = prob.get_real(f, L.STALE)
    data["time"] = np.float64(prob.clock.t)
    path = _filename(file_path_and_name, file_number)
    np.savez(path, **data)
    return path


def readMHDFlows(path):
    """readMHDFlows (utils/IC.jl:245-257): the datasets of one dump as a dict."""
    with np.load(path) as f:
        return {k: f[k] for k in f.files}


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
