"""The C-ABI library loads and exports every symbol include/mhdflows_b200.h declares; without a GPU it fails
loudly (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "mhdflows_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mhdf_[a-z_A-Z0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    from mhdflows_jl_b200 import build
    return ctypes.CDLL(build.build(verbose=False))


def test_header_declares_the_expected_entry_points():
    syms = _header_symbols()
    for s in ("mhdf_create", "mhdf_destroy", "mhdf_step", "mhdf_calcN", "mhdf_set_real", "mhdf_get_real", "mhdf_set_spectral",
              "mhdf_get_spectral", "mhdf_cfl_dt", "mhdf_energy", "mhdf_helicity", "mhdf_spectrum", "mhdf_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(lib):
    for s in _header_symbols():
        assert hasattr(lib, s), f"libmhdflows_b200.so does not export {s}"


def test_python_binding_lists_the_same_symbols():
    from mhdflows_jl_b200 import _lib
    assert sorted(_lib.SYMBOLS) == _header_symbols()


def test_config_struct_layout_matches_header():
    from mhdflows_jl_b200 import _lib
    names = [f[0] for f in _lib.Config._fields_]
    src = open(os.path.join(ROOT, "include", "mhdflows_b200.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} mhdf_config;", src, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    decl = []
    for stmt in body.split(";"):
        stmt = stmt.strip()
        if not stmt:
            continue
        stmt = re.sub(r"^(const\s+)?(int|double|void\s*\*)\s*", "", stmt)
        decl += [x.strip().lstrip("*").strip() for x in stmt.split(",")]
    assert names == decl
    assert ctypes.sizeof(_lib.Config) == 112


def test_invalid_arguments_are_status_codes_not_crashes(lib):
    from mhdflows_jl_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    assert L.mhdf_create(None, ctypes.byref(h)) == _lib.ERR_INVALID
    cfg = _lib.Config(nx=48, ny=32, nz=32, Lx=1.0, Ly=1.0, Lz=1.0, physics=_lib.MHD, stepper=_lib.RK4, dtype=_lib.F32, nranks=1)
    assert L.mhdf_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.ERR_INVALID
    assert b"powers of two" in L.mhdf_last_error(None)
    assert L.mhdf_destroy(None) == _lib.ERR_INVALID
    assert L.mhdf_step(None, 1) == _lib.ERR_INVALID


def test_no_gpu_means_loud_failure_not_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import mhdflows_jl_b200 as M
    with pytest.raises(M.MHDFlowsError) as ei:
        M.Problem(M.GPU(), nx=32)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mhdflows_jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "oracle" not in txt.replace("no oracle", ""), f"{f} references the oracle"


def test_julia_shim_binds_only_declared_entry_points_with_the_declared_arity():
    """julia/MHDFlowsB200.jl cannot be executed here (no Julia): check statically that every `ccall((:mhdf_x, lib), ...)` names
    an entry point of include/mhdflows_b200.h and passes as many arguments as the prototype declares, and that the
    `MhdfConfig` / `MhdfA99` structs list the fields of `mhdf_config` / `mhdf_a99` in order."""
    import re
    hdr = open(os.path.join(ROOT, "include", "mhdflows_b200.h")).read()
    jl = open(os.path.join(ROOT, "julia", "MHDFlowsB200.jl")).read()
    protos = {}
    for m in re.finditer(r"^(?:int|long long|const char\*)\s+(mhdf_\w+)\(([^;]*)\);", hdr, re.M):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(args.split(","))
    calls = re.findall(r"ccall\(\(:(mhdf_\w+), lib\),\s*\w+,\s*\(([^()]*(?:\{[^()]*\}[^()]*)*)\)", jl)
    assert len(calls) >= 20
    for name, types in calls:
        assert name in protos, f"{name} is not declared in the header"
        n = len([t for t in types.split(",") if t.strip()])
        assert n == protos[name], f"{name}: shim passes {n} arguments, the header declares {protos[name]}"
    bound = {c[0] for c in calls}
    for must in ("mhdf_create", "mhdf_destroy", "mhdf_set_real", "mhdf_get_real", "mhdf_step", "mhdf_cfl_dt", "mhdf_energy",
                 "mhdf_set_random_phase", "mhdf_div_correction", "mhdf_set_forcing_a99", "mhdf_spectrum", "mhdf_helicity"):
        assert must in bound, must

    def c_fields(struct):
        end = hdr.index("} " + struct + ";")
        body = hdr[hdr.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            if decl.strip():
                names += [x.strip().split()[-1].lstrip("*") for x in decl.split(",")]
        return names

    def jl_fields(struct):
        body = re.search(r"struct " + struct + r"\n(.*?)\nend", jl, re.S).group(1)
        return re.findall(r"(\w+)::", body)

    assert jl_fields("MhdfConfig") == c_fields("mhdf_config")
    assert jl_fields("MhdfA99") == c_fields("mhdf_a99")
