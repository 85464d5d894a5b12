"""In-tree nvcc build of the C-ABI library (sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmhdflows_b200.so")
# one translation unit per precision (the Solver<T> template with all its kernels) + the extern "C" boundary: they compile
# in parallel and are linked into one shared object
SOURCES = [os.path.join(CSRC, f) for f in ("api.cu", "solver_f32.cu", "solver_f64.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "solver.cuh"), os.path.join(CSRC, "kernels.cuh"), os.path.join(CSRC, "fft_core.cuh"),
                  os.path.join(ROOT, "include", "mhdflows_b200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-I/usr/include", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
LINK_FLAGS = ["-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-ldl"]


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found; cannot build libmhdflows_b200.so")
    return p


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


# Tuning variants: built next to the default library as libmhdflows_b200_<name>.so and selected at run time with MHDF_LIB
# (tools/ab.sh runs the same-box A/B).  f32x2 = Float32 butterflies / products on the packed sm_100 instructions
# (add/mul/fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2), see csrc/fft_core.cuh.
# Measured in round 2 (profiles/r02_c1_ab_f32x2.log): strided passes -3...5 %, x pass -3 % at 256^3 but +4 % at 512^3 -- stays opt-in.
VARIANTS = {"f32x2": ["-DMHDF_F32X2"],
            # strided passes on the scalar FP32 forms (the default uses the packed float2p arithmetic there): A/B partner
            "pass_scalar": ["-DMHDF_PASS_SCALAR"],
            # 16 columns per block in the 1024-point y passes as well (the z passes have them by default): measured 2 % slower per
            # 1024^3 step with 16 points per thread (profiles/r02_c13_time1024.log) and with 32 (r02_c17_time1024.log)
            "ytx16": ["-DMHDF_Y_TX16"],
            "e16": ["-DMHDF_PASS_E32=0"],            # 16 points per thread in every strided pass (the default until call 16)
            "fwdp": ["-DMHDF_PASS_FWD_PACKED"]}      # packed arithmetic in the forward 1024-point z passes too (measured 2 % slower there)      # packed arithmetic in the forward 1024-point passes too


def _compile_and_link(out: str, extra: list, tag: str, verbose: bool) -> str:
    """nvcc -c every source concurrently (objects under build/<tag>/), then link `out`; ptxas -v output is kept in
    build/ptxas<_tag>.log."""
    nvcc = nvcc_path()
    objdir = os.path.join(ROOT, "build", tag or "default")
    os.makedirs(objdir, exist_ok=True)
    jobs = []
    for src in SOURCES:
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", "-o", obj, src]
        if verbose:
            print("[mhdflows_jl_b200] " + " ".join(cmd), file=sys.stderr)
        jobs.append((obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, failed = [], False
    for obj, p in jobs:
        o, _ = p.communicate()
        log.append(o)
        failed |= p.returncode != 0
    with open(os.path.join(ROOT, "build", f"ptxas{'_' + tag if tag else ''}.log"), "w") as f:
        f.write("".join(log))
    if failed:
        raise RuntimeError("nvcc failed:\n" + "".join(log)[-4000:])
    cmd = [nvcc] + LINK_FLAGS + ["-o", out] + [obj for obj, _ in jobs]
    if verbose:
        print("[mhdflows_jl_b200] " + " ".join(cmd), file=sys.stderr)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + (res.stdout + res.stderr)[-4000:])
    return out


def build_variant(name: str, verbose: bool = True) -> str:
    return _compile_and_link(os.path.join(HERE, f"libmhdflows_b200_{name}.so"), VARIANTS[name], name, verbose)


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("MHDF_NVCC_EXTRA", "").split()     # tuning builds, e.g. -DMHDF_SPEC_MINB=3
    return _compile_and_link(LIB, extra, "", verbose)


if __name__ == "__main__":
    variants = [a[len("--variant="):] for a in sys.argv if a.startswith("--variant=")]
    if variants:
        for v in variants:
            print(build_variant(v))
    else:
        print(build(force="--force" in sys.argv))
