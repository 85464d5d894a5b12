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
SOURCES = [os.path.join(CSRC, "api.cu")]
DEPS = SOURCES + [os.path.join(CSRC, "kernels.cuh"), os.path.join(CSRC, "fft_core.cuh"),
                  os.path.join(ROOT, "include", "mhdflows_b200.h")]

NVCC_FLAGS = ["-O3", "-std=c++17", "-I/usr/include", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-Xptxas", "-v", "-ldl"]


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
VARIANTS = {"f32x2": ["-DMHDF_F32X2"]}


def build_variant(name: str, verbose: bool = True) -> str:
    out = os.path.join(HERE, f"libmhdflows_b200_{name}.so")
    cmd = [nvcc_path()] + NVCC_FLAGS + VARIANTS[name] + ["-o", out] + SOURCES
    if verbose:
        print("[mhdflows_jl_b200] " + " ".join(cmd), file=sys.stderr)
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(ROOT, "build", f"ptxas_{name}.log"), "w") as f:
        f.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + (res.stdout + res.stderr)[-4000:])
    return out


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return LIB
    extra = os.environ.get("MHDF_NVCC_EXTRA", "").split()     # tuning builds, e.g. -DMHDF_SPEC_MINB=3
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + ["-o", LIB] + SOURCES
    if verbose:
        print("[mhdflows_jl_b200] " + " ".join(cmd), file=sys.stderr)
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(ROOT, "build", "ptxas.log"), "w") as f:
        f.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + (res.stdout + res.stderr)[-4000:])
    return LIB


if __name__ == "__main__":
    variants = [a[len("--variant="):] for a in sys.argv if a.startswith("--variant=")]
    if variants:
        for v in variants:
            print(build_variant(v))
    else:
        print(build(force="--force" in sys.argv))
