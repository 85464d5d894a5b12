"""The packed-FP32 forms (float2p in the strided passes of the default build; everything with -DMHDF_F32X2) must keep compiling for sm_100a and must really
map the Float32 butterflies onto the packed instructions: cross-compile one small strided pass and one fused x kernel
and look for FADD2 / FMUL2 / FFMA2 in the SASS (no GPU needed)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = r'''
#include "kernels.cuh"
using namespace mhdf;
template __global__ void mhdf::k_pass<float, 64, 8, 16, -1, false, 0>(PassArgs<float>);
template __global__ void mhdf::k_xfused<float, 64, 8, 16, PHYS_MHD, false>(XArgs<float>);
template __global__ void mhdf::k_pass<double, 64, 8, 16, -1, false, 0>(PassArgs<double>);
'''


def _opcodes(cubin):
    out = subprocess.run(["cuobjdump", "-sass", cubin], capture_output=True, text=True, check=True).stdout
    ops = {}
    for line in out.splitlines():
        parts = line.split()
        if len(parts) > 2 and parts[0].startswith("/*") and parts[0].endswith("*/"):
            op = parts[2] if parts[1].startswith("@") else parts[1]
            op = op.split(".")[0].rstrip(";")
            ops[op] = ops.get(op, 0) + 1
    return ops


@pytest.mark.skipif(shutil.which("nvcc") is None or shutil.which("cuobjdump") is None, reason="CUDA toolchain not available")
def test_packed_variant_compiles_to_packed_sass(tmp_path):
    cu = tmp_path / "inst.cu"
    cu.write_text(SRC)
    counts = {}
    for name, flags in (("scalar", ["-DMHDF_PASS_SCALAR"]), ("default", []), ("f32x2", ["-DMHDF_F32X2"])):
        cubin = str(tmp_path / f"{name}.cubin")
        cmd = ["nvcc", "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", *flags,
               "-I", os.path.join(ROOT, "mhdflows_jl_b200", "csrc"), "-o", cubin, str(cu)]
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
        assert res.returncode == 0, res.stderr[-3000:]
        counts[name] = _opcodes(cubin)
    s, d, p = counts["scalar"], counts["default"], counts["f32x2"]
    npacked = lambda c: c.get("FADD2", 0) + c.get("FMUL2", 0) + c.get("FFMA2", 0)
    packed = npacked(p)
    assert npacked(s) == 0                            # -DMHDF_PASS_SCALAR: no packed instruction anywhere
    assert 30 < npacked(d) < packed                   # default: the strided Float32 passes are packed (float2p), the x kernel is not
    assert packed > 500, p
    # Float64 is untouched by the variant, Float32 scalar FP work almost disappears
    assert p.get("DADD", 0) == s.get("DADD", 0) and p.get("DFMA", 0) == s.get("DFMA", 0)
    scalar_fp = lambda c: c.get("FADD", 0) + c.get("FMUL", 0) + c.get("FFMA", 0)
    assert scalar_fp(p) < 0.1 * scalar_fp(s), (scalar_fp(p), scalar_fp(s))
    total = lambda c: sum(v for k, v in c.items() if k != "NOP")
    assert total(p) < 0.85 * total(s), (total(p), total(s))
