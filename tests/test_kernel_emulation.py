"""The CUDA kernels, compiled as plain C++ (-DMHDF_CPU_EMU) and run on the CPU with one OS thread per CUDA thread
(tests/cpu_emu/cuda_emu.h): index arithmetic, barrier placement, warp-shuffle patterns and the blocked exchange
addressing of the very source the GPU runs are checked against naive DFTs / direct formulas on tiny problems.
Covers: strided passes (16..128 points, f32/f64), one- and two-level slab addressing (bit-identical to the single-rank
passes), the fused x kernel for HD / MHD / EMHD up to 1024-point rows, the plain x passes, the spectral kernel in every
stage mode (with forcing, hyperviscosity and the gathered mirror plane), emhd_derive and pack/unpack.
Runs three times: the default (scalar) code shape, the packed-FP32 shape (-DMHDF_F32X2: the formulas that map onto the
sm_100 add/mul/fma.f32x2 instructions, with the packed primitives emulated lane by lane), and the default shape under
AddressSanitizer + UBSan -- every global array of the harness is a host vector of exactly the size the solver would
allocate, so an out-of-bounds index in a kernel is a heap-buffer-overflow here (the CPU stand-in for compute-sanitizer
memcheck; MHDF_EMU_TSAN=1 adds a ThreadSanitizer run = racecheck between the emulated CUDA threads, ~10 min)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


PARAMS = ["scalar", "f32x2", "asan"] + (["tsan"] if os.environ.get("MHDF_EMU_TSAN") == "1" else [])


@pytest.fixture(scope="module", params=PARAMS)
def emu_binary(request, tmp_path_factory):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    out = str(tmp_path_factory.mktemp("emu") / "emu_test")
    extra = {"f32x2": ["-DMHDF_F32X2"], "asan": ["-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"],
             "tsan": ["-g", "-fsanitize=thread"]}.get(request.param, [])
    cmd = [gxx, "-std=c++20", "-O1", "-pthread", "-DMHDF_CPU_EMU", *extra, "-I", os.path.join(ROOT, "tests", "cpu_emu"),
           "-I", os.path.join(ROOT, "mhdflows_jl_b200", "csrc"), "-I", "/usr/local/cuda/include",
           "-o", out, os.path.join(ROOT, "tests", "cpu_emu", "test_kernels.cpp")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-4000:]
    return out


def test_kernels_on_the_cpu_emulator(emu_binary):
    res = subprocess.run([emu_binary], capture_output=True, text=True, timeout=3000, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert "ERROR: AddressSanitizer" not in res.stderr and "runtime error:" not in res.stderr and "WARNING: ThreadSanitizer" not in res.stderr, res.stderr[-4000:]
    lines = res.stdout.strip().splitlines()
    fails = [l for l in lines if l.startswith("FAIL")]
    assert res.returncode == 0 and not fails, "\n".join(fails) + res.stderr[-2000:]
    assert lines[-1].startswith("ALL PASS")
    names = " ".join(lines)
    for needle in ("pass forward N=128", "slab inverse leg bit-identical, NZC=4", "xfused MHD N=1024", "xfused EMHD N=128",
                   "xplain c2r N=1024", "spectral phys=1 mode=5", "spectral2 == spectral phys=1 mode=2", "P=2 rank=1", "emhd_derive", "pack / unpack",
                   "philox4x32-10 known answers", "A99 forcing host variant f32", "A99 forcing gpu variant f64", "divclean f32", "xfused VP MHD N=128",
                   "spectral VP phys=1", "xfused EMHD second form == first form N=1024"):
        assert needle in names, needle
