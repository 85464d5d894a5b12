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
import shutil

import pytest

from tests import emu_build

PARAMS = emu_build.KERNEL_PARAMS


@pytest.fixture(scope="module")
def emu_runs():
    """Every variant is built and then run by its own background job, started at session start by tests/conftest.py:
    {variant: (returncode, stdout, stderr)}."""
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    jobs = emu_build.start("kernels")
    res = {}
    for v, job in jobs.items():
        job.join(timeout=3900)
        assert not job.is_alive(), f"{v}: emulator run did not finish"
        assert job.build_rc == 0, f"{v}: " + job.build_log[-4000:]
        res[v] = job.result
    return res


@pytest.mark.parametrize("variant", PARAMS)
def test_kernels_on_the_cpu_emulator(emu_runs, variant):
    rc, stdout, stderr = emu_runs[variant]
    assert "ERROR: AddressSanitizer" not in stderr and "runtime error:" not in stderr and "WARNING: ThreadSanitizer" not in stderr, stderr[-4000:]
    lines = stdout.strip().splitlines()
    fails = [l for l in lines if l.startswith("FAIL")]
    assert rc == 0 and not fails, "\n".join(fails) + stderr[-2000:]
    assert lines[-1].startswith("ALL PASS")
    names = " ".join(lines)
    for needle in ("pass forward N=128", "slab inverse leg bit-identical, NZC=4", "xfused MHD N=1024", "xfused EMHD N=128",
                   "xplain c2r N=1024", "spectral phys=1 mode=5", "P=2 rank=1", "emhd_derive", "pack / unpack",
                   "philox4x32-10 known answers", "A99 forcing host variant f32", "A99 forcing gpu variant f64", "divclean f32", "xfused VP MHD N=128",
                   "spectral VP phys=1", "xfused EMHD second form == first form N=1024"):
        assert needle in names, needle
