"""The whole library on the CPU emulator: api.cu + solver.cuh (buffer sizing, launch arguments, register rotation, the API
boundary) and the kernel sources are compiled as plain C++ against tests/cpu_emu/cuda_host_emu.h (heap = device memory,
inert streams, kernel launches = one OS thread per CUDA thread) into a throw-away libmhdflows_b200_emu.so; the end-to-end
cases of tests/emu_lib_cases.py then run through the ordinary Python mirror with MHDF_LIB pointing at it, against the oracle,
on 16-point grids.  This is what exercises the HOST side of features written without GPU access (A99 driving, volume
penalisation, divergence corrections, second EMHD kernel form, HDF5 dumps) before their first hardware run.

The emulated library is test infrastructure: it is built into a temporary directory, never by mhdflows_jl_b200.build, and
the product has no path to it other than the MHDF_LIB tuning variable (tests/test_abi.py checks the no-GPU failure of the
real library)."""
import os
import shutil
import subprocess
import sys

import pytest

from tests import emu_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mhdflows_jl_b200", "csrc")

# case-name filters of the concurrent workers (balanced by measured run time)
GROUPS = [["smoke", "a99_host", "a99_float64", "hdf5", "hm89"],
          ["a99_gpu", "a99_lsrk54", "a99_reproducible", "div_b_correction_emhd"],
          ["closure"],
          ["div_corrections", "volume_penalisation_hd", "random_phase", "on_device"],
          ["volume_penalisation_mhd", "volume_penalisation_time", "second_emhd", "negative_damping"],
          ["regress_"]]


# slab-decomposed runs, ranks = threads of tests/cpu_emu/test_library_ranks.cpp: (ranks, environment)
RANK_RUNS = {"P=2 peer pushes + flags, z-chunk pipeline, field groups at the ends": (2, {"MHDF_ZCHUNKS": "2", "MHDF_FGROUPS": "2"}),
             "P=4 send/recv, field groups in every chunk": (4, {"MHDF_PEER": "0", "MHDF_FGROUPS": "1"})}
# Found with this harness and fixed: the multiply-high division of the blocked exchange addressing cannot represent a divisor
# of 1 -- one z plane per pipeline chunk (or per rank) silently scrambled the exchange; such shapes now fall back / are refused.


@pytest.fixture(scope="module")
def emu_results():
    """Build the emulated library (+ the multi-rank driver) once (compiles started at session start by tests/conftest.py), then run
    every case group and rank configuration concurrently."""
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    d, objs, drv, procs = emu_build.start("library")
    for p in procs:
        out, _ = p.communicate(timeout=900)
        assert p.returncode == 0, out[-4000:]
    lib, exe = os.path.join(d, "libmhdflows_b200_emu.so"), os.path.join(d, "ranks_emu")
    for cmd in ([gxx, "-shared", "-pthread", "-o", lib] + objs + ["-ldl"], [gxx, "-pthread", "-o", exe, drv] + objs + ["-ldl"]):
        res = subprocess.run(cmd, capture_output=True, text=True)
        assert res.returncode == 0, res.stderr[-4000:]
    env = dict(os.environ, MHDF_LIB=lib)
    env.pop("MHDF_EMHD2", None)
    env.pop("MHDF_ZCHUNKS", None)
    running = {("cases", i): subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "emu_lib_cases.py")] + g, stdout=subprocess.PIPE,
                                              stderr=subprocess.STDOUT, text=True, env=env, cwd=ROOT) for i, g in enumerate(GROUPS)}
    for name, (P, extra) in RANK_RUNS.items():
        running[("ranks", name)] = subprocess.Popen([exe, str(P)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                                    env=dict(env, MHDF_RANKS_QUICK="1", **extra))
    results = {}
    for key, p in running.items():
        out, _ = p.communicate(timeout=2400)
        results[key] = (p.returncode, out)
    return results


def test_end_to_end_cases_on_the_emulated_library(emu_results):
    runs = [v for k, v in emu_results.items() if k[0] == "cases"]
    text = "\n".join(out for _, out in runs)
    fails = [l for l in text.splitlines() if l.startswith("FAIL")]
    assert not fails and all(rc == 0 for rc, _ in runs), text[-6000:]
    passed = [l.split()[1] for l in text.splitlines() if l.startswith("PASS")]
    # every case ran exactly once
    import ast
    tree = ast.parse(open(os.path.join(ROOT, "tests", "emu_lib_cases.py")).read())
    names = [n.name for n in tree.body if isinstance(n, ast.FunctionDef) and any(getattr(d, "id", "") == "case" for d in n.decorator_list)]
    assert sorted(passed) == sorted(names), (sorted(passed), sorted(names))


@pytest.mark.parametrize("config", list(RANK_RUNS))
def test_slab_runs_on_the_emulated_library_equal_the_single_rank_run(emu_results, config):
    """Ranks as threads, NCCL / CUDA IPC replaced by in-process stand-ins (cuda_host_emu.h): the spectral state of the P-rank run is
    bit-identical to the single-rank run -- EMHD Float64, MHD with A99 driving + volume penalisation + DivVCorrection! and the EMHD HM89
    stepper (fixed-point error norm reduced over the ranks) and a calcF! host callback that uploads per-rank forcing slabs between the
    stages of a step, through the copy-push transport with the z-chunk pipelined path and
    through send/recv on 4 ranks.  Streams are synchronous here: this checks layouts, offsets, slab bounds and reductions, not event ordering."""
    rc, out = emu_results[("ranks", config)]
    lines = [l for l in out.splitlines() if "ranks-vs-single" in l]
    assert rc == 0 and len(lines) == 4 and all(l.startswith("PASS") for l in lines), out[-4000:]
    if "pipeline" in config:     # the pipelined path really ran: it launches its passes once per z chunk
        for l in lines:
            a, b = l.split("launches ")[1].rstrip(")").split(" -> ")
            assert int(b) > int(a), l


@pytest.mark.skipif(os.environ.get("MHDF_EMU_SANITIZE_LIB") != "1", reason="about 2 h: the whole library under ASan + UBSan + LeakSanitizer "
                    "(set MHDF_EMU_SANITIZE_LIB=1); last result in profiles/r02_emulator_sanitizers.txt")
def test_whole_library_under_address_sanitizer(tmp_path):
    """api.cu + solver.cuh + kernels + the C-ABI driver tests/cpu_emu/test_library_sanitize.cpp, all with -fsanitize=address,undefined:
    the solver's device buffers are heap blocks of exactly the computed sizes, so any overrun is reported."""
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not available")
    san = ["-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"]
    objs, procs = [], []
    for src in ("api.cu", "solver_f32.cu", "solver_f64.cu"):
        obj = str(tmp_path / (src[:-3] + ".o"))
        objs.append(obj)
        procs.append(subprocess.Popen([gxx, "-std=c++20", "-O1", "-pthread", "-DMHDF_CPU_EMU", *san, "-I", os.path.join(ROOT, "tests", "cpu_emu"),
                                       "-I", "/usr/local/cuda/include", "-x", "c++", "-c", os.path.join(CSRC, src), "-o", obj],
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    drv = str(tmp_path / "drv.o")
    res = subprocess.run([gxx, "-std=c++20", "-O1", *san, "-c", os.path.join(ROOT, "tests", "cpu_emu", "test_library_sanitize.cpp"), "-o", drv],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    for p in procs:
        out, _ = p.communicate(timeout=1800)
        assert p.returncode == 0, out[-4000:]
    exe = str(tmp_path / "libsan")
    res = subprocess.run([gxx, "-fsanitize=address,undefined", "-pthread", "-o", exe, drv] + objs + ["-ldl"], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-4000:]
    res = subprocess.run([exe], capture_output=True, text=True, timeout=7200, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=1"))
    assert res.returncode == 0 and "ERROR: AddressSanitizer" not in res.stderr and "runtime error:" not in res.stderr \
        and "LeakSanitizer" not in res.stderr, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("PASS") == 15 and "0 failure(s)" in res.stdout
