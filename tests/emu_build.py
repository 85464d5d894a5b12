"""Compile steps of the two CPU-emulator harnesses (tests/test_kernel_emulation.py, tests/test_emulated_library.py), kept here so
tests/conftest.py can start all of them at the beginning of the session: the g++ runs (2-3 minutes each over the kernel templates)
then overlap with each other and with the rest of the suite instead of running back to back inside two module fixtures.
Test infrastructure only; nothing here is on the product path."""
import atexit
import os
import shutil
import subprocess
import tempfile
import threading

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "mhdflows_jl_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "cpu_emu")

KERNEL_PARAMS = ["scalar", "f32x2", "asan"] + (["tsan"] if os.environ.get("MHDF_EMU_TSAN") == "1" else [])
KERNEL_FLAGS = {"scalar": [], "f32x2": ["-DMHDF_F32X2"], "asan": ["-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"],
                "tsan": ["-g", "-fsanitize=thread"]}

_started = {}


def _tmpdir(tag):
    d = tempfile.mkdtemp(prefix=f"mhdf_{tag}_")
    atexit.register(shutil.rmtree, d, ignore_errors=True)
    return d


def _popen(cmd):
    return subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


class KernelJob(threading.Thread):
    """Compile one code shape of tests/cpu_emu/test_kernels.cpp, then run it straight away (the runs of the variants overlap
    with whatever the suite is doing meanwhile).  After join(): build_rc / build_log, result = (returncode, stdout, stderr)."""

    def __init__(self, cmd, exe):
        super().__init__(daemon=True)
        self.cmd, self.exe = cmd, exe
        self.build_rc, self.build_log, self.result = None, "", None

    def run(self):
        b = subprocess.run(self.cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        self.build_rc, self.build_log = b.returncode, b.stdout
        if b.returncode == 0:
            r = subprocess.run([self.exe], capture_output=True, text=True, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
            self.result = (r.returncode, r.stdout, r.stderr)


def _start_kernels(gxx):
    """{variant: KernelJob (started)}."""
    d = _tmpdir("emu")
    jobs = {}
    for v in KERNEL_PARAMS:
        out = os.path.join(d, f"emu_test_{v}")
        jobs[v] = KernelJob([gxx, "-std=c++20", "-O1", "-pthread", "-DMHDF_CPU_EMU", *KERNEL_FLAGS[v], "-I", EMU, "-I", CSRC,
                             "-I", "/usr/local/cuda/include", "-o", out, os.path.join(EMU, "test_kernels.cpp")], out)
        jobs[v].start()
    return jobs


def _start_library(gxx):
    """Objects of the whole library on the emulator + the multi-rank driver: (dir, objs, driver object, [Popen])."""
    d = _tmpdir("emu_lib")
    inc = ["-I", EMU, "-I", "/usr/local/cuda/include"]
    objs, procs = [], []
    for src in ("api.cu", "solver_f32.cu", "solver_f64.cu"):
        obj = os.path.join(d, src[:-3] + ".o")
        objs.append(obj)
        procs.append(_popen([gxx, "-std=c++20", "-O1", "-fPIC", "-pthread", "-DMHDF_CPU_EMU", *inc, "-x", "c++", "-c", os.path.join(CSRC, src), "-o", obj]))
    drv = os.path.join(d, "ranks.o")
    procs.append(_popen([gxx, "-std=c++20", "-O1", "-fPIC", "-pthread", "-c", os.path.join(EMU, "test_library_ranks.cpp"), "-o", drv]))
    return d, objs, drv, procs


_STARTERS = {"kernels": _start_kernels, "library": _start_library}


def start(name):
    """Start the compiles of harness `name` unless they are already running; returns their handle (None without g++)."""
    if name not in _started:
        gxx = shutil.which("g++")
        _started[name] = _STARTERS[name](gxx) if gxx else None
    return _started[name]
