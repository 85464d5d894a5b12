import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_finish(session):
    """Start the g++ compiles of the CPU-emulator harnesses now, so they overlap with each other and with the rest of the suite."""
    if session.config.option.collectonly:
        return
    from tests import emu_build
    files = {os.path.basename(str(item.fspath)) for item in session.items}
    for name, f in (("library", "test_emulated_library.py"), ("kernels", "test_kernel_emulation.py")):
        if f in files and any(os.path.basename(str(i.fspath)) == f and not i.get_closest_marker("skip") for i in session.items):
            emu_build.start(name)
