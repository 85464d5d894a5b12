"""Slab-decomposed (multi-GPU) path: needs >= 2 GPUs on the box; skipped otherwise.  One process per GPU (torchrun),
NCCL rendezvous on 127.0.0.1.  The distributed run must reproduce the single-GPU run of the same problem -- the same
kernels do the same arithmetic per row / column, so the comparison is for exact equality of the spectral state."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("peer", ["1", "0"])
def test_slab_run_equals_single_gpu(peer):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    env = dict(os.environ, MHDF_PEER=peer)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "dist_check.py"), "check64"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("dist-vs-single")]
    assert len(lines) == 3, res.stdout[-3000:]
    for l in lines:
        m = re.search(r"spectral max rel diff ([0-9.e+-]+)\s+real ([0-9.e+-]+)\s+diag rel ([0-9.e+-]+)", l)
        assert m, l
        assert float(m.group(1)) == 0.0 and float(m.group(2)) == 0.0, l      # bit-identical state
        assert float(m.group(3)) < 1e-10, l                                 # reductions: summation order differs


def test_pipelined_slab_run_equals_single_gpu():
    if _ngpu() < 2:
        pytest.skip("needs at least 2 GPUs")
    env = dict(os.environ, MHDF_PEER="1", MHDF_ZCHUNKS="2")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29543", os.path.join(ROOT, "tools", "dist_check.py"), "check64"],
                         capture_output=True, text=True, timeout=150, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("dist-vs-single")]
    assert len(lines) == 3
    for l in lines:
        m = re.search(r"spectral max rel diff ([0-9.e+-]+)\s+real ([0-9.e+-]+)", l)
        assert m and float(m.group(1)) == 0.0 and float(m.group(2)) == 0.0, l


def test_driven_slab_run_equals_single_gpu():
    """A99 random driving + DivVCorrection!/DivBCorrection! on 2 GPUs: the Philox counter is the global mode index, so the
    slab run draws the same random numbers as the single-GPU run and the state stays bit-identical."""
    if _ngpu() < 2:
        pytest.skip("needs at least 2 GPUs")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29545", os.path.join(ROOT, "tools", "dist_check.py"), "forcing64"],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, MHDF_PEER="1"), cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("dist-vs-single driven")]
    assert len(lines) == 2
    for l in lines:
        m = re.search(r"spectral max rel diff ([0-9.e+-]+)\s+real ([0-9.e+-]+)", l)
        assert m and float(m.group(1)) == 0.0 and float(m.group(2)) == 0.0, l


def test_hm89_slab_run_equals_single_gpu():
    """HM89TimeStepper on 2 GPUs: the error norm of the fixed-point loop is reduced over the ranks (max), so every rank leaves the
    loop in the same iteration and the state stays bit-identical to the single-GPU run."""
    if _ngpu() < 2:
        pytest.skip("needs at least 2 GPUs")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29547", os.path.join(ROOT, "tools", "dist_check.py"), "hm89_64"],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, MHDF_PEER="1"), cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("dist-vs-single emhd")]
    assert len(lines) == 2
    for l in lines:
        m = re.search(r"spectral max rel diff ([0-9.e+-]+)\s+real ([0-9.e+-]+)", l)
        assert m and float(m.group(1)) == 0.0 and float(m.group(2)) == 0.0, l
