"""N > 1 host logic on CPU: world_size-2 (and 4) gloo runs of the slab layout -- the blocked exchange layout and
row tables used by the CUDA path reproduce the global (masked) rfftn."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from mhdflows_jl_b200.dist import SlabLayout, emulated_forward
dist.init_process_group("gloo")
r, P = dist.get_rank(), dist.get_world_size()
nx, ny, nz = 32, 16, 8 * P
lay = SlabLayout(nx, ny, nz, P, r)
x = np.random.default_rng(0).standard_normal((nz, ny, nx))
loc = emulated_forward(lay.scatter_real(x), lay)
ref = np.fft.rfftn(x, axes=(0, 1, 2))
# expected local compact slab
rows = lay.local_ky_rows()
exp = np.zeros((lay.Kz, lay.Kyl, lay.Kxp), dtype=complex)
ok = rows >= 0
exp[:, ok, :lay.Kx] = ref[lay.kz_full_index()][:, rows[ok], :lay.Kx]
err = np.abs(loc - exp).max() / np.abs(exp).max()
assert err < 1e-12, err
# assemble on every rank through an object gather and compare with the masked global spectrum
full_loc = np.zeros((nz, lay.Kyl, lay.nkr), dtype=complex)
full_loc[lay.kz_full_index(), :, :lay.Kx] = loc[:, :, :lay.Kx]
objs = [None] * P
dist.all_gather_object(objs, full_loc)
full = lay.assemble_spectral(objs)
msk = np.zeros_like(ref, dtype=bool)
msk[np.ix_(lay.kz_full_index(), lay.ky_full_index(), np.arange(lay.Kx))] = True
assert np.abs(full - np.where(msk, ref, 0)).max() / np.abs(ref).max() < 1e-12
assert np.array_equal(lay.local_spectral_from_full(np.where(msk, ref, 0)), full_loc)
dist.barrier()
if r == 0: print("OK", P)
"""


@pytest.mark.parametrize("world", [2, 4])
def test_slab_exchange_layout_over_gloo(world, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert f"OK {world}" in res.stdout


def test_layout_arithmetic():
    from mhdflows_jl_b200.dist import SlabLayout
    lay = SlabLayout(256, 256, 256, 8, rank=7)
    assert (lay.Kx, lay.Kxp, lay.Ky, lay.Kz, lay.Kyl, lay.nzl) == (85, 88, 170, 170, 22, 32)
    rows = lay.local_ky_rows()
    assert (rows >= 0).sum() == 170 - 7 * 22 and rows[0] == 256 - 16
    allrows = np.concatenate([SlabLayout(256, 256, 256, 8, r).local_ky_rows() for r in range(8)])
    assert sorted(allrows[allrows >= 0]) == sorted(lay.ky_full_index())
    with pytest.raises(ValueError):
        SlabLayout(32, 32, 32, 8)
    t = lay.tab_zfull(6)
    assert t[32] == lay.block_elems(6) and t[33] - t[32] == lay.Kyl * lay.Kxp


IO_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch.distributed as dist
from types import SimpleNamespace as NS
from mhdflows_jl_b200 import io
dist.init_process_group("gloo")
r, P = dist.get_rank(), dist.get_world_size()
nx, ny, nz = 16, 8, 4 * P
nzl = nz // P
full = {n: np.random.default_rng(i).standard_normal((nz, ny, nx)).astype(np.float32) for i, n in enumerate(("ux", "uy", "uz", "bx", "by", "bz"))}
class Prob:   # the attributes savefile / Restart use, backed by host arrays (no GPU here)
    def __init__(self):
        self.rank, self.nranks = r, P
        self.flag = NS(e=False, b=True)
        self.grid = NS(nx=nx, ny=ny, nz=nz)
        self.clock = NS(t=1.25)
        self._real_shape = (nzl, ny, nx)
        self.got = {}
    def get_real(self, f, which):
        return full[f][r * nzl:(r + 1) * nzl].copy()
    def set_real(self, f, a):
        assert a.shape == self._real_shape
        self.got[f] = np.array(a)
p = Prob()
path = io.savefile(p, 3, file_path_and_name=%r)
assert path.endswith("_t_0003.h5")
d = io.readMHDFlows(path)
for ds, f in io._U + io._B:          # ONE file holding the whole grid, whatever the number of ranks
    assert d[ds].shape == (nz, ny, nx) and np.array_equal(d[ds], full[f]), ds
assert float(d["time"]) == 1.25
q = Prob()
q.clock.t = 0.0
io.Restart(q, path)
for f in full:                        # every rank gets back its own z planes
    assert np.array_equal(q.got[f], full[f][r * nzl:(r + 1) * nzl]), f
assert q.clock.t == 1.25
dist.barrier()
if r == 0: print("IO OK", P)
"""


def test_savefile_and_restart_of_a_slab_decomposed_problem(tmp_path):
    """with nranks > 1 every rank used to write its own slab to the same file.  Now the slabs are gathered to
    rank 0 (one file, full datasets) and Restart! slices the rank's planes -- checked over gloo with 2 ranks."""
    script = tmp_path / "io_worker.py"
    script.write_text(IO_WORKER % (ROOT, str(tmp_path / "run")))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "IO OK 2" in res.stdout


def test_savefile_of_a_slab_problem_without_a_process_group_is_refused():
    from types import SimpleNamespace as NS
    from mhdflows_jl_b200 import io
    p = NS(rank=0, nranks=2, flag=NS(e=False, b=False), clock=NS(t=0.0),
           get_real=lambda f, w: np.zeros((2, 4, 4), np.float32))
    with pytest.raises(NotImplementedError):
        io.savefile(p, 0, file_path_and_name="/tmp/never_written")
