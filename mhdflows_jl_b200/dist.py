"""Host-side logic of the slab decomposition (one process per GPU).

Real space is split along z (`nzl = nz / P` planes per rank); the compact spectral state along the retained ky rows
(`Kyl = ceil(Ky / P)` rows per rank, the last slab zero-padded).  A 3D transform needs one all-to-all between the
z and the y passes; the exchange buffers are `[peer][field][z'][ky'][kx]` so every piece is contiguous.  The same
index formulas are used by csrc/solver.cuh (`tabs_for`); tests/test_dist_gloo.py runs them over gloo on CPU.

The reference has no multi-device mode (README.md:40-41), so the distributed array conventions are ours:
  local real field      (nzl, ny, nx)        z planes  [rank*nzl, (rank+1)*nzl)
  local spectral field  (nz, Kyl, nx/2+1)    row j is the global compact ky row rank*Kyl + j (see `local_ky_rows`)
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib as L


def aliased_range(nk, af=1 / 3):
    return math.floor((1 - af) / 2 * nk) + 1, math.ceil((1 + af) / 2 * nk)


class SlabLayout:
    def __init__(self, nx, ny, nz, nranks, rank=0):
        self.nx, self.ny, self.nz, self.P, self.rank = nx, ny, nz, nranks, rank
        self.nkr = nx // 2 + 1
        iL, _ = aliased_range(nx)
        self.Kx = iL - 1
        self.Kxp = (self.Kx + 7) // 8 * 8
        iL, iR = aliased_range(ny)
        self.ylo, self.yhi0 = iL - 1, iR
        iL, iR = aliased_range(nz)
        self.zlo, self.zhi0 = iL - 1, iR
        self.Ky = self.ylo + (ny - self.yhi0)
        self.Kz = self.zlo + (nz - self.zhi0)
        if nz % nranks:
            raise ValueError("nz must be divisible by the number of ranks")
        self.nzl = nz // nranks
        self.Kyl = -(-self.Ky // nranks)
        if nranks > 1 and ((nranks - 1) * self.Kyl >= self.Ky or self.nzl < 2 or self.Kyl < 2):
            raise ValueError("grid too small for this many ranks (every rank needs a non-empty ky slab, "
                             "at least 2 z planes and 2 retained ky rows)")
        self.ky0 = rank * self.Kyl

    # full index of every compact row (ky and kz)
    def ky_full_index(self):
        return np.concatenate([np.arange(self.ylo), np.arange(self.yhi0, self.ny)])

    def kz_full_index(self):
        return np.concatenate([np.arange(self.zlo), np.arange(self.zhi0, self.nz)])

    def local_ky_rows(self, rank=None):
        """Full ky index of each local compact row of `rank` (-1 for the zero padding rows of the last slab)."""
        rank = self.rank if rank is None else rank
        full = self.ky_full_index()
        out = np.full(self.Kyl, -1, dtype=np.int64)
        lo, hi = rank * self.Kyl, min((rank + 1) * self.Kyl, self.Ky)
        out[: hi - lo] = full[lo:hi]
        return out

    def block_elems(self, nfields=1):
        return nfields * self.nzl * self.Kyl * self.Kxp

    def tab_zfull(self, nfields=1):
        z = np.arange(self.nz)
        return (z // self.nzl) * self.block_elems(nfields) + (z % self.nzl) * self.Kyl * self.Kxp

    def tab_ky(self, nfields=1):
        j = np.arange(self.Ky)
        return (j // self.Kyl) * self.block_elems(nfields) + (j % self.Kyl) * self.Kxp

    # host helpers for tests / drivers
    def scatter_real(self, full, rank=None):
        rank = self.rank if rank is None else rank
        return np.ascontiguousarray(full[rank * self.nzl:(rank + 1) * self.nzl])

    def local_spectral_from_full(self, full, rank=None):
        rows = self.local_ky_rows(rank)
        out = np.zeros((self.nz, self.Kyl, self.nkr), dtype=full.dtype)
        ok = rows >= 0
        out[:, ok, :] = full[:, rows[ok], :]
        return out

    def assemble_spectral(self, slabs):
        """slabs[r] = local spectral array of rank r -> full (nz, ny, nkr) array."""
        full = np.zeros((self.nz, self.ny, self.nkr), dtype=slabs[0].dtype)
        for r, s in enumerate(slabs):
            rows = self.local_ky_rows(r)
            ok = rows >= 0
            full[:, rows[ok], :] = s[:, ok, :]
        return full


def nccl_unique_id():
    buf = (C.c_char * 128)()
    code = L.lib().mhdf_nccl_unique_id(buf)
    if code != L.OK:
        raise L.MHDFlowsError(code, (L.lib().mhdf_last_error(None) or b"").decode())
    return bytes(buf)


def nccl_id_via_torch(group=None):
    """Rank 0 creates the ncclUniqueId, torch.distributed (any backend) hands it to the other ranks."""
    import torch.distributed as dist
    obj = [nccl_unique_id() if dist.get_rank(group) == 0 else None]
    dist.broadcast_object_list(obj, src=0, group=group)
    return obj[0]


def enable_peer_exchange(prob, group=None):
    """Exchange CUDA IPC handles of the transpose buffers between the ranks (through torch.distributed) so the global
    transposes become copy-engine pushes into peer HBM over NVLink (overlapping the axis passes) instead of NCCL
    send/recv kernels."""
    import torch.distributed as dist
    n = L.lib().mhdf_ipc_blob_size(prob._h)
    blob = C.create_string_buffer(n)
    L.check(prob._h, L.lib().mhdf_ipc_export(prob._h, blob))
    blobs = [None] * dist.get_world_size(group)
    dist.all_gather_object(blobs, blob.raw, group=group)
    allb = C.create_string_buffer(b"".join(blobs), n * len(blobs))
    L.check(prob._h, L.lib().mhdf_ipc_import(prob._h, allb))


def emulated_forward(real_slab, lay: SlabLayout, group=None):
    """NumPy emulation of the distributed forward transform (x r2c -> y pass -> all-to-all -> z pass) over
    torch.distributed, using the SAME blocked exchange layout and row tables as the CUDA path.
    Returns the local compact spectral field [Kz][Kyl][Kxp]."""
    import torch
    import torch.distributed as dist
    P, r = lay.P, lay.rank
    ct = np.complex64 if real_slab.dtype == np.float32 else np.complex128
    xs = np.fft.rfft(real_slab.astype(np.float64), axis=2)[:, :, : lay.Kx]                # [nzl][ny][Kx]
    ys = np.fft.fft(xs, axis=1)[:, lay.ky_full_index(), :]                                 # [nzl][Ky][Kx]
    send = np.zeros(P * lay.block_elems(1), dtype=np.complex128)
    tab = lay.tab_ky(1)
    for zl in range(lay.nzl):
        for j in range(lay.Ky):
            o = tab[j] + zl * lay.Kyl * lay.Kxp
            send[o:o + lay.Kx] = ys[zl, j]
    recv = np.zeros_like(send)
    if P > 1:
        ts, tr = torch.from_numpy(send.view(np.float64)), torch.from_numpy(recv.view(np.float64))
        dist.all_to_all_single(tr, ts, group=group)
    else:
        recv[:] = send
    ztab = lay.tab_zfull(1)
    cols = np.zeros((lay.nz, lay.Kyl * lay.Kxp), dtype=np.complex128)
    for z in range(lay.nz):
        cols[z] = recv[ztab[z]: ztab[z] + lay.Kyl * lay.Kxp]
    zs = np.fft.fft(cols, axis=0)[lay.kz_full_index()]
    return zs.reshape(lay.Kz, lay.Kyl, lay.Kxp).astype(ct)
