"""Host-side mirror of the MHDFlows.jl problem API over the C ABI (include/mhdflows_b200.h).

Same names, keyword arguments and error behaviour as the reference for the hot path:
`Problem`, `SetUpProblemIC!`, `stepforward!`, `TimeIntegrator!`, `getCFL!`, `ProbDiagnostic`,
`Diagnostic`, `DivFreeSpectraMap` (Python cannot spell `!`, so the bang is dropped).  All compute
happens in the CUDA library; this file only marshals arguments, like julia/MHDFlowsB200.jl does with
`ccall`.  Arrays follow the reference's column-major layout: a Julia `(nx, ny, nz)` array is a NumPy
C-order array of shape `(nz, ny, nx)`; `sol` is `(Nfield, nz, ny, nx/2+1)`.

file:line citations are into the reference tree.
"""
from __future__ import annotations

import ctypes as C
import math
import time

import numpy as np

from . import _lib as L


class CPU:  # FourierFlows.CPU()
    pass


class GPU:  # FourierFlows.GPU()
    def __init__(self, device: int = 0):
        self.device = device


def nothingfunction(*args, **kw):  # pgen.jl:7
    return None


def N97ForceDriving(*args, **kw):
    """N97ForceDriving! (pgen/TaylorGreenDynamo.jl:12-16): pass as `calcF` to Problem, then call SetUpN97(prob).
    The forcing itself is applied inside the CUDA library (mhdf_set_forcing); this object is only the selector."""
    raise RuntimeError("N97ForceDriving is applied by the library; it is not called from the host")


def GetN97vars_And_function(dev=None, nx=None, ny=None, nz=None, T=np.float32):
    """GetN97vars_And_function(dev, nx, ny, nz; T) (pgen/TaylorGreenDynamo.jl:36-40) -> (usr_vars, calcF)."""
    return {}, N97ForceDriving


def SetUpN97(prob, F0=1, kf=2):
    """SetUpN97!(prob; F0, kf) (pgen/TaylorGreenDynamo.jl:18-34): Taylor-Green forcing of Nore et al. (1997),
    f = F0 (sin kf x cos kf y cos kf z, -cos kf x sin kf y cos kf z, 0), added to N[ux], N[uy] on every RHS."""
    if prob.params.calcF is not N97ForceDriving:
        raise ValueError("construct the problem with calcF=N97ForceDriving (GetN97vars_And_function)")
    g = prob.grid
    z0 = prob.rank * prob._real_shape[0] if prob.nranks > 1 else 0
    X = g.x.astype(np.float64).reshape(1, 1, -1)
    Y = g.y.astype(np.float64).reshape(1, -1, 1)
    Z = g.z.astype(np.float64)[z0:z0 + prob._real_shape[0]].reshape(-1, 1, 1)
    prob.set_forcing("ux", F0 * np.sin(kf * X) * np.cos(kf * Y) * np.cos(kf * Z))
    prob.set_forcing("uy", -F0 * np.cos(kf * X) * np.sin(kf * Y) * np.cos(kf * Z))


class ND_vars:
    """ND_vars (pgen/NegativeDamping.jl:8-13): the power P and the three real profiles fx, fy, fz of the negative-damping force
    F_i = f_i u_i.  The profiles live on the device (mhdf_set_forcing_nd); this object mirrors what the user set."""

    def __init__(self, T):
        self.T = np.dtype(T).type
        self.P, self.fx, self.fy, self.fz = 0.0, None, None, None


def NDForceDriving(*args, **kw):
    """NDForceDriving! (pgen/NegativeDamping.jl:23-45): pass as `calcF` to Problem together with the `usr_vars` of
    GetNDvars_And_function, then call SetUpND(prob, P, fx, fy, fz).  Applied inside the CUDA library (mhdf_set_forcing_nd)."""
    raise RuntimeError("NDForceDriving is applied by the library; it is not called from the host")


def GetNDvars_And_function(dev=None, nx=None, ny=None, nz=None, T=np.float32):
    """GetNDvars_And_function(dev, nx, ny, nz; T) (pgen/NegativeDamping.jl:47-51) -> (usr_vars, calcF)."""
    return ND_vars(T), NDForceDriving


def SetUpND(prob, P, fx, fy, fz):
    """SetUpND!(prob, P, fx, fy, fz) (pgen/NegativeDamping.jl:14-21)."""
    uv = prob.vars.usr_vars
    if not isinstance(uv, ND_vars) or prob.params.calcF is not NDForceDriving:
        raise ValueError("construct the problem with the usr_vars and calcF of GetNDvars_And_function")
    ptrs, keep = [], []
    for f in (fx, fy, fz):
        ptr, k = prob._real_ptr(f)
        ptrs.append(ptr)
        keep.append(k)
    L.check(prob._h, L.lib().mhdf_set_forcing_nd(prob._h, float(P), *ptrs))
    uv.P, uv.fx, uv.fy, uv.fz = float(P), fx, fy, fz


class A99_vars:
    """A99_vars (pgen/A99ForceDriving.jl:5-16; module A99GPU: pgen/A99ForceDriving_GPU.jl:7-12): `A` and `b` are the
    user-visible knobs; the spectral tables of the reference (Fk, e1x ... e2z) are not stored -- the spectral kernel
    evaluates them per mode.  `seed` keys the device random-number stream (ours: Julia's stream cannot be reproduced)."""

    def __init__(self, variant, T, seed=0):
        self.variant, self.T = variant, np.dtype(T).type
        self.A, self.b = self.T(1.0), self.T(1.0)
        self.σ2, self.kf = self.T(1.0), self.T(1.0)      # A99GPU fields σ², kf
        self.Fk_A = None                                  # normalisation inside the Fk table, set by SetUpFk
        self.seed = int(seed)
        self._pushed = None


def A99ForceDriving(*args, **kw):
    """A99ForceDriving! (pgen/A99ForceDriving.jl:33-60): pass as `calcF` to Problem together with the `usr_vars` of
    GetA99vars_And_function, then call SetUpFk(prob).  Applied inside the CUDA library (mhdf_set_forcing_a99)."""
    raise RuntimeError("A99ForceDriving is applied by the library; it is not called from the host")


def GetA99vars_And_function(dev=None, nx=None, ny=None, nz=None, T=np.float32, C=False, seed=0):
    """GetA99vars_And_function(dev, nx, ny, nz; T, C) (pgen/A99ForceDriving.jl:18-31) -> (usr_vars, calcF)."""
    if C:
        raise NotImplementedError("A99ForceDriving_Compressible! belongs to the compressible solver (outside SURVEY 8)")
    return A99_vars(L.A99_HOST, T, seed), A99ForceDriving


def _a99_integral(grid, kf, sigma2):
    """sum(exp(-(k-kf)^2/sigma2) / (k+1)^2) over the whole (nkr, nl, nm) array, one z plane at a time."""
    kr = grid.kr.astype(np.float64).reshape(1, -1)
    l = grid.l.astype(np.float64).reshape(-1, 1)
    tot = 0.0
    for m in grid.m.astype(np.float64).ravel():
        k = np.sqrt(kr * kr + l * l + m * m)
        tot += float(np.sum(np.exp(-(k - kf) ** 2 / sigma2) / (k + 1.0) ** 2))
    return tot


def SetUpFk(prob, kf=2, P=1, σ2=1, **greek):
    """SetUpFk(prob; kf, P, σ²) (pgen/A99ForceDriving.jl:93-127): A = sqrt(3 P (Lx/dx)(Ly/dy)(Lz/dz) / ∫Fk dk / dV),
    Fk = A sqrt(exp(-(k-kf)²/σ²)/2π)/k with the kr = 0 plane zeroed.  (`σ²` is accepted as a keyword too.)"""
    σ2 = greek.pop("σ²", σ2)
    if greek:
        raise TypeError(f"SetUpFk() got unexpected keyword arguments {sorted(greek)}")
    uv = prob.vars.usr_vars
    if not isinstance(uv, A99_vars) or uv.variant != L.A99_HOST or prob.params.calcF is not A99ForceDriving:
        raise ValueError("construct the problem with the usr_vars and calcF of GetA99vars_And_function")
    g = prob.grid
    integral = _a99_integral(g, float(kf), float(σ2))
    uv.Fk_A = math.sqrt(P * 3 * (g.Lx / g.dx) * (g.Ly / g.dy) * (g.Lz / g.dz) / integral * (1 / g.dx / g.dy / g.dz))
    uv.kf, uv.σ2 = uv.T(kf), uv.T(σ2)
    prob._sync_forcing(force=True)


class A99GPU:
    """module A99GPU (pgen/A99ForceDriving_GPU.jl): the same driving with its own basis vectors, a real Φ, |g_i|
    clipped to 1, Im N_u = 0 on the kr = 0 plane; `SetUpFk!` stores kf = b (:45) and the kernel multiplies A twice
    (:60-61, :89) -- both restated."""

    @staticmethod
    def A99ForceDriving(*args, **kw):
        raise RuntimeError("A99GPU.A99ForceDriving is applied by the library; it is not called from the host")

    @staticmethod
    def GetA99vars_And_function(dev=None, nx=None, ny=None, nz=None, T=np.float32, seed=0):
        """-> (usr_vars, calcF, SetUpFk!) (pgen/A99ForceDriving_GPU.jl:14-23)."""
        return A99_vars(L.A99_GPU, T, seed), A99GPU.A99ForceDriving, A99GPU.SetUpFk

    @staticmethod
    def SetUpFk(prob, kf=2.0, P=1.0, σ=1.0, b=1.0):
        """SetUpFk!(prob; kf, P, σ, b) (pgen/A99ForceDriving_GPU.jl:25-48)."""
        uv = prob.vars.usr_vars
        if not isinstance(uv, A99_vars) or uv.variant != L.A99_GPU or prob.params.calcF is not A99GPU.A99ForceDriving:
            raise ValueError("construct the problem with the usr_vars and calcF of A99GPU.GetA99vars_And_function")
        g = prob.grid
        integral = _a99_integral(g, float(kf), float(σ) ** 2)
        A = math.sqrt(P * 3 * (g.Lx / g.dx) * (g.Ly / g.dy) * (g.Lz / g.dz) / integral * (1 / g.dx / g.dy / g.dz))
        uv.A, uv.σ2, uv.b = uv.T(A), uv.T(σ ** 2), uv.T(b)
        uv.kf = uv.T(b)              # sic: `prob.vars.usr_vars.kf = T(b)` (:45)
        uv.Fk_A = float(uv.A)        # Fk = A sqrt(...) k⁻¹ inside the kernel, then N += A Fk ... (:89, :106)
        prob._sync_forcing(force=True)


class _Clock:
    """FourierFlows.Clock{T}(dt, t, step) (Problems.jl:120) backed by the library's clock."""

    def __init__(self, prob):
        self._p = prob

    def _get(self):
        t, dt, s = C.c_double(), C.c_double(), C.c_longlong()
        L.check(self._p._h, L.lib().mhdf_get_clock(self._p._h, C.byref(t), C.byref(dt), C.byref(s)))
        return t.value, dt.value, s.value

    @property
    def t(self):
        return self._get()[0]

    @t.setter
    def t(self, v):
        L.check(self._p._h, L.lib().mhdf_set_clock(self._p._h, float(v), self._get()[2]))

    @property
    def dt(self):
        return self._get()[1]

    @dt.setter
    def dt(self, v):
        L.check(self._p._h, L.lib().mhdf_set_dt(self._p._h, float(v)))

    @property
    def step(self):
        return self._get()[2]

    @step.setter
    def step(self, v):
        L.check(self._p._h, L.lib().mhdf_set_clock(self._p._h, self._get()[0], int(v)))


class _Flag:  # Problems.jl:68-79
    def __init__(self, b, e, vp=False, c=False, s=False):
        self.b, self.e, self.vp, self.c, self.s = b, e, vp, c, s


class _Grid:
    """The ThreeDGrid fields user code reads (mirror: utils/utils.jl:42-98)."""

    def __init__(self, nx, ny, nz, Lx, Ly, Lz, T):
        self.nx, self.ny, self.nz = nx, ny, nz
        self.Lx, self.Ly, self.Lz = Lx, Ly, Lz
        self.nkr, self.nl, self.nm = nx // 2 + 1, ny, nz
        self.nk = nx
        self.dx, self.dy, self.dz = Lx / nx, Ly / ny, Lz / nz
        self.T = np.dtype(T).type
        T_ = self.T
        self.x = (T_(-Lx / 2) + T_(self.dx) * np.arange(nx)).astype(T_)
        self.y = (T_(-Ly / 2) + T_(self.dy) * np.arange(ny)).astype(T_)
        self.z = (T_(-Lz / 2) + T_(self.dz) * np.arange(nz)).astype(T_)
        self.kr = (np.arange(self.nkr) * (2 * math.pi / Lx)).astype(T_).reshape(1, 1, -1)
        self.l = (np.fft.fftfreq(ny, 1.0 / ny) * (2 * math.pi / Ly)).astype(T_).reshape(1, -1, 1)
        self.m = (np.fft.fftfreq(nz, 1.0 / nz) * (2 * math.pi / Lz)).astype(T_).reshape(-1, 1, 1)
        self.aliased_fraction = 1 / 3  # Problem's aliased_fraction is not forwarded (pgen.jl:107)

    @staticmethod
    def aliased_range(nk, aliased_fraction=1 / 3):
        """FourierFlows.getaliasedwavenumbers, 1-based inclusive (iL, iR)."""
        Lf, Rf = (1 - aliased_fraction) / 2, (1 + aliased_fraction) / 2
        return math.floor(Lf * nk) + 1, math.ceil(Rf * nk)

    @property
    def Krsq(self):
        return (self.kr ** 2 + self.l ** 2 + self.m ** 2).astype(self.T)

    def retained_mask(self):
        """Modes kept by dealias!(fh, grid): shape (nz, ny, nkr)."""
        msk = np.ones((self.nm, self.nl, self.nkr), dtype=bool)
        iL, _ = self.aliased_range(self.nx)
        msk[..., iL - 1:] = False
        iL, iR = self.aliased_range(self.ny)
        msk[:, iL - 1:iR, :] = False
        iL, iR = self.aliased_range(self.nz)
        msk[iL - 1:iR, :, :] = False
        return msk


_VP_FIELDS = {"χ": 0, "chi": 0, "U₀x": 1, "U₀y": 2, "U₀z": 3, "B₀x": 4, "B₀y": 5, "B₀z": 6,
              "U0x": 1, "U0y": 2, "U0z": 3, "B0x": 4, "B0y": 5, "B0z": 6}


class _Params:
    """MHDParams / HDParams / EMHDParams (MHDParams.jl:42-94, HDParams.jl:30-51); 1-based indices.  With VP_method the
    reference's params also hold the real fields χ, U₀x, U₀y, U₀z (, B₀x, B₀y, B₀z) (datastructure.jl:80-81,94-95): assigning
    one (`prob.params.χ = mask`, the mirror of `copyto!(prob.params.χ, mask)`) sends it to the device."""

    def __init__(self, prob=None):
        object.__setattr__(self, "_p", prob)

    def __setattr__(self, name, value):
        if name in _VP_FIELDS:
            object.__getattribute__(self, "_p").set_vp_field(name, value)
        object.__setattr__(self, name, value)


class _Vars:
    """vars.ux ... : real-space fields.  Reading one runs a c2r on demand of the LAST STAGE INPUT,
    which is what the reference's vars hold after a step."""

    def __init__(self, prob, names):
        object.__setattr__(self, "_p", prob)
        object.__setattr__(self, "_names", names)

    def __getattr__(self, name):
        names = object.__getattribute__(self, "_names")
        if name in names:
            return object.__getattribute__(self, "_p").get_real(names[name], which=L.STALE)
        raise AttributeError(name)


class Problem:
    """Problem(dev; nx, ny, nz, Lx, Ly, Lz, c_s, dt, nu, n_nu, eta, n_eta, B_field, EMHD, Compressibility, Shear,
    VP_method, Dye_Module, stepper, calcF, T, aliased_fraction, usr_vars, usr_params, usr_func)  (pgen.jl:64-127).

    Greek keyword names of the reference are accepted as well (ν, nν, η, nη)."""

    def __init__(self, dev=None, *, nx=64, ny=None, nz=None, Lx=2 * math.pi, Ly=None, Lz=None, cs=0.0, dt=0.0,
                 nu=0.0, n_nu=0, eta=0.0, n_eta=0, B_field=False, EMHD=False, Compressibility=False, Shear=False,
                 VP_method=False, Dye_Module=False, stepper="RK4", calcF=nothingfunction, T=np.float32,
                 aliased_fraction=1 / 3, usr_vars=None, usr_params=None, usr_func=None,
                 rank=0, nranks=1, nccl_id=None, **greek):
        for k, v in greek.items():
            if k == "ν":
                nu = v
            elif k == "η":
                eta = v
            elif k == "nν":
                n_nu = v
            elif k == "nη":
                n_eta = v
            else:
                raise TypeError(f"Problem() got an unexpected keyword argument {k!r}")
        if isinstance(dev, CPU):
            raise L.MHDFlowsError(L.ERR_INVALID, "this build is the B200 path only: Problem(GPU(); ...) (no CPU fallback)")
        if cs == 0.0 and Compressibility:
            raise ValueError("You should define cₛ")                       # pgen.jl:98-100
        if Shear:
            raise ValueError("Shear haven't fully implemented yet!")        # pgen.jl:103-105
        if Compressibility or Dye_Module:
            raise NotImplementedError("Compressibility / Dye_Module are outside the B200 hot path (SURVEY 8)")
        if VP_method and EMHD:
            raise ValueError("VP_method: the EMHD equation has no volume-penalisation terms (MHDSolver.jl:183-270)")
        if calcF is None:
            calcF = nothingfunction
        builtin = calcF in (nothingfunction, N97ForceDriving, A99ForceDriving, A99GPU.A99ForceDriving, NDForceDriving)
        if not builtin and not callable(calcF):
            raise TypeError("calcF must be callable: calcF(N, sol, t, clock, vars, params, grid)")
        if calcF is NDForceDriving and not isinstance(usr_vars, ND_vars):
            raise ValueError("NDForceDriving needs usr_vars = the ND_vars of GetNDvars_And_function")
        if calcF is NDForceDriving and VP_method:
            raise NotImplementedError("NDForceDriving together with VP_method is not supported on this path")
        if calcF in (A99ForceDriving, A99GPU.A99ForceDriving) and not isinstance(usr_vars, A99_vars):
            raise ValueError("A99ForceDriving needs usr_vars = the A99_vars of GetA99vars_And_function")
        if EMHD and not B_field:
            raise ValueError("EMHD requires B_field=true (datastructure.jl:78-88)")
        if stepper == "HM89" and not EMHD:
            # Problems.jl:124-128: without EFlag the name goes to FourierFlows.TimeStepper, which does not know it
            raise ValueError("stepper \"HM89\" exists for EMHD problems only (Problems.jl:124-126)")
        if stepper == "HM89" and calcF is not nothingfunction:
            raise NotImplementedError("HM89 with a forcing function: RK3linearterm! (HM89.jl:182-196) is built for calcF = nothingfunction")
        if stepper not in ("RK4", "LSRK54", "HM89"):
            raise ValueError(f"stepper {stepper!r}: \"RK4\", \"LSRK54\" and (EMHD) \"HM89\" are on the B200 path")
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        Ly = Lx if Ly is None else Ly
        Lz = Lx if Lz is None else Lz
        T = np.dtype(T).type
        if T not in (np.float32, np.float64):
            raise ValueError("T must be Float32 or Float64")
        self.T = T
        self.CT = np.complex64 if T is np.float32 else np.complex128
        self.grid = _Grid(nx, ny, nz, Lx, Ly, Lz, T)
        self.flag = _Flag(B_field, EMHD, vp=bool(VP_method))
        self.stepper = stepper
        self.usr_func = list(usr_func) if usr_func else [nothingfunction]
        p = _Params(self)
        if EMHD:
            p.η, p.nη, p.bx_ind, p.by_ind, p.bz_ind = eta, 0, 1, 2, 3
            names = {"bx": 0, "by": 1, "bz": 2}
            physics, self.Nl = L.EMHD, 3
        elif B_field:
            p.ν, p.η, p.nν, p.nη = nu, eta, n_nu, 0        # nη is never forwarded (pgen.jl:114-116)
            p.ux_ind, p.uy_ind, p.uz_ind, p.bx_ind, p.by_ind, p.bz_ind = 1, 2, 3, 4, 5, 6
            names = {"ux": 0, "uy": 1, "uz": 2, "bx": 3, "by": 4, "bz": 5}
            physics, self.Nl = L.MHD, 6
        else:
            p.ν, p.nν, p.ux_ind, p.uy_ind, p.uz_ind = nu, n_nu, 1, 2, 3
            names = {"ux": 0, "uy": 1, "uz": 2}
            physics, self.Nl = L.HD, 3
        p.calcF = calcF
        p.usr_params = usr_params
        self.params = p
        self._names = names
        self.vars = _Vars(self, names)
        object.__setattr__(self.vars, "usr_vars", usr_vars)
        # slab decomposition over `nranks` processes (ours: the reference is single-device, README.md:40-41)
        self.rank, self.nranks = int(rank), int(nranks)
        self._idbuf = None
        if self.nranks > 1:
            from .dist import SlabLayout
            self.layout = SlabLayout(nx, ny, nz, self.nranks, self.rank)
            if nccl_id is None or len(nccl_id) != 128:
                raise ValueError("nranks > 1 needs the 128-byte nccl_id shared by rank 0 (dist.nccl_id_via_torch())")
            self._idbuf = C.create_string_buffer(bytes(nccl_id), 128)
            self._real_shape = (self.layout.nzl, ny, nx)
            self._spec_shape = (nz, self.layout.Kyl, nx // 2 + 1)
        else:
            self.layout = None
            self._real_shape = (nz, ny, nx)
            self._spec_shape = (nz, ny, nx // 2 + 1)
        cfg = L.Config(nx=nx, ny=ny, nz=nz, Lx=Lx, Ly=Ly, Lz=Lz, nu=float(nu), eta=float(eta), n_nu=int(n_nu),
                       dt=float(dt), physics=physics, stepper={"RK4": L.RK4, "LSRK54": L.LSRK54, "HM89": L.HM89}[stepper],
                       dtype=L.F32 if T is np.float32 else L.F64,
                       device=dev.device if isinstance(dev, GPU) else 0, rank=self.rank, nranks=self.nranks,
                       nccl_id=C.cast(self._idbuf, C.c_void_p) if self._idbuf is not None else None,
                       vp=1 if VP_method else 0, nd=1 if calcF is NDForceDriving else 0)
        h = C.c_void_p()
        code = L.lib().mhdf_create(C.byref(cfg), C.byref(h))
        if code != L.OK:
            raise L.MHDFlowsError(code, (L.lib().mhdf_last_error(None) or b"").decode())
        self._h = h
        self.clock = _Clock(self)
        self._cb = self._cb_error = None
        if not builtin:
            self._install_calcF(calcF)

    # -- arbitrary calcF! closures ----------------------------------------------------------------
    def _install_calcF(self, calcF):
        """`calcF(N, sol, t, clock, vars, params, grid)` as in the reference (pgen.jl:231-234), any Python callable: the library
        calls back at the beginning of every right-hand-side evaluation (mhdf_set_forcing_callback).  `sol` is a host copy of the
        evaluation's input (Nfield, nz, ny, nkr), `N` a zero array of that shape: what the function adds to N is uploaded and
        added to the right-hand side (additive forcings -- the reference's `N[..., ind] += F` idiom; rows are 0-based here, i.e.
        `params.ux_ind - 1`).  `vars.*` are the stale fields.  A device round trip per evaluation: the compatibility path for
        user code; N97 / A99 / negative-damping forcings run on the device."""
        def cb(_user, t):
            try:
                sol = np.stack([self.get_spectral(i, L.STAGE) for i in range(self.Nl)])
                N = np.zeros_like(sol)
                calcF(N, sol, t, self.clock, self.vars, self.params, self.grid)
                for i in range(self.Nl):
                    a = np.ascontiguousarray(N[i], dtype=self.CT)
                    L.check(self._h, L.lib().mhdf_set_forcing_spectral(self._h, i, a.ctypes.data if a.any() else None))
                return 0
            except BaseException as e:      # never unwind through the C frames: report, re-raise after the call returns
                self._cb_error = e
                return -1
        self._cb = L.FORCING_FN(cb)          # keeps the trampoline alive as long as the problem
        L.check(self._h, L.lib().mhdf_set_forcing_callback(self._h, self._cb, None))

    def _check(self, code):
        """L.check that re-raises an exception thrown inside the forcing callback."""
        err, self._cb_error = getattr(self, "_cb_error", None), None
        if err is not None:
            raise err
        L.check(self._h, code)

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        h = getattr(self, "_h", None)
        if h:
            L.lib().mhdf_destroy(h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- fields ---------------------------------------------------------------------------------
    def _field_id(self, f):
        return self._names[f] if isinstance(f, str) else int(f)

    def _real_ptr(self, arr, writable=False):
        """Address of a real field for the C ABI: a NumPy array (host) or any object with `data_ptr()` -- a torch tensor,
        host or CUDA: the library copies with cudaMemcpyDefault, so device-resident fields (the reference's CuArray
        `vars.*`) never bounce through the host.  Returns (pointer, keep-alive object)."""
        if hasattr(arr, "data_ptr"):
            want = "float32" if self.T is np.float32 else "float64"
            if tuple(arr.shape) != self._real_shape or str(arr.dtype).split(".")[-1] != want or not arr.is_contiguous():
                raise ValueError(f"expected a contiguous {want} tensor of shape {self._real_shape}")
            if getattr(arr, "is_cuda", False):      # the library copies on its own stream: the producer's stream must be done
                import torch
                torch.cuda.current_stream(arr.device).synchronize()
            return arr.data_ptr(), arr
        if writable:
            if arr.shape != self._real_shape or arr.dtype != self.T or not arr.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous array of the field's shape and dtype")
            return arr.ctypes.data, arr
        a = np.ascontiguousarray(arr, dtype=self.T)
        if a.shape != self._real_shape:
            raise ValueError(f"expected shape {self._real_shape}, got {a.shape}")
        return a.ctypes.data, a

    def set_real(self, f, arr):
        ptr, keep = self._real_ptr(arr)
        L.check(self._h, L.lib().mhdf_set_real(self._h, self._field_id(f), ptr))

    def set_forcing(self, f, arr):
        """Constant real-space forcing of field `f` (the calcF! hook for time-independent forcings); None removes it."""
        if arr is None:
            L.check(self._h, L.lib().mhdf_set_forcing(self._h, self._field_id(f), None))
            return
        a = np.ascontiguousarray(arr, dtype=self.T)
        if a.shape != self._real_shape:
            raise ValueError(f"expected shape {self._real_shape}, got {a.shape}")
        L.check(self._h, L.lib().mhdf_set_forcing(self._h, self._field_id(f), a.ctypes.data))

    def _sync_forcing(self, force=False):
        """Push the A99 parameters (usr_vars.A, b, kf, σ², seed) to the library when they changed -- the reference reads
        usr_vars on every forcing call, so user code may retune A or b between steps."""
        uv = getattr(self.vars, "usr_vars", None)
        if not isinstance(uv, A99_vars) or uv.Fk_A is None:
            return
        key = (uv.variant, float(uv.A) * float(uv.Fk_A), float(uv.kf), float(uv.σ2), float(uv.b), uv.seed)
        if not force and key == uv._pushed:
            return
        a = L.A99(variant=key[0], amp=key[1], kf=key[2], sigma2=key[3], b=key[4], seed=key[5], call=self.a99_calls())
        L.check(self._h, L.lib().mhdf_set_forcing_a99(self._h, C.byref(a)))
        uv._pushed = key

    def a99_calls(self):
        """Forcing evaluations so far (the counter word of the device random-number stream)."""
        n = C.c_ulonglong()
        L.check(self._h, L.lib().mhdf_forcing_a99_calls(self._h, C.byref(n)))
        return n.value

    def set_vp_field(self, name, arr):
        """params.χ / U₀x ... / B₀x ... of a VP_method problem (real fields of the problem's shape)."""
        which = _VP_FIELDS[name] if isinstance(name, str) else int(name)
        a = np.ascontiguousarray(arr, dtype=self.T)
        if a.shape != self._real_shape:
            raise ValueError(f"expected shape {self._real_shape}, got {a.shape}")
        L.check(self._h, L.lib().mhdf_set_vp_field(self._h, which, a.ctypes.data))

    def div_correction(self, group):
        L.check(self._h, L.lib().mhdf_div_correction(self._h, int(group)))

    def get_real(self, f, which=L.FRESH, out=None):
        """Real-space field (c2r on demand).  `out`: optional preallocated (e.g. pinned) array to receive it."""
        if out is None:
            out = np.empty(self._real_shape, dtype=self.T)
        ptr, keep = self._real_ptr(out, writable=True)
        L.check(self._h, L.lib().mhdf_get_real(self._h, self._field_id(f), which, ptr))
        return out

    def set_spectral(self, f, arr):
        a = np.ascontiguousarray(arr, dtype=self.CT)
        if a.shape != self._spec_shape:
            raise ValueError(f"expected shape {self._spec_shape}, got {a.shape}")
        L.check(self._h, L.lib().mhdf_set_spectral(self._h, self._field_id(f), a.ctypes.data))

    def get_spectral(self, f, which=L.FRESH):
        out = np.empty(self._spec_shape, dtype=self.CT)
        L.check(self._h, L.lib().mhdf_get_spectral(self._h, self._field_id(f), which, out.ctypes.data))
        return out

    @property
    def sol(self):
        """prob.sol as a host copy (Nfield, nz, ny, nkr); dealiased modes are zero."""
        return np.stack([self.get_spectral(i) for i in range(self.Nl)])

    @sol.setter
    def sol(self, arr):
        for i in range(self.Nl):
            self.set_spectral(i, arr[i])

    def calcN(self):
        """eqn.calcN!(N, sol, t, clock, vars, params, grid) on the current sol -> N (host copy)."""
        self._sync_forcing()
        out = np.empty((self.Nl,) + self._spec_shape, dtype=self.CT)
        self._check(L.lib().mhdf_calcN(self._h, out.ctypes.data))
        return out

    # -- diagnostics ----------------------------------------------------------------------------
    def energy(self, which=L.STALE):
        ke, me = C.c_double(), C.c_double()
        L.check(self._h, L.lib().mhdf_energy(self._h, which, C.byref(ke), C.byref(me)))
        return ke.value, me.value

    def helicity(self):
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        L.check(self._h, L.lib().mhdf_helicity(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def stale_stats(self):
        mx = (C.c_double * 6)()
        sm = (C.c_double * 6)()
        L.check(self._h, L.lib().mhdf_stale_stats(self._h, mx, sm))
        return np.array(mx), np.array(sm)

    def info(self):
        v = [C.c_int() for _ in range(5)]
        b = C.c_longlong()
        L.check(self._h, L.lib().mhdf_info(self._h, *[C.byref(x) for x in v], C.byref(b)))
        return dict(nfields=v[0].value, Kx=v[1].value, Kxp=v[2].value, Ky=v[3].value, Kz=v[4].value, bytes_device=b.value)

    def step_timed(self, nsteps):
        self._sync_forcing()
        ms = C.c_double()
        self._check(L.lib().mhdf_step_timed(self._h, int(nsteps), C.byref(ms)))
        return ms.value

    def profile(self, enable=True):
        L.check(self._h, L.lib().mhdf_profile(self._h, 1 if enable else 0))

    def profile_get(self):
        ms = (C.c_double * 8)()
        cnt = (C.c_longlong * 8)()
        L.check(self._h, L.lib().mhdf_profile_get(self._h, ms, cnt, 8))
        names = ["z_inverse", "y_inverse", "x_fused", "y_forward", "z_forward", "spectral", "emhd_derive", "exchange"]
        return {n: (ms[i], cnt[i]) for i, n in enumerate(names)}

    def launch_count(self):
        return L.lib().mhdf_launch_count(self._h)

    def stepper_stats(self):
        """HM89TimeStepper: (fixed-point iterations of the last step, its last error norm max |Bⁿ - B¹|), HM89.jl:61-84."""
        it, eps = C.c_longlong(), C.c_double()
        L.check(self._h, L.lib().mhdf_stepper_stats(self._h, C.byref(it), C.byref(eps)))
        return it.value, eps.value

    def __repr__(self):  # Problems.jl:142-159
        b = "ON (EMHD)" if self.flag.e else ("ON (Ideal MHD)" if self.flag.b else "OFF")
        return ("MHDFlows Problem\n  │    Funtions\n  │     ├ Compressibility: OFF\n"
                f"  │     ├──────── B-field: {b}\n  │     ├────────── Shear: OFF\n  ├─────├────── VP Method: OFF\n"
                "  │     ├──────────── Dye: OFF\n  │     └── user function: OFF\n  │\n  │     Features\n"
                "  │     ├─────────── grid: grid (on GPU)\n  │     ├───── parameters: params\n"
                "  │     ├────── variables: vars\n  └─────├─── state vector: sol\n        ├─────── equation: eqn\n"
                f"        ├────────── clock: clock\n        └──── timestepper: {self.stepper}TimeStepper")


def SetUpProblemIC(prob, *, ux=None, uy=None, uz=None, bx=None, by=None, bz=None, **unsupported):
    """SetUpProblemIC!(prob; ux, uy, uz, bx, by, bz, U₀x, U₀y, U₀z, B₀x, B₀y, B₀z) (utils/IC.jl:41-109): copy each given real
    field in and r2c it into sol; no dealias, no projection; velocity is skipped for EMHD (:69); with VP_method the wall
    velocities / fields go to params (:93-106)."""
    for k in [k for k in unsupported if k in _VP_FIELDS and k not in ("χ", "chi")]:
        v = unsupported.pop(k)
        if prob.flag.vp and v is not None and np.size(v) and (_VP_FIELDS[k] <= 3 or prob.flag.b):
            setattr(prob.params, k, v)
    if any(v is not None and len(np.shape(v)) for v in unsupported.values()):
        raise NotImplementedError(f"unsupported IC fields on this path: {sorted(unsupported)}")
    if not prob.flag.e:
        for name, arr in (("ux", ux), ("uy", uy), ("uz", uz)):
            if arr is not None and np.size(arr):
                prob.set_real(name, arr)
    if prob.flag.b:
        for name, arr in (("bx", bx), ("by", by), ("bz", bz)):
            if arr is not None and np.size(arr):
                prob.set_real(name, arr)
    return None


def Cylindrical_Mask_Function(grid, R2=0.82 * math.pi, R1=0.0, **greek):
    """Cylindrical_Mask_Function(grid; R₂, R₁) (utils/IC.jl:5-27): the VP mask χ -- 0 in the fluid R₁ <= sqrt(x²+y²) <= R₂,
    1 in the solid -- as a (nz, ny, nx) array for `prob.params.χ`.  (`R₂` / `R₁` are accepted as keywords via **{...}.)"""
    R2, R1 = greek.pop("R₂", R2), greek.pop("R₁", R1)
    if greek:
        raise TypeError(f"Cylindrical_Mask_Function() got unexpected keyword arguments {sorted(greek)}")
    x = grid.x.reshape(1, 1, -1)
    y = grid.y.reshape(1, -1, 1)
    R = np.sqrt(x * x + y * y)                       # evaluated in the grid's element type like √(xᵢ^2+yᵢ^2)
    S = np.where((R2 >= R) & (R >= R1), 0, 1).astype(grid.T)
    return np.broadcast_to(S, (grid.nz, grid.ny, grid.nx)).copy()


def stepforward(prob, nsteps=1):
    """stepforward!(prob.sol, prob.clock, prob.timestepper, prob.eqn, prob.vars, prob.params, prob.grid)
    (timestepper/timestepper.jl:4-6)."""
    prob._sync_forcing()
    prob._check(L.lib().mhdf_step(prob._h, int(nsteps)))


def DivVCorrection(prob):
    """DivVCorrection!(prob) (Solver/VPSolver.jl:101-137): sol_u -= k (k·sol_u)/k², vars.u* refreshed."""
    prob.div_correction(0)


def DivBCorrection(prob):
    """DivBCorrection!(prob) (Solver/VPSolver.jl:61-99): sol_b -= k (k·sol_b)/k², vars.b* refreshed."""
    prob.div_correction(1)


def getCFL(prob, t_diff, Coef=0.3):
    """getCFL!(prob, t_diff; Coef) (integrator.jl:158-198)."""
    dt = C.c_double()
    L.check(prob._h, L.lib().mhdf_cfl_dt(prob._h, float(Coef), float(t_diff), C.byref(dt)))
    return dt.value


def _round_sig(x, sig=3):
    if x == 0 or not math.isfinite(x):
        return x
    return round(x, sig - int(math.floor(math.log10(abs(x)))) - 1)


def ProbDiagnostic(prob):
    """ProbDiagnostic(prob) (utils/UserInterface.jl:65-86): KE, ME = round(sum(u^2) dV; sigdigits=3) from vars."""
    try:
        ke, me = prob.energy(L.STALE)
    except L.MHDFlowsError as e:
        if e.code == L.ERR_NONFINITE:
            raise FloatingPointError("detected NaN! Quit the simulation right now.") from e
        raise
    if prob.flag.e:
        return _round_sig(me)
    if prob.flag.b:
        return _round_sig(ke), _round_sig(me)
    return _round_sig(ke)


class Diagnostic:
    """Diagnostic(calc, prob; freq, nsteps, ndata) (DiagnosticWrapper.jl:14-105)."""

    def __init__(self, calc, prob, freq=1, nsteps=100, ndata=None):
        ndata = math.ceil((nsteps + 1) / freq) if ndata is None else ndata
        self.calc, self.prob, self.freq, self.N = calc, prob, freq, ndata
        self.data = [None] * ndata
        self.t = [0.0] * ndata
        self.steps = [0] * ndata
        self.data[0], self.t[0], self.steps[0] = calc(prob), prob.clock.t, prob.clock.step
        self.i = 1

    def extend(self, n=None):
        n = self.N if n is None else n
        self.data += [None] * n
        self.t += [0.0] * n
        self.steps += [0] * n

    def update(self, i):
        if i > len(self.steps):
            self.extend()
        self.data[i - 1], self.t[i - 1], self.steps[i - 1] = self.calc(self.prob), self.prob.clock.t, self.prob.clock.step
        self.i = i

    def increment(self):
        if self.prob.clock.step % self.freq == 0:
            self.update(self.i + 1)

    def __getitem__(self, s):
        if isinstance(s, str):
            return getattr(self, s)[: self.i]
        return self.data[s]

    def __call__(self):
        return self.calc(self.prob)


def increment(diags):
    for d in (diags if isinstance(diags, (list, tuple)) else [diags]):
        d.increment()


def TimeIntegrator(prob, t0, N0, *, usr_dt=0.0, CFL_Coef=0.25, CFL_function=nothingfunction, diags=(),
                   dynamic_dashboard=True, loop_number=100, save=False, save_loc="", filename="", file_number=0,
                   dump_dt=0, quiet=True):
    """TimeIntegrator!(prob, t0, N0; ...) (integrator.jl:31-156): CFL -> stepforward! -> diagnostics per step.
    Keeps the reference's quirks: clock.step is reset to 0 (:76) and the loop runs while
    N0 >= step && t0 >= t, i.e. N0+1 steps (:104).  `save=True` dumps through mhdflows_jl_b200.io.savefile (HDF5 files with the
    reference's names and datasets)."""
    file_path_and_name = ""
    if save:
        from .io import savefile
        if len(save_loc) == 0 or len(filename) == 0 or dump_dt == 0:
            raise ValueError("Save Function Turned ON but save_loc/filename/dump_dt is not declared!\n")   # :47
        file_path_and_name = save_loc + filename
        savefile(prob, file_number, file_path_and_name=file_path_and_name)                                # :50-51
        file_number += 1
    if CFL_function is not nothingfunction and usr_dt > 0.0:
        raise ValueError("User define both CFL_function and usr_dt")                                       # :202
    updateCFL = getCFL if CFL_function is nothingfunction else CFL_function
    p = prob.params
    if prob.flag.b:
        vi = p.η if prob.flag.e else max(p.ν, p.η)
        nv = p.nη if prob.flag.e else max(p.nν, p.nη)
    else:
        vi, nv = p.ν, p.nν
    g = prob.grid
    dl = min(g.Lx / g.nx, g.Ly / g.ny, g.Lz / g.nz)
    if vi == 0:
        t_diff = math.inf
    else:
        t_diff = CFL_Coef * dl ** nv / vi if nv > 1 else CFL_Coef * dl ** 2 / vi                         # :72
    t_next_save = prob.clock.t + dump_dt                                                               # :75
    prob.clock.step = 0
    usr_declared_dt = usr_dt != 0.0
    if usr_declared_dt:
        prob.clock.dt = usr_dt
    if prob.flag.vp:                                                                                   # :85-88
        DivVCorrection(prob)
        if prob.flag.b:
            DivBCorrection(prob)
    t_start = time.perf_counter()
    while N0 >= prob.clock.step and t0 >= prob.clock.t:
        if not usr_declared_dt:
            updateCFL(prob, t_diff, Coef=CFL_Coef)
        stepforward(prob)
        increment(list(diags))
        if prob.flag.vp:                                                                               # :118-122
            DivVCorrection(prob)
            if prob.flag.b:
                DivBCorrection(prob)
        for foo in prob.usr_func:
            foo(prob)
        if save and prob.clock.t >= t_next_save:                                                       # :136-141
            ProbDiagnostic(prob)
            savefile(prob, file_number, file_path_and_name=file_path_and_name)
            t_next_save += dump_dt
            file_number += 1
        if not quiet and (dynamic_dashboard or prob.clock.step % loop_number == 0):
            d = ProbDiagnostic(prob)
            print(f"           n = {prob.clock.step:8d}, t = {_round_sig(prob.clock.t):8}, diag = {d}")
    elapsed = time.perf_counter() - t_start
    ntotal = g.nx * g.ny * g.nz
    if not quiet:
        print(f"Total CPU/GPU time run = {elapsed:.3f} s, zone update per second = {prob.clock.step * ntotal / elapsed:.3f} ")
    return elapsed


def spectralline(prob, field, nbins=None):
    """spectralline(A; Lx) (utils/MHDAnalysis.jl:237-255) of a state field: (Pk, kr)."""
    g = prob.grid
    kmax = math.sqrt((g.nx // 2 * 2 * math.pi / g.Lx) ** 2 + (g.ny // 2 * 2 * math.pi / g.Ly) ** 2 + (g.nz // 2 * 2 * math.pi / g.Lz) ** 2)
    krmax = int(np.rint(kmax + 1)) if nbins is None else nbins
    Pk = (C.c_double * krmax)()
    L.check(prob._h, L.lib().mhdf_spectrum(prob._h, prob._field_id(field), Pk, krmax))
    Pk = np.array(Pk)
    kr = np.where(Pk > 0, np.arange(1, krmax + 1), 0).astype(prob.T)
    return Pk.astype(prob.T), kr


def ScaleDecomposition(prob, group="b", kf=(1, 5), which=L.STALE):
    """ScaleDecomposition(B1, B2, B3, grid; kf) (utils/MHDAnalysis.jl:54-82) of the problem's velocity (group "u") or magnetic
    field ("b"), on the device: the three components restricted to kf[0] <= |k| <= kf[1].  `which`: the reference's vars.* (STALE,
    what a script would pass) or the true state (FRESH)."""
    out = np.empty((3,) + prob._real_shape, dtype=prob.T)
    g = {"u": 0, "b": 1}[group] if isinstance(group, str) else int(group)
    L.check(prob._h, L.lib().mhdf_scale_decomposition(prob._h, g, which, float(min(kf)), float(max(kf)), out.ctypes.data))
    return out[0], out[1], out[2]


def VectorPotential(prob, which=L.STALE):
    """VectorPotential(B1, B2, B3) (utils/MHDAnalysis.jl:129-174) of the problem's magnetic field, on the device: a with
    curl a = b, div a = 0."""
    out = np.empty((3,) + prob._real_shape, dtype=prob.T)
    L.check(prob._h, L.lib().mhdf_vector_potential(prob._h, which, out.ctypes.data))
    return out[0], out[1], out[2]


def CF(prob, group="b", which=L.STALE):
    """CF(V) = fftshift(real(ifft(abs.(fft(V)).^2))) (utils/TurbStatTool.jl:67) of the three components of the problem's velocity
    ("u") or magnetic field ("b"): periodic autocorrelation functions, transforms on the device (mhdf_correlation), the fftshift
    here.  The argument is the (dealiased) state, cf. ScaleDecomposition."""
    if prob.nranks > 1:
        raise NotImplementedError("CF of a slab-decomposed problem: mhdf_correlation returns this rank's z planes without the fftshift")
    out = np.empty((3,) + prob._real_shape, dtype=prob.T)
    g = {"u": 0, "b": 1}[group] if isinstance(group, str) else int(group)
    L.check(prob._h, L.lib().mhdf_correlation(prob._h, g, which, out.ctypes.data))
    return tuple(np.fft.fftshift(out[i]) for i in range(3))


def SFC(prob, group="b", which=L.STALE):
    """SFC(V) = 2 (mean(V) - CF(V)) (TurbStatTool.jl:72, as written: the mean of V, not of V^2) per component."""
    g = {"u": 0, "b": 1}[group] if isinstance(group, str) else int(group)
    f0 = 0 if (g == 0 or prob.flag.e) else 3
    cf = CF(prob, group, which)
    n3 = prob.grid.nx * prob.grid.ny * prob.grid.nz
    means = [prob.T(prob.get_spectral(f0 + i, which)[0, 0, 0].real / n3) for i in range(3)]      # mean(V) = V^(k = 0) / N^3
    return tuple((2 * (m - c)).astype(prob.T) for m, c in zip(means, cf))


def SF2_1D(prob, group="b", which=L.STALE):
    """SF₂1D(Vx, Vz, Vy) (TurbStatTool.jl:90-120): radial two-point structure function of a vector field of the problem -- the sum of
    the three SFC cubes averaged over shells of round(|r|), r measured from element (N/2, N/2, N/2) (1-based) like the reference's
    loop (one element off the zero lag of the fftshift: reproduced)."""
    sx, sy, sz = SFC(prob, group, which)
    sfv = (sx + sy + sz).astype(np.float64)
    nz, ny, nx = sfv.shape
    R = int(math.ceil(math.sqrt((nx // 2) ** 2 + (ny // 2) ** 2 + (nz // 2) ** 2)))
    i = np.arange(1, nx + 1).reshape(1, 1, -1) - nx // 2
    j = np.arange(1, ny + 1).reshape(1, -1, 1) - ny // 2
    k = np.arange(1, nz + 1).reshape(-1, 1, 1) - nz // 2
    kk = np.rint(np.sqrt((i * i + j * j + k * k).astype(np.float64))).astype(np.int64)
    ok = kk > 0
    mask = np.bincount(kk[ok], minlength=2 * R + 1)[1:2 * R + 1].astype(np.float64)
    tot = np.bincount(kk[ok], weights=sfv[ok], minlength=2 * R + 1)[1:2 * R + 1]
    with np.errstate(invalid="ignore", divide="ignore"):
        return tot / mask


DFSM_CALL = 0x7FFFFFFF44465350      # counter tag of the device random-phase stream (csrc/kernels.cuh: DFSM_CALL_HI/LO)


def SetUpRandomPhaseIC(prob, *, seed_u=None, seed_b=None, k_peak=0.0, P=1, k0=-5 / 3 / 2):
    """`Fx, Fy, Fz = DivFreeSpectraMap(grid; ...)` followed by `SetUpProblemIC!(prob; ux = Fx, ...)` (utils/IC.jl:130-179, 41-109)
    without leaving the device: the velocity (seed_u) and / or the magnetic field (seed_b) of `prob` are set to random-phase
    power-law solenoidal fields (mhdf_set_random_phase).  Works on slab-decomposed problems (every rank fills its modes)."""
    if seed_u is not None and not prob.flag.e:
        L.check(prob._h, L.lib().mhdf_set_random_phase(prob._h, 0, int(seed_u), float(k0), float(P), float(k_peak)))
    if seed_b is not None and prob.flag.b:
        L.check(prob._h, L.lib().mhdf_set_random_phase(prob._h, 1, int(seed_b), float(k0), float(P), float(k_peak)))


def DivFreeSpectraMap(grid, *, k_peak=0.0, P=1, k0=-5 / 3 / 2, b=1, seed=0, dev=None):
    """DivFreeSpectraMap(grid; k_peak, P, k0, b) (utils/IC.jl:130-179): random-phase power-law solenoidal field, returned as
    the three real arrays Fx, Fy, Fz like the reference.  Built on the device (the reference builds it on grid.device): a
    scratch HD problem on `grid` runs mhdf_set_random_phase and the fields are read back.  `seed` keys the Philox stream of the
    phases (Julia's rand stream cannot be reproduced); `b` is unused by the reference as well.  To initialise a problem
    directly, without the host round trip, use SetUpRandomPhaseIC."""
    p = Problem(dev if dev is not None else GPU(), nx=grid.nx, ny=grid.ny, nz=grid.nz, Lx=grid.Lx, Ly=grid.Ly, Lz=grid.Lz, T=grid.T)
    try:
        SetUpRandomPhaseIC(p, seed_u=seed, k_peak=k_peak, P=P, k0=k0)
        return tuple(p.get_real(i, L.FRESH) for i in range(3))
    finally:
        p.close()
