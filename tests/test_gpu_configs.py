"""GPU parity at the BASELINE.json configuration sizes (configs 2, 3, 5) against frozen digests of the CPU oracle
(tests/golden/cfg*.npz, made by tests/golden/make_golden_configs.py -- the oracle needs minutes to an hour of pocketfft at
these sizes, far outside a GPU test budget).  A digest = the spectral values at 20000 random retained modes per field, the L2
norm of every field over all retained modes and the shell spectra.  Tolerances: Float32 relative L2 <= 1e-5 (north_star)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
F32_TOL = 1e-5


@pytest.fixture(scope="module")
def M():
    import mhdflows_jl_b200 as M
    return M


def _load(name):
    p = os.path.join(GOLD, name)
    if not os.path.exists(p):
        pytest.skip(f"{name} has not been generated (tests/golden/make_golden_configs.py)")
    return np.load(p)


def _check(fields, gold, prefix, groups, tol=F32_TOL, nretained=None):
    """fields: (F, nz, ny, nkr) host array of the CUDA path; groups: list of field-index lists compared together.
    The sampled modes carry the fraction NSAMPLE / N_retained of the squared-error budget tol^2 ||field||^2 (for a dense
    spectrum this is the usual relative-L2 estimate; on the sparse Taylor-Green spectra, where most samples are exact zeros
    on both sides, the field norms, the shell spectra and the energies carry the comparison)."""
    idx, ref, norms = gold["index"], gold[prefix + "_samples"], gold[prefix + "_norms"]
    nret = nretained or _nretained(fields.shape)
    for grp in groups:
        got = np.stack([fields[f].ravel()[idx[f]] for f in grp]).astype(np.complex128)
        want = ref[grp].astype(np.complex128)
        nr = np.sqrt(sum(norms[f] ** 2 for f in grp))
        err = np.linalg.norm((got - want).ravel()) / (nr * np.sqrt(idx.shape[1] / nret))
        assert err < tol, (prefix, grp, err)
        dense = np.linalg.norm(want.ravel())
        if dense > 0.05 * nr * np.sqrt(idx.shape[1] / nret):      # the sample really carries signal: plain relative L2 too
            assert np.linalg.norm((got - want).ravel()) / dense < 3 * tol, (prefix, grp, "sample rel L2")
        nn = np.sqrt(sum(np.linalg.norm(fields[f].astype(np.complex128).ravel()) ** 2 for f in grp))
        assert abs(nn - nr) / nr < tol, (prefix, grp, "norm", nn, nr)


def _nretained(shape):
    """Modes FourierFlows' dealias! keeps on a (nz, ny, nkr) half spectrum (aliased_fraction = 1/3)."""
    import math
    _, nz, ny, nkr = shape
    nx = 2 * (nkr - 1)
    keep = lambda nk: (math.floor((1 - 1 / 3) / 2 * nk)) + (nk - math.ceil((1 + 1 / 3) / 2 * nk))
    return (math.floor((1 - 1 / 3) / 2 * nx)) * keep(ny) * keep(nz)


def _check_peaks(fields, gold, prefix, groups, tol=F32_TOL, noise=3e-3):
    """Sparse (Taylor-Green) spectra: the 256 largest modes of every field carry the signal -- compared at `tol` -- and everything
    else is rounding noise, which second derivatives amplify by k^2 in Float32 (the Float32 oracle's own EMHD calcN! carries
    7e-4 of ||N|| of it at 512^3): the noise norm of the CUDA path must stay below `noise` ||N||."""
    pi, pv = gold[prefix + "_peak_index"], gold[prefix + "_peak_values"]
    for grp in groups:
        got = np.stack([fields[f].ravel()[pi[f]] for f in grp]).astype(np.complex128)
        want = pv[grp].astype(np.complex128)
        keep = np.abs(want) >= 1e-3 * np.abs(want).max()          # below that the "peaks" of a sparse field are noise themselves
        assert keep.sum() >= 4
        got, want = np.where(keep, got, 0), np.where(keep, want, 0)
        sig = np.linalg.norm(want.ravel())
        assert np.linalg.norm((got - want).ravel()) / sig < tol, (prefix, grp, "peaks")
        total = np.sqrt(sum(np.linalg.norm(fields[f].astype(np.complex128).ravel()) ** 2 for f in grp))
        rest = np.sqrt(max(total ** 2 - np.linalg.norm(got.ravel()) ** 2, 0.0))
        assert rest < noise * total, (prefix, grp, "noise", rest / total)


def _check_spectra(M, prob, gold, prefix, nfields):
    for f in range(nfields):
        Pk, _ = M.spectralline(prob, f)
        ref = gold[prefix + "_spectra"][f]
        assert len(Pk) == len(ref)
        assert np.abs(Pk - ref).max() / ref.max() < 1e-4, (prefix, f)


def test_config3_mhd512_lsrk54_random_phase(M):
    """BASELINE config 3: MHD decaying turbulence 512^3 Float32 LSRK54, random-phase IC built ON THE DEVICE
    (mhdf_set_random_phase; IC.jl:130-179) -- IC, one RHS evaluation and the state after 2 steps against the oracle digest."""
    g = _load("cfg3_mhd512_lsrk54.npz")
    n = int(g["n"])
    gp = M.Problem(M.GPU(), nx=n, nu=float(g["nu"]), eta=float(g["eta"]), dt=float(g["dt"]), stepper="LSRK54", B_field=True)
    ic = dict(seed_u=int(g["seeds"][0]), seed_b=int(g["seeds"][1]), k0=float(g["k0"]), P=1, k_peak=0.0)
    import time
    t0 = time.perf_counter()
    M.SetUpRandomPhaseIC(gp, **ic)
    t_ic = time.perf_counter() - t0
    assert t_ic < 1.0, f"device-side DivFreeSpectraMap + SetUpProblemIC! took {t_ic:.2f} s at 512^3"
    grp = [[0, 1, 2], [3, 4, 5]]
    _check(gp.sol, g, "sol0", grp)
    _check_spectra(M, gp, g, "sol0", 6)
    _check(gp.calcN(), g, "N", grp)
    M.SetUpRandomPhaseIC(gp, **ic)
    M.stepforward(gp, 2)
    _check(gp.sol, g, "sol2", grp)
    _check_spectra(M, gp, g, "sol2", 6)
    ke, me = gp.energy(M.STALE)
    assert abs(ke - g["energy_stale"][0]) < F32_TOL * g["energy_stale"][0] and abs(me - g["energy_stale"][1]) < F32_TOL * g["energy_stale"][1]
    gp.close()


def test_config5_emhd512_rk4(M):
    """BASELINE config 5: EMHD 512^3 Float32 RK4 (Hall term), Taylor-Green b: one RHS evaluation and one step."""
    import bench
    g = _load("cfg5_emhd512_rk4.npz")
    n = int(g["n"])
    gp = M.Problem(M.GPU(), nx=n, dt=float(g["dt"]), stepper="RK4", B_field=True, EMHD=True)
    fields = bench.tg_fields(n)
    bench.set_ic(M, gp, "emhd", fields)
    _check_peaks(gp.calcN(), g, "N", [[0, 1, 2]])
    bench.set_ic(M, gp, "emhd", fields)          # calcN! refreshed the stale b: start the step from the IC state again
    M.stepforward(gp, 1)
    _check(gp.sol, g, "sol1", [[0, 1, 2]])
    _check_spectra(M, gp, g, "sol1", 3)
    me = gp.energy(M.STALE)[1]
    assert abs(me - g["energy_stale"][0]) < F32_TOL * g["energy_stale"][0]
    gp.close()


def test_config2_mhd256_energy_helicity_series_100_steps(M):
    """BASELINE config 2 over 100 steps: KE / ME of the stale vars after EVERY step, helicities every 10th step and the final
    state, against the oracle's series (north_star: "matching energy and helicity time series over 100 steps")."""
    import bench
    g = _load("cfg2_mhd256_rk4_100steps.npz")
    n = int(g["n"])
    gp = M.Problem(M.GPU(), nx=n, nu=float(g["nu"]), eta=float(g["eta"]), dt=float(g["dt"]), stepper="RK4", B_field=True)
    bench.set_ic(M, gp, "mhd", bench.tg_fields(n))
    KE, ME, hel = g["KE"], g["ME"], {int(r[0]): r[1:] for r in g["helicity"]}
    dV = (2 * np.pi / n) ** 3
    for s in range(len(KE)):
        M.stepforward(gp)
        ke, me = gp.energy(M.STALE)
        assert abs(ke - KE[s]) < F32_TOL * KE[s] and abs(me - ME[s]) < F32_TOL * ME[s], (s, ke, KE[s], me, ME[s])
        if s + 1 in hel:
            hk, hm, hc = gp.helicity()
            Hk, Hm, Hc = hel[s + 1]
            scale = KE[s] + ME[s]
            assert abs(hk - Hk) < F32_TOL * scale and abs(hm - Hm) < F32_TOL * scale / dV and abs(hc - Hc) < F32_TOL * scale, (s, hk, Hk, hm, Hm, hc, Hc)
    _check(gp.sol, g, "sol100", [[0, 1, 2], [3, 4, 5]])
    gp.close()
