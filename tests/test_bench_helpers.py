"""bench.py host-side helpers (no GPU): weak-scaling grids, analytic Taylor-Green slabs, workload table."""
import math

import numpy as np

import bench


def test_weak_scaling_grids_keep_points_per_gpu():
    for w in (1, 2, 4, 8):
        nx, ny, nz = bench.grid_for(256, w)
        assert nx * ny * nz == w * 256 ** 3 and nz % w == 0
    assert bench.grid_for(256, 8) == (512, 512, 512)


def test_taylor_green_slabs_tile_the_full_field_and_are_solenoidal():
    n, dims = 16, (16, 32, 32)
    full = bench.tg_fields(n, dims=dims)
    lo = bench.tg_fields(n, dims=dims, zrange=(0, 16))
    hi = bench.tg_fields(n, dims=dims, zrange=(16, 32))
    for f, a, b in zip(full, lo, hi):
        assert np.array_equal(np.concatenate([a[0], b[0]], axis=0), f[0])
    # divergence of u and b by centred spectral derivative on the periodic box
    nx, ny, nz = dims
    k = [np.fft.fftfreq(m, 1.0 / m) * (2 * math.pi / (2 * math.pi * m / n)) for m in (nx, ny, nz)]
    for base in (0, 3):
        fh = [np.fft.fftn(full[base + i][0].astype(np.float64)) for i in range(3)]
        div = (1j * k[0].reshape(1, 1, -1) * fh[0] + 1j * k[1].reshape(1, -1, 1) * fh[1] + 1j * k[2].reshape(-1, 1, 1) * fh[2])
        assert np.abs(div).max() < 1e-4 * np.abs(fh[0]).max()


def test_workload_table_matches_baseline_configs():
    assert bench.WORKLOADS["mhd256"][:3] == ("mhd", 256, "RK4")
    assert bench.WORKLOADS["mhd512_lsrk"][:3] == ("mhd", 512, "LSRK54")
    assert bench.WORKLOADS["emhd512"][:3] == ("emhd", 512, "RK4")
    assert bench.ALG_S_PER_STEP[("mhd", "RK4")] == 384 and bench.XPASS_S_PER_LAUNCH["mhd"] == 15
    assert abs(bench.PUBLISHED_PTS_STEPS_PER_S["mhd256"] - 6.19e7) / 6.19e7 < 1e-2


def test_cufft_reference_point_never_takes_the_bench_down():
    """Without a GPU the cuFFT reference point reports `unavailable` instead of raising; with torch on the CPU device the
    bookkeeping (FFT counts of the fused and of the reference formulation) is checked through a patched device."""
    import torch
    if not torch.cuda.is_available():
        r = bench.cufft_reference_point("mhd", (16, 16, 16), reps=1)
        assert set(r) == {"unavailable"}
    # counts used for the per-RHS extrapolation
    import inspect
    src = inspect.getsource(bench.cufft_reference_point)
    assert bench.XPASS_FIELDS["mhd"] == (6, 9) and '"mhd": 36' in src and '"emhd": 51' in src


def test_taylor_green_fields_equal_the_direct_formula():
    """The plane-by-plane builder must reproduce (sin x cos y) cos z ... evaluated in Float64 and rounded, bit for bit."""
    n, dims = 16, (16, 32, 16)
    nx, ny, nz = dims
    dx = 2 * math.pi / n
    X = (np.float32(-math.pi * nx / n) + np.float32(dx) * np.arange(nx)).astype(np.float64).reshape(1, 1, -1)
    Y = (np.float32(-math.pi * ny / n) + np.float32(dx) * np.arange(ny)).astype(np.float64).reshape(1, -1, 1)
    Z = (np.float32(-math.pi * nz / n) + np.float32(dx) * np.arange(nz)).astype(np.float64).reshape(-1, 1, 1)
    want = [np.sin(X) * np.cos(Y) * np.cos(Z), -np.cos(X) * np.sin(Y) * np.cos(Z), np.zeros((nz, ny, nx)),
            np.cos(X) * np.sin(Y) * np.sin(Z), np.sin(X) * np.cos(Y) * np.sin(Z), -2 * np.sin(X) * np.sin(Y) * np.cos(Z)]
    got = bench.tg_fields(n, dims=dims)
    for w, (g, _) in zip(want, got):
        assert np.array_equal(w.astype(np.float32), g)
    import torch
    for i in range(6):
        assert np.array_equal(bench.tg_field_device(n, dims, i, "cpu").numpy(), got[i][0])


def test_defaults_are_the_north_star_configuration_and_both_arms_share_the_config():
    assert bench.DEFAULT_WORKLOAD == "mhd1024" and bench.WORKLOADS["mhd1024"][:3] == ("mhd", 1024, "RK4")
    a = bench.config_for("mhd1024", (1024, 1024, 1024), 8, True)
    b = bench.config_for("mhd1024", (1024, 1024, 1024), 8, True)
    assert a == b and a["workload"] == "mhd1024" and a["grid"] == [1024, 1024, 1024]
    (kind, ns, stepper, *_), text = bench.ref_sample("mhd1024")
    assert (kind, ns, stepper) == ("mhd", bench.REF_SAMPLE_N, "RK4") and "bounded sample" in text


def test_pruned_byte_model_counts_every_pass_once():
    info = {"Kxp": 344, "Ky": 682, "Kz": 682}
    step, x = bench.pruned_bytes("mhd", "RK4", info, (1024, 1024, 1024), 1, 6)
    cf, zf, xf = 344 * 682 * 682 * 8, 1024 * 682 * 344 * 8, 1024 * 1024 * 344 * 8
    assert x == 15 * xf
    stage = 6 * (cf + zf) + 6 * (zf + xf) + 15 * xf + 9 * (xf + zf) + 9 * (zf + cf)
    assert step == 4 * stage + (4 * 9 + 16 * 6) * cf
    assert 0.75e12 < step < 0.95e12          # ~0.85 TB per 1024^3 step
