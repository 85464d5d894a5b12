"""NumPy simulation of the register/shared-memory Stockham plan used by csrc/fft_core.cuh.
Mirrors the per-thread index arithmetic (thread t, register slot m, butterfly q, radix R, Ns) so the
formulas can be validated on a CPU-only box before the CUDA kernels run on a B200."""
import numpy as np


def fft_sim(x, E, sign=-1):
    N = len(x)
    Tn = N // E
    tw = np.exp(sign * 2j * np.pi * np.arange(N) / N)
    # regs[t][m] = element t + Tn*m
    regs = np.array([[x[t + Tn * m] for m in range(E)] for t in range(Tn)], dtype=complex)
    NS = 1
    while True:
        R = min(E, N // NS)
        Q = E // R
        smem = np.zeros(N, dtype=complex)
        for t in range(Tn):
            for q in range(Q):
                j = t + q * Tn
                v = np.array([regs[t][q + r * Q] for r in range(R)])
                k = j % NS
                if NS > 1:
                    for r in range(1, R):
                        v[r] *= tw[(r * k) * (N // (NS * R))]
                y = np.array([sum(v[r] * np.exp(sign * 2j * np.pi * r * rp / R) for r in range(R)) for rp in range(R)])
                for rp in range(R):
                    regs[t][q + rp * Q] = y[rp]
                    smem[(j // NS) * NS * R + k + rp * NS] = y[rp]
        if NS * R == N:
            break
        for t in range(Tn):
            for m in range(E):
                regs[t][m] = smem[t + Tn * m]
        NS *= R
    out = np.zeros(N, dtype=complex)
    for t in range(Tn):
        for m in range(E):
            out[t + Tn * m] = regs[t][m]
    return out


def c2r_sim(X, N, E):
    """X: half spectrum k=0..M (M=N/2), returns unnormalised irfft (N reals) via an M-point complex FFT."""
    M = N // 2
    k = np.arange(M)
    Xm = np.conj(X[M - k])
    Z = (X[:M] + Xm) + 1j * np.exp(2j * np.pi * k / N) * (X[:M] - Xm)
    z = fft_sim(Z, E, +1)
    out = np.empty(N)
    out[0::2] = z.real
    out[1::2] = z.imag
    return out


def r2c_sim(x, E):
    N = len(x)
    M = N // 2
    z = x[0::2] + 1j * x[1::2]
    Z = fft_sim(z, E, -1)
    k = np.arange(M)
    Zm = np.conj(Z[(M - k) % M])
    Ev = 0.5 * (Z + Zm)
    Od = -0.5j * (Z - Zm)
    return Ev + np.exp(-2j * np.pi * k / N) * Od   # k = 0..M-1 (Nyquist not produced)


if __name__ == "__main__":
    rng = np.random.default_rng(0)
    for N, E in [(8, 8), (16, 4), (16, 16), (32, 8), (64, 8), (128, 8), (128, 16), (256, 16), (512, 8), (512, 16), (1024, 16), (1024, 8), (2048, 16)]:
        x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
        for s in (-1, 1):
            ref = np.fft.fft(x) if s < 0 else np.fft.ifft(x) * N
            err = np.abs(fft_sim(x, E, s) - ref).max() / np.abs(ref).max()
            assert err < 1e-12, (N, E, s, err)
    for N, E in [(32, 4), (64, 8), (256, 8), (1024, 8)]:
        x = rng.standard_normal(N)
        X = np.fft.rfft(x)
        assert np.abs(r2c_sim(x, E) - X[:-1]).max() < 1e-10
        X[0] = X[0].real; X[-1] = 0
        xr = np.fft.irfft(X, N) * N
        assert np.abs(c2r_sim(X, N, E) - xr).max() < 1e-10
    print("fft plan simulation OK")
