"""The oracle reproduces the committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py).
Guards the oracle against silent drift; the same vectors check the CUDA path in test_gpu_golden.py."""
import os

import numpy as np
import pytest

from tests.golden import make_golden as G

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("name", sorted(G.CASES))
def test_oracle_matches_golden(name):
    ref = np.load(os.path.join(HERE, "golden", name + ".npz"))
    got = G.run_case(name)
    f32 = G.CASES[name][2] is np.float32
    tol = 2e-6 if f32 else 1e-12
    from oracle import mhdflows_oracle as O
    for key in ("N0", "sol"):
        for i in range(ref[key].shape[0]):
            if np.linalg.norm(ref[key][i]) > 0:
                assert O.rel_l2(got[key][i], ref[key][i]) < tol, (name, key, i)
    assert np.allclose(got["energies"], ref["energies"], rtol=10 * tol)
    assert abs(float(got["t"]) - float(ref["t"])) < 1e-12
