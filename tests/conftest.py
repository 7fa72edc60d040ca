import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def random_system(D, K, seed, hermitian=True, scale=1.0, unitary_targets=False):
    """Seeded random drift/controls/states.  Non-Hermitian generators exercise the Liouvillian case."""
    rng = np.random.default_rng(seed)

    def mat():
        X = rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))
        if hermitian:
            X = (X + X.conj().T) / 2
        return X * scale / np.sqrt(D)

    A = mat()
    B = [mat() for _ in range(K)]
    if unitary_targets:
        Xi = np.linalg.qr(rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D)))[0]
        Xt = np.linalg.qr(rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D)))[0]
    else:
        Xi = rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))
        Xt = rng.standard_normal((D, D)) + 1j * rng.standard_normal((D, D))
        Xi = Xi @ Xi.conj().T / D
        Xt = Xt @ Xt.conj().T / D
    return A, B, Xi, Xt


def assert_parity(F, G, Fo, Go, ftol=1e-10, gtol=1e-8):
    """north_star tolerances: 1e-10 relative on the figure of merit, 1e-8 (inf-norm relative) on the gradient."""
    assert abs(F - Fo) <= ftol * max(1.0, abs(Fo)), f"fom {F} vs oracle {Fo}"
    if Go is not None:
        scale = max(np.max(np.abs(Go)), 1e-6)      # absolute floor 1e-14: entries that vanish by symmetry are rounding noise
        err = np.max(np.abs(np.asarray(G) - Go))
        assert err <= gtol * scale, f"gradient max err {err:.3e} vs scale {scale:.3e}"
