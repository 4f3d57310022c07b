"""The C port of the reference's per-point schedule (the reported CPU baseline) agrees with the NumPy oracle."""
import numpy as np
from oracle import batched, port, unisgp, kernels


def test_c_port_equals_numpy_oracle():
    rng = np.random.default_rng(0)
    N, D, M = 300, 8, 40
    X = rng.normal(size=(N, D)); y = rng.normal(size=N); Z = rng.normal(size=(M, D))
    ell = 1.0 + rng.random(D); var = 1.4; w = 12.0
    xi, Lam = port.sweep(X, y, Z, var, ell, w, Lambda0=np.eye(M) / 50.0)
    _, p1, p2, _ = batched.psi_stats_point(X, y, Z, var, ell)
    assert np.linalg.norm(Lam - (np.eye(M) / 50.0 + w * p2)) / np.linalg.norm(Lam) < 1e-13
    assert np.linalg.norm(xi - w * p1) / np.linalg.norm(xi) < 1e-13
    mu, Sigma, Uv = port.flush(Lam, xi)
    o_mu, o_Sig, o_Uv, _, _ = batched.posterior_v(np.zeros(M), np.eye(M) / 50.0, w, p1, p2)
    assert np.linalg.norm(mu - o_mu) / np.linalg.norm(o_mu) < 1e-9
    assert np.linalg.norm(Sigma - o_Sig) / np.linalg.norm(o_Sig) < 1e-9
    assert np.linalg.norm(Uv - o_Uv) / np.linalg.norm(o_Uv) < 1e-9
