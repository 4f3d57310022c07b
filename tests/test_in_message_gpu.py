"""`@rule MultiSGP(:in)` (GPnode/MultiSGPnode.jl:162-236) through sgp_in_logmessage and its host mirror
(nodes.multi_rule_in, multi_prod_gaussian_logpdf, multi_rule_in_laplace) against the per-node closures restated in oracle.multisgp.
Pendulum-GPSSM shape: N = 300 nodes, d_in = 2, D_out = 2, M = 48.  Tolerances: values 1e-10 relative (north_star's free-energy class of
quantity is 1e-8), analytic derivatives against central differences of the oracle closure 1e-5, Laplace mode 1e-6."""
import numpy as np
import pytest

from oracle import cubature as cub, multisgp
from test_multisgp_nodes_gpu import _setup

pytestmark = pytest.mark.gpu


def _metas(c, nd):
    M = c["M"]
    ometa = multisgp.MultiSGPMeta(cub.SRCUBATURE, c["Z"], np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), c["Kinv"], c["kern"])
    meta = nd.MultiSGPMeta(cub.SRCUBATURE, c["Z"], np.zeros((1, 1)), np.zeros((M, 1)), np.zeros((M, M)), None, c["kern"], N=c["N"])
    meta.kuu_jitter = 1e-8                                   # the oracle's K_uu^-1 was built with the same jitter (_setup)
    return ometa, meta


@pytest.mark.parametrize("gaussian_out", [True, False])
def test_log_backward_message_values(gaussian_out):
    from gaussianprocessnode_b200 import nodes as nd
    c = _setup()
    N = c["N"]
    ometa, meta = _metas(c, nd)
    q_outs = [nd.MvNormalMeanCovariance(c["Y"][n], c["Sy"][n]) if gaussian_out else nd.PointMass(c["Y"][n]) for n in range(N)]
    logpdf = nd.multi_rule_in(q_outs, nd.MvNormalMeanCovariance(c["mu_v"], c["Sigma_v"]), nd.PointMass(c["W"]), nd.PointMass(c["theta"]), meta)
    # the points ReactiveMP would ask for: the 2d+1 spherical-radial points of every node's forward message
    Xp = np.stack([cub.sigma_points(cub.SRCUBATURE, c["means"][n], c["covs"][n])[0] for n in range(N)])
    f = logpdf(Xp)
    ref = np.empty_like(f)
    for n in range(N):
        fn = multisgp.rule_in_logpdf(c["Y"][n], c["mu_v"], c["Sigma_v"], c["W"], c["theta"], ometa)
        ref[n] = [fn(x) for x in Xp[n]]
    assert np.max(np.abs(f - ref) / np.maximum(np.abs(ref), 1.0)) <= 1e-10


def test_analytic_gradient_and_hessian():
    from gaussianprocessnode_b200 import nodes as nd
    c = _setup(N=24)
    N = c["N"]
    ometa, meta = _metas(c, nd)
    q_outs = [nd.PointMass(c["Y"][n]) for n in range(N)]
    logpdf = nd.multi_rule_in(q_outs, nd.MvNormalMeanCovariance(c["mu_v"], c["Sigma_v"]), nd.PointMass(c["W"]), nd.PointMass(c["theta"]), meta)
    Xp = c["means"][:, None, :] + 0.3 * np.random.default_rng(3).normal(size=(N, 3, 2))
    f, g, H = logpdf(Xp, hess=True)
    h = 1e-5
    for n in range(0, N, 5):
        fn = multisgp.rule_in_logpdf(c["Y"][n], c["mu_v"], c["Sigma_v"], c["W"], c["theta"], ometa)
        for p in range(3):
            x = Xp[n, p]
            gfd = np.array([(fn(x + h * e) - fn(x - h * e)) / (2 * h) for e in np.eye(2)])
            assert np.linalg.norm(g[n, p] - gfd) <= 1e-5 * max(np.linalg.norm(gfd), 1.0)
            hh = 1e-4
            Hfd = np.array([[(fn(x + hh * a + hh * b) - fn(x + hh * a - hh * b) - fn(x - hh * a + hh * b) + fn(x - hh * a - hh * b)) / (4 * hh * hh)
                             for b in np.eye(2)] for a in np.eye(2)])
            assert np.linalg.norm(H[n, p] - Hfd) <= 1e-4 * max(np.linalg.norm(Hfd), 1.0)
            assert np.array_equal(H[n, p], H[n, p].T)


def test_prod_with_the_forward_message():
    # ReactiveMP.prod(::GenericProd, left::MvGaussian, right::ContinuousMultivariateLogPdf) (MultiSGPnode.jl:38-45) for the whole chain.
    # The message is a scaled log density: W = 100 I makes exp() span hundreds of orders of magnitude, exactly as in the notebook, so the
    # test uses the pendulum's own precision scale on a damped message (W / 50) to stay inside the double range for random inputs.
    from gaussianprocessnode_b200 import nodes as nd
    c = _setup(N=60)
    c["W"] = c["W"] / 50.0
    N = c["N"]
    ometa, meta = _metas(c, nd)
    q_outs = [nd.PointMass(c["Y"][n]) for n in range(N)]
    q_v = nd.MvNormalMeanCovariance(c["mu_v"] * 0.1, c["Sigma_v"] * 0.01)
    logpdf = nd.multi_rule_in(q_outs, q_v, nd.PointMass(c["W"]), nd.PointMass(c["theta"]), meta)
    lefts = [nd.MvNormalMeanCovariance(c["means"][n], c["covs"][n]) for n in range(N)]
    mu, cov = nd.multi_prod_gaussian_logpdf(lefts, logpdf)
    for n in range(N):
        fn = multisgp.rule_in_logpdf(c["Y"][n], c["mu_v"] * 0.1, c["Sigma_v"] * 0.01, c["W"], c["theta"], ometa)
        om, oc = multisgp.prod_gaussian_logpdf(c["means"][n], c["covs"][n], fn)
        assert np.linalg.norm(mu[n] - om) <= 1e-9 * max(np.linalg.norm(om), 1.0)
        assert np.linalg.norm(cov[n] - oc) <= 1e-8 * max(np.linalg.norm(oc), 1e-12)


def test_laplace_variant_finds_the_oracle_mode():
    from gaussianprocessnode_b200 import nodes as nd
    c = _setup(N=8)
    c["W"] = c["W"] / 50.0
    N = c["N"]
    ometa, meta = _metas(c, nd)
    q_outs = [nd.PointMass(c["Y"][n]) for n in range(N)]
    mu_v, Sigma_v = c["mu_v"] * 0.1, c["Sigma_v"] * 0.01
    q_ins = [nd.MvNormalMeanCovariance(c["means"][n], c["covs"][n]) for n in range(N)]
    xi, Wz, mz = nd.multi_rule_in_laplace(q_outs, q_ins, nd.MvNormalMeanCovariance(mu_v, Sigma_v), nd.PointMass(c["W"]), nd.PointMass(c["theta"]), meta)
    hits = 0
    for n in range(N):
        oxi, oW, omz = multisgp.rule_in_laplace(c["Y"][n], c["means"][n], mu_v, Sigma_v, c["W"], c["theta"], ometa)
        fn = multisgp.rule_in_logpdf(c["Y"][n], mu_v, Sigma_v, c["W"], c["theta"], ometa)
        # both are stationary points reached by ascent from the same start; the mirror must be at least as good a maximum
        assert fn(mz[n]) >= fn(omz) - 1e-9 * max(abs(fn(omz)), 1.0)
        if np.linalg.norm(mz[n] - omz) <= 1e-5 * max(np.linalg.norm(omz), 1.0):
            hits += 1
            assert np.linalg.norm(Wz[n] - oW) <= 1e-3 * max(np.linalg.norm(oW), 1e-12)
            assert np.linalg.norm(xi[n] - Wz[n] @ mz[n]) <= 1e-12 * max(np.linalg.norm(xi[n]), 1.0)
    assert hits >= N // 2          # (multi-modal messages may send the two optimisers to different modes)
