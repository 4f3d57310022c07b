"""Pins the oracle against the reference's own saved artefacts (golden chains)."""
import numpy as np
from oracle import batched, kernels


def test_kin40k_chain_reproduces_printed_smse(kin40k):
    # experiments/regression_kin40k.ipynb cell[11]: softplus(theta*) printed to 16 digits
    sp = kernels.softplus(kin40k["theta_raw"])
    np.testing.assert_allclose(sp[:3], [0.17636613718898136, 2.994391934274809, 2.905302600576806], rtol=1e-15)
    Xu = kin40k["xtrain"][kin40k["xu_ids"]]
    pred = batched.predict_mean(kin40k["xtest"], Xu, sp[0], sp[1:], kin40k["mu_v"])
    smse = batched.smse(kin40k["ytest"], pred)
    # notebook prints 0.08343114079545057 (regression_kin40k.ipynb:315); summation-order noise only
    assert abs(smse - float(kin40k["smse_printed"])) < 1e-13
    np.testing.assert_allclose(kin40k["mu_v"][:3], [-51.2111428, 36.2406217, -30.7765060], rtol=1e-8)


def test_banana_chain_reproduces_125_errors(banana):
    sp = kernels.softplus(banana["theta_raw"])
    np.testing.assert_allclose(sp, [0.98563, 1.02806, 1.02154], atol=1e-5)
    x = banana["x"]; lab = banana["label"]
    Xu = x[:4000][banana["xu_ids"]]
    xt, yt = x[4000:5300], (lab[4000:5300] > 0).astype(float)        # -1 -> 0 (classification_banana.ipynb:66-73)
    m = batched.predict_mean(xt, Xu, sp[0], sp[1:], banana["mu_v"])
    errors = int(np.sum(np.abs((m > 0).astype(float) - yt)))
    assert errors == int(banana["errors_printed"]) == 125
    assert errors / 1300 == 0.09615384615384616


def test_toy_sets(toy):
    x, y = toy["xtest_toyregression"], toy["ytest_toyregression"]
    assert x.size == 600 and toy["xtrain_toyregression"].size == 50
    np.testing.assert_allclose(y, np.sinc(x), atol=3e-16)
    lab = toy["ytrain_toyclassification"]
    assert lab.size == 100 and set(np.unique(lab)) == {0.0, 1.0} and int(lab.sum()) == 60


def test_inducing_points_are_training_rows(kin40k, banana):
    assert kin40k["xu_ids"][:8].tolist() == [4185, 6438, 5980, 7904, 3036, 2976, 6951, 4021]
    assert banana["xu_ids"][:4].tolist() == [3312, 757, 444, 1147]
