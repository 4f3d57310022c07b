"""Regenerates tests/golden/*.npz from the reference's own data and saved posteriors.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
The fixtures pin the oracle: the kin40k chain must reproduce the notebook's printed
SMSE 0.08343114079545057 (experiments/regression_kin40k.ipynb:315) and the banana chain exactly 125 errors out of
1300 (experiments/classification_banana.ipynb:316-317).
"""
import os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import fixtures  # noqa: E402


def main():
    k = fixtures.kin40k()
    assert k["xu_ids"].size == 600 and int(k["xu_ids"].sum()) == 2913678
    S = k["Sigma_v"]
    np.savez_compressed(os.path.join(HERE, "kin40k_chain.npz"),
                        xtrain=k["xtrain"], ytrain=k["ytrain"], xtest=k["xtest"], ytest=k["ytest"],
                        xu_ids=k["xu_ids"], theta_raw=k["theta_raw"], mu_v=k["mu_v"],
                        Sigma_v_diag=np.diag(S).copy(), Sigma_v_fro=np.linalg.norm(S), Sigma_v_rowsum=S.sum(1),
                        smse_printed=np.float64(0.08343114079545057))
    b = fixtures.banana()
    assert b["xu_ids"].size == 500
    np.savez_compressed(os.path.join(HERE, "banana_chain.npz"), x=b["x"], label=b["label"], xu_ids=b["xu_ids"],
                        theta_raw=b["theta_raw"], mu_v=b["mu_v"], Sigma_v_diag=np.diag(b["Sigma_v"]).copy(),
                        errors_printed=np.int64(125))
    np.savez_compressed(os.path.join(HERE, "toy_sets.npz"), **fixtures.toy())
    for f in ("kin40k_chain.npz", "banana_chain.npz", "toy_sets.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
