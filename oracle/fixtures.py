"""Decoders for the reference's saved artefacts (oracle; test infrastructure only).

Used ONLY by ``tests/golden/make_golden.py`` in the build container, where /root/reference exists; the GPU box
never sees /root/reference, so tests read the committed ``tests/golden/*.npz`` instead.
JLD (HDF5) payloads are read by byte offset -- no h5py here -- following SURVEY.md section 9.5.
"""
import os
import numpy as np

REF = "/root/reference"


def _raw(name):
    with open(os.path.join(REF, "savefiles", name), "rb") as f:
        return f.read()


def flat_f8(name, count, offset=4188):
    return np.frombuffer(_raw(name)[offset:offset + 8 * count], dtype="<f8").copy()


def rows_in_file(name, table):
    """Inducing points were saved as Vector{Vector{Float64}}: each is a bit-exact row of ``table``; recover the
    row ids in file order by scanning for the row images."""
    raw = _raw(name)
    width = table.shape[1] * 8
    first = {}
    for i in range(table.shape[0]):
        first.setdefault(table[i, :1].tobytes(), []).append(i)
    ids, i = [], 0
    while i + width <= len(raw):
        hit = None
        for r in first.get(raw[i:i + 8], ()):
            if raw[i:i + width] == table[r].tobytes():
                hit = r
                break
        if hit is None:
            i += 1
        else:
            ids.append(hit)
            i += width
    return np.array(ids)


def kin40k():
    import scipy.io as sio
    d = os.path.join(REF, "data", "kin40k")
    xtrain = sio.loadmat(os.path.join(d, "kin40k_xtrain.mat"))["xtrain"].astype(np.float64)
    ytrain = sio.loadmat(os.path.join(d, "kin40k_ytrain.mat"))["ytrain"].ravel().astype(np.float64)
    xtest = sio.loadmat(os.path.join(d, "kin40k_xtest.mat"))["xtest"].astype(np.float64)
    ytest = sio.loadmat(os.path.join(d, "kin40k_ytest.mat"))["ytest"].ravel().astype(np.float64)
    ids = rows_in_file("Xu_kin40k.jld", xtrain)
    theta = flat_f8("params_optimal_kin40k.jld", 9)
    mu_v = flat_f8("qv_kin40k.jld", 600, 6244)
    Sigma_v = flat_f8("qv_kin40k.jld", 600 * 600, 11624).reshape(600, 600)
    return dict(xtrain=xtrain, ytrain=ytrain, xtest=xtest, ytest=ytest, xu_ids=ids, theta_raw=theta, mu_v=mu_v,
                Sigma_v=Sigma_v)


def banana():
    tab = np.loadtxt(os.path.join(REF, "data", "banana", "banana.csv"), delimiter=",", skiprows=1)
    x = np.ascontiguousarray(tab[:, :2]); lab = tab[:, 2]
    ids = rows_in_file("Xu_banana.jld", x[:4000])
    theta = flat_f8("params_optimal_banana.jld", 3)
    mu_v = flat_f8("qv_banana.jld", 500, 6244)
    Sigma_v = flat_f8("qv_banana.jld", 500 * 500, 10824).reshape(500, 500)
    return dict(x=x, label=lab, xu_ids=ids, theta_raw=theta, mu_v=mu_v, Sigma_v=Sigma_v)


def toy():
    out = {}
    for kind, ntr, nte in (("toyregression", 50, 600), ("toyclassification", 100, 400)):
        out["xtrain_" + kind] = flat_f8("xtrain_%s.jld" % kind, ntr)
        out["ytrain_" + kind] = flat_f8("ytrain_%s.jld" % kind, ntr)
        out["xtest_" + kind] = flat_f8("xtest_%s.jld" % kind, nte)
        out["ytest_" + kind] = flat_f8("ytest_%s.jld" % kind, nte)
    return out
