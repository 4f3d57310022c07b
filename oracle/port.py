"""ctypes wrapper of oracle/sweep_port.c (oracle; test infrastructure / reported CPU baseline only)."""
import ctypes
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "sweep_port.c")
PREBUILT = os.path.join(HERE, "_build", "libsgp_port.so")
_lib = None


def build(native=False, out=None):
    """gcc -O3 [-march=native].  The prebuilt copy is generic x86-64 so that it runs on any box; bench.py asks for a
    native rebuild on the box it times on (falls back to the prebuilt one if gcc is missing there)."""
    out = out or PREBUILT
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cmd = ["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", SRC, "-o", out, "-lm"]
    if native:
        cmd.insert(2, "-march=native")
    subprocess.run(cmd, check=True)
    return out


def load(native=False):
    global _lib
    if _lib is not None and not native:
        return _lib
    path = PREBUILT
    if native:
        try:
            path = build(True, os.path.join(tempfile.mkdtemp(prefix="sgp_port_"), "libsgp_port_native.so"))
        except Exception:
            path = PREBUILT
    if not os.path.exists(path):
        build(False)
    lib = ctypes.CDLL(path)
    dp = ctypes.POINTER(ctypes.c_double)
    lib.sgp_port_sweep.restype = None
    lib.sgp_port_sweep.argtypes = [ctypes.c_long, ctypes.c_int, ctypes.c_int, dp, dp, dp, ctypes.c_double, dp, ctypes.c_double, dp, dp, dp, dp]
    lib.sgp_port_sweep_mt.restype = ctypes.c_int
    lib.sgp_port_sweep_mt.argtypes = lib.sgp_port_sweep.argtypes + [ctypes.c_int]
    lib.sgp_port_flush.restype = ctypes.c_int
    lib.sgp_port_flush.argtypes = [ctypes.c_int, dp, dp, dp, dp, dp]
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def sweep(X, y, Z, variance, ell, w, Lambda0=None, xi0=None, native=False, threads=1):
    """Reference-schedule sweep over the rows of X; returns (xi, Lambda) = prior + sum of the N per-point messages.
    threads != 1: the OpenMP variant (0 = all the host's threads); `sweep.last_threads` holds the number actually used."""
    lib = load(native)
    X = np.ascontiguousarray(X, dtype=np.float64); Z = np.ascontiguousarray(Z, dtype=np.float64)
    X = X[:, None] if X.ndim == 1 else X
    Z = Z[:, None] if Z.ndim == 1 else Z
    N, D = X.shape; M = Z.shape[0]
    y = np.ascontiguousarray(y, dtype=np.float64)
    ell = np.ascontiguousarray(np.broadcast_to(np.asarray(ell, dtype=np.float64), (D,)))
    Lam = np.zeros((M, M), order="F") if Lambda0 is None else np.asfortranarray(np.array(Lambda0, dtype=np.float64))
    xi = np.zeros(M) if xi0 is None else np.array(xi0, dtype=np.float64)
    buf = np.empty((M, M), order="F"); k = np.empty(M)
    if threads == 1:
        lib.sgp_port_sweep(N, D, M, _p(X), _p(y), _p(Z), float(variance), _p(ell), float(w), _p(Lam), _p(xi), _p(buf), _p(k))
        sweep.last_threads = 1
    else:
        sweep.last_threads = lib.sgp_port_sweep_mt(N, D, M, _p(X), _p(y), _p(Z), float(variance), _p(ell), float(w), _p(Lam), _p(xi), _p(buf), _p(k),
                                                   int(threads))
    return xi, Lam


sweep.last_threads = 1


def flush(Lambda, xi):
    lib = load()
    M = xi.size
    Lam = np.asfortranarray(np.array(Lambda, dtype=np.float64))
    mu = np.empty(M); Sigma = np.empty((M, M), order="F"); UvL = np.empty((M, M), order="F")
    info = lib.sgp_port_flush(M, _p(Lam), _p(np.ascontiguousarray(xi, dtype=np.float64)), _p(mu), _p(Sigma), _p(UvL))
    if info:
        raise np.linalg.LinAlgError("not positive definite at pivot %d" % info)
    return mu, Sigma, UvL.T
