"""Thin object wrapper over the libsgp C ABI (host NumPy arrays in / out).  One SGPContext per GPU."""
import ctypes
import numpy as np

from . import _lib

SE, MATERN32, MATERN52 = 0, 1, 2
SRCUBATURE, GENUT, GAUSSHERMITE, CLOSED_FORM_SE, POINT = 0, 1, 2, 3, 4


class SGPError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsgp error %d: %s" % (code, msg))
        self.code = code


def _p(a):
    return None if a is None else a.ctypes.data_as(_lib.c_double_p)


def _is_plain(a):
    return a is None or (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous)


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        a = a.reshape(shape)
    return a


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            _lib.load().sgp_pinned_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, order="C"):
    """Float64 NumPy array in page-locked host memory (sgp_pinned_alloc): H2D / D2H copies of such arrays run at full PCIe
    speed.  The memory is released when the array (and every view of it) is garbage-collected."""
    n = int(np.prod(shape))
    ptr = ctypes.c_void_p()
    rc = _lib.load().sgp_pinned_alloc(max(n, 1) * 8, ctypes.byref(ptr))
    if rc != 0:
        raise SGPError(rc, "sgp_pinned_alloc failed")
    buf = (ctypes.c_double * max(n, 1)).from_address(ptr.value)
    buf._owner = _PinnedOwner(ptr)          # keeps the allocation alive as long as the ctypes buffer is referenced
    return np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape, order=order)


def pack_lower(A):
    """Lower triangle of a square matrix column by column (LAPACK 'L' packed storage) -- the layout of sgp_sweep_psi_host_packed."""
    A = np.asarray(A)
    M = A.shape[0]
    return np.concatenate([A[j:, j] for j in range(M)]) if M else np.empty(0)


def unpack_lower(packed, M):
    """Full symmetric matrix (column-major) from the packed lower triangle."""
    out = np.empty((M, M), order="F")
    r, c = np.tril_indices(M)
    order = np.lexsort((r, c))                 # column by column
    r, c = r[order], c[order]
    out[r, c] = packed
    out[c, r] = packed
    return out


class SGPContext:
    """Owns one ``sgp_ctx``.  Layout notes: the C ABI takes Julia's column-major D x N; a C-contiguous NumPy array of
    shape (N, D) is the same memory, so inputs here are (N, D) / (M, D) row-per-point arrays.  M x M results come back
    column-major; they are symmetric or explicitly triangular, and are returned as Fortran-ordered views."""

    def __init__(self, device=0):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self.lib.sgp_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise SGPError(rc, "sgp_create failed (no CUDA device / not sm_100?)")
        self.h = h
        self.M = 0
        self.D = 0
        self.N = 0
        self._host_call = None      # cached argument marshalling of sweep_psi_host (same arrays every step)

    def close(self):
        if getattr(self, "h", None):
            self.lib.sgp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise SGPError(rc, self.lib.sgp_last_error(self.h).decode())

    # ---- state ----
    def set_kernel(self, variance, lengthscale, D=None, kind=SE):
        ell = np.atleast_1d(np.asarray(lengthscale, dtype=np.float64))
        D = ell.size if D is None else int(D)
        ell = np.ascontiguousarray(np.broadcast_to(ell, (D,)))
        self._ck(self.lib.sgp_set_kernel(self.h, int(kind), D, float(variance), _p(ell)))
        self.D = D

    def set_inducing(self, Z):
        Z = _f64(Z)
        Z = Z.reshape(-1, self.D)
        self._ck(self.lib.sgp_set_inducing(self.h, Z.shape[0], _p(Z)))
        self.M = Z.shape[0]

    def set_data(self, X, ybar=None, yvar=None, wts=None):
        X = _f64(X).reshape(-1, self.D)
        N = X.shape[0]
        ybar, yvar, wts = _f64(ybar, (N,)), _f64(yvar, (N,)), _f64(wts, (N,))
        self._ck(self.lib.sgp_set_data(self.h, N, _p(X), _p(ybar), _p(yvar), _p(wts)))
        self.N = N

    def set_targets(self, ybar, yvar=None):
        ybar, yvar = _f64(ybar, (self.N,)), _f64(yvar, (self.N,))
        self._ck(self.lib.sgp_set_targets(self.h, _p(ybar), _p(yvar)))

    def set_data_dev(self, N, X_ptr, y_ptr, yv_ptr=None, w_ptr=None):
        self._ck(self.lib.sgp_set_data_dev(self.h, int(N), X_ptr, y_ptr, yv_ptr, w_ptr))
        self.N = int(N)

    # ---- sweep ----
    def sweep_psi(self, fetch=True, out=None):
        """`out` = (psi1, psi2) preallocated arrays (e.g. pinned_empty) to receive the results."""
        M = self.M
        if not fetch:
            self._ck(self.lib.sgp_sweep_psi(self.h, None, None, None, None))
            return None
        psi0, sy2 = ctypes.c_double(), ctypes.c_double()
        if out is not None:
            psi1, psi2 = out
            assert psi1.size == M and psi2.shape == (M, M) and psi2.flags.f_contiguous and psi1.dtype == psi2.dtype == np.float64
        else:
            psi1 = np.empty(M)
            psi2 = np.empty((M, M), order="F")
        self._ck(self.lib.sgp_sweep_psi(self.h, ctypes.byref(psi0), _p(psi1), _p(psi2), ctypes.byref(sy2)))
        return psi0.value, psi1, psi2, sy2.value

    def sweep_psi_host(self, X, ybar=None, yvar=None, wts=None, out=None, packed=False):
        """set_data + sweep_psi in one call (one host synchronisation); `out` = (psi1, psi2) preallocated arrays.
        packed=True (sgp_sweep_psi_host_packed): ALL statistics come back in ONE buffer with ONE device-to-host copy --
        [packed lower triangle of Psi2 (M (M + 1) / 2 doubles, column by column = LAPACK 'L' packed storage; `unpack_lower` expands it) | Psi1 (M) |
        Psi0, sum_y2, sum_w, n] -- half the bytes over the bus; `out` is then that ONE 1-D array of M (M + 1) / 2 + M + 4 doubles (e.g. pinned_empty),
        and the returned psi1 / psi2 are views into it."""
        # Steady-state callers pass the SAME (pinned) arrays every step: the argument marshalling (~20 us of ctypes / numpy work, a tenth of the
        # call at the kin40k shape) is cached on the identity of the array objects -- their buffers cannot move while we hold references.
        c = self._host_call
        o0, o1 = (out, None) if (packed or out is None) else out
        if (c is not None and out is not None and c[0] is X and c[1] is ybar and c[2] is yvar and c[3] is wts and c[4] is o0 and c[5] is o1
                and c[6] == X.shape and c[7] == (self.M, self.D, bool(packed))):
            N, args, psi0, sy2, psi1, psi2 = c[8:14]
        else:
            Xc = _f64(X).reshape(-1, self.D)
            N = Xc.shape[0]; M = self.M
            yb, yv, w = _f64(ybar, (N,)), _f64(yvar, (N,)), _f64(wts, (N,))
            if packed:
                tri = M * (M + 1) // 2
                buf = out if out is not None else np.empty(tri + M + 4)
                assert buf.shape == (tri + M + 4,) and buf.flags.c_contiguous and buf.dtype == np.float64
                psi2, psi1 = buf[:tri], buf[tri:tri + M]
                psi0, sy2 = buf[tri + M:tri + M + 1], buf[tri + M + 1:tri + M + 2]                  # the scalars are read from the buffer's tail
                args = (self.h, N, _p(Xc), _p(yb), _p(yv), _p(w), None, None, _p(buf), None)
            else:
                psi1, psi2 = out if out is not None else (np.empty(M), np.empty((M, M), order="F"))
                assert psi1.size == M and psi2.shape == (M, M) and psi2.flags.f_contiguous and psi1.dtype == psi2.dtype == np.float64
                psi0, sy2 = ctypes.c_double(), ctypes.c_double()
                args = (self.h, N, _p(Xc), _p(yb), _p(yv), _p(w), ctypes.byref(psi0), _p(psi1), _p(psi2), ctypes.byref(sy2))
            cacheable = out is not None and all(_is_plain(a) for a in (X, ybar, yvar, wts))      # (no converted temporaries behind the pointers)
            self._host_call = (X, ybar, yvar, wts, o0, o1, X.shape, (M, self.D, bool(packed)), N, args, psi0, sy2, psi1, psi2) if cacheable else None
        self._ck((self.lib.sgp_sweep_psi_host_packed if packed else self.lib.sgp_sweep_psi_host)(*args))
        self.N = N
        if packed:
            return float(psi0[0]), psi1, psi2, float(sy2[0])
        return psi0.value, psi1, psi2, sy2.value

    def fetch_psi2_packed(self, out=None):
        """Psi2 of the last sweep as its packed lower triangle (sgp_fetch_psi2_packed)."""
        M = self.M
        out = np.empty(M * (M + 1) // 2) if out is None else out
        assert out.shape == (M * (M + 1) // 2,) and out.dtype == np.float64 and out.flags.c_contiguous
        self._ck(self.lib.sgp_fetch_psi2_packed(self.h, _p(out)))
        return out

    def sweep_psi_uncertain(self, method, mean, cov, R=None, D_out=1, p=21, want_psi1_n=False):
        mean = _f64(mean).reshape(-1, self.D)
        N = mean.shape[0]
        cov = None if cov is None else _f64(cov).reshape(N, self.D, self.D)
        M = self.M
        if R is not None:
            R = _f64(R).reshape(N, D_out)
            R = np.ascontiguousarray(R.T).T if False else np.asfortranarray(R)   # N x D_out column-major
        psi0 = ctypes.c_double()
        psi1 = np.empty((M, D_out), order="F")
        psi2 = np.empty((M, M), order="F")
        psi1_n = np.empty((M, N), order="F") if want_psi1_n else None
        self._ck(self.lib.sgp_sweep_psi_uncertain(self.h, int(method), int(p), N, _p(mean), _p(cov), int(D_out), _p(R),
                                                  ctypes.byref(psi0), _p(psi1), _p(psi2), _p(psi1_n)))
        return psi0.value, (psi1[:, 0] if D_out == 1 else psi1), psi2, (None if psi1_n is None else psi1_n.T)

    def sweep_timed(self, reps=1):
        a, b = ctypes.c_float(), ctypes.c_float()
        self._ck(self.lib.sgp_sweep_timed(self.h, int(reps), ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def sweep_timed_flushed(self, reps=1, flush_mb=256, main_kernel=True):
        """Mean device time of one sweep (and of its main kernel) over `reps` back-to-back repetitions, each preceded by an
        L2 flush on the library's stream (outside the timed intervals); no host synchronisation between repetitions.
        main_kernel=False: no inner event pair around the main kernel (the timed interval holds the sweep alone); returns (ms, None)."""
        a, b = ctypes.c_float(), ctypes.c_float()
        self._ck(self.lib.sgp_sweep_timed_flushed(self.h, int(reps), int(flush_mb), ctypes.byref(a), ctypes.byref(b) if main_kernel else None))
        return a.value, (b.value if main_kernel else None)

    def sweep_debug_clocks(self, cap=8192):
        """First call switches per-segment clock recording on; later calls return an (n, 4) int64 array
        {chunks, SM clocks, is_diagonal_tile, cta} of the last sweep (rows of -1 = unused slots)."""
        import numpy as np
        out = np.full((cap, 4), -1, dtype=np.int64)
        n = ctypes.c_int(0)
        self._ck(self.lib.sgp_sweep_debug_clocks(self.h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), cap, ctypes.byref(n)))
        return out[:n.value]

    def last_sweep_info(self):
        v = [ctypes.c_int() for _ in range(4)]
        self._ck(self.lib.sgp_last_sweep_info(self.h, *[ctypes.byref(x) for x in v]))
        return dict(launches=v[0].value, grid=v[1].value, block=v[2].value, smem_bytes=v[3].value)

    # ---- factorisations ----
    def kuu_factor(self, jitter=0.0, fetch=True):
        L = np.empty((self.M, self.M), order="F") if fetch else None
        self._ck(self.lib.sgp_kuu_factor(self.h, float(jitter), _p(L)))
        return L

    def kuu_solve(self, B):
        B = np.asfortranarray(np.array(B, dtype=np.float64).reshape(self.M, -1))
        self._ck(self.lib.sgp_kuu_solve(self.h, B.shape[1], _p(B)))
        return B

    def posterior_v(self, xi0, Lambda0, w, want_Uv=True, out=None):
        """`out` = (mu, Sigma, Uv) preallocated Fortran-ordered arrays (e.g. pinned_empty) to receive the results."""
        M = self.M
        xi0 = _f64(xi0, (M,))
        Lam0 = np.asfortranarray(np.asarray(Lambda0, dtype=np.float64).reshape(M, M))
        if out is not None:
            mu, Sigma, Uv = out
        else:
            mu = np.empty(M)
            Sigma = np.empty((M, M), order="F")
            Uv = np.empty((M, M), order="F") if want_Uv else None
        self._ck(self.lib.sgp_posterior_v(self.h, _p(xi0), _p(Lam0), float(w), _p(mu), _p(Sigma), _p(Uv)))
        return mu, Sigma, Uv

    def prior_set(self, xi0, Lambda0):
        M = self.M
        self._ck(self.lib.sgp_prior_set(self.h, _p(_f64(xi0, (M,))), _p(np.asfortranarray(np.asarray(Lambda0, dtype=np.float64).reshape(M, M)))))

    def prior_set_isotropic(self, variance):
        self._ck(self.lib.sgp_prior_set_isotropic(self.h, float(variance)))

    def posterior_v_stream(self, w, carry=True, fetch=False, out=None, want_Uv=True):
        """Posterior from the RESIDENT prior and the last sweep; with `carry` it becomes the next prior.  fetch=False copies
        nothing back (mu_v / Sigma_v stay resident for w_terms(None, None) / theta_objective(None, None, ...)) and skips the
        Cholesky factor Uv of Sigma_v + mu_v mu_v' altogether."""
        M = self.M
        if out is not None:
            mu, Sigma, Uv = out
        elif fetch:
            mu, Sigma, Uv = np.empty(M), np.empty((M, M), order="F"), (np.empty((M, M), order="F") if want_Uv else None)
        else:
            mu = Sigma = Uv = None
        self._ck(self.lib.sgp_posterior_v_stream(self.h, float(w), 1 if carry else 0, _p(mu), _p(Sigma), _p(Uv)))
        return mu, Sigma, Uv

    def w_terms(self, mu_v, Uv):
        M = self.M
        if mu_v is not None:
            mu_v = _f64(mu_v, (M,))
            Uv = np.asfortranarray(np.asarray(Uv, dtype=np.float64).reshape(M, M))
        a, b = ctypes.c_double(), ctypes.c_double()
        self._ck(self.lib.sgp_w_terms(self.h, _p(mu_v), _p(Uv), ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def predict_mean(self, Xt, mu_v):
        Xt = _f64(Xt).reshape(-1, self.D)
        mu_v = _f64(mu_v, (self.M,))
        out = np.empty(Xt.shape[0])
        self._ck(self.lib.sgp_predict_mean(self.h, Xt.shape[0], _p(Xt), _p(mu_v), _p(out)))
        return out

    def predict_probit(self, Xt, mu_v, w_bar):
        """(mean_f, var_f, prob_y) of `predict_new` in the classification drivers (sgp_predict_probit)."""
        Xt = _f64(Xt).reshape(-1, self.D)
        mu_v = _f64(mu_v, (self.M,))
        mean_f = np.empty(Xt.shape[0]); prob = np.empty(Xt.shape[0]); var_f = ctypes.c_double()
        self._ck(self.lib.sgp_predict_probit(self.h, Xt.shape[0], _p(Xt), _p(mu_v), float(w_bar), _p(mean_f), ctypes.byref(var_f), _p(prob)))
        return mean_f, var_f.value, prob

    def dense_timed(self, what, w=1.0, jitter=0.0, reps=10):
        """Device ms per call of kuu_factor (0) / posterior_v_stream with (1) or without (2) Uv / w_terms (3) on the resident state."""
        ms = ctypes.c_float()
        self._ck(self.lib.sgp_dense_timed(self.h, int(what), float(w), float(jitter), int(reps), ctypes.byref(ms)))
        return ms.value

    def fp64_peak(self, ms_target=200):
        """FP64 DMMA peak of this device, TFLOP/s (sgp_fp64_peak)."""
        v = ctypes.c_double()
        self._ck(self.lib.sgp_fp64_peak(self.h, int(ms_target), ctypes.byref(v)))
        return v.value

    def uncertain_node_terms(self, mu_v, Uv, N):
        """Per-node (Psi0_n, tr(Kuu^-1 Psi2_n), Psi1_n' mu_v, tr(Uv'Uv Psi2_n)) of the last sweep_psi_uncertain's N nodes, plus tr(Kuu^-1) and |Uv|_F^2
        (sgp_uncertain_node_terms)."""
        M = self.M
        mu_v = _f64(mu_v, (M,))
        Uv = np.asfortranarray(np.asarray(Uv, dtype=np.float64).reshape(M, M))
        outs = [np.empty(int(N)) for _ in range(4)]
        a, b = ctypes.c_double(), ctypes.c_double()
        self._ck(self.lib.sgp_uncertain_node_terms(self.h, _p(mu_v), _p(Uv), *[_p(o) for o in outs], ctypes.byref(a), ctypes.byref(b)))
        return outs[0], outs[1], outs[2], outs[3], a.value, b.value

    def in_logmessage(self, Xp, Mv, S, trW, R=None, grad=False, hess=False):
        """log backward message of `@rule MultiSGP(:in)` at P points for each of N nodes (sgp_in_logmessage).
        Xp: (N, P, d); Mv: (M, D_out) columns mu_v^(d); S: (M, M) = sumRvblk_W; R: (N, D_out) rows (W mu_y,n)' or None.
        Returns f (N, P) [, grad (N, P, d)] [, hess (N, P, d, d)].  Needs kuu_factor."""
        Xp = _f64(Xp)
        N, P, d = Xp.shape
        M = self.M
        Mv = np.asfortranarray(np.asarray(Mv, dtype=np.float64).reshape(M, -1))
        Dout = Mv.shape[1]
        S = np.asfortranarray(np.asarray(S, dtype=np.float64).reshape(M, M))
        Rc = None if R is None else _f64(R, (N, Dout))
        f = np.empty((N, P)); g = np.empty((N, P, d)) if (grad or hess) else None; h = np.empty((N, P, d, d)) if hess else None
        self._ck(self.lib.sgp_in_logmessage(self.h, N, P, _p(Xp), Dout, _p(Rc), _p(Mv), _p(S), float(trW), _p(f), _p(g), _p(h)))
        out = (f,) + ((g,) if grad or hess else ()) + ((h,) if hess else ())
        return out if len(out) > 1 else f

    def theta_objective(self, mu_v, Uv, w, jitter=0.0, grad=True):
        """(F, dF/dvariance, dF/dlengthscale[D]) of the theta step on the resident data (sgp_theta_objective)."""
        M = self.M
        if mu_v is not None:
            mu_v = _f64(mu_v, (M,))
            Uv = np.asfortranarray(np.asarray(Uv, dtype=np.float64).reshape(M, M))
        val = ctypes.c_double(); dv = ctypes.c_double(); dl = np.zeros(self.D)
        self._ck(self.lib.sgp_theta_objective(self.h, _p(mu_v), _p(Uv), float(w), float(jitter), ctypes.byref(val),
                                              ctypes.byref(dv) if grad else None, _p(dl) if grad else None))
        return (val.value, dv.value, dl) if grad else val.value

    # ---- multi-GPU ----
    @staticmethod
    def comm_unique_id():
        buf = ctypes.create_string_buffer(128)
        rc = _lib.load().sgp_comm_unique_id(buf)
        if rc != 0:
            raise SGPError(rc, "sgp_comm_unique_id failed (libnccl.so.2 not found?)")
        return buf.raw

    def comm_init(self, nranks, rank, uid):
        self._ck(self.lib.sgp_comm_init(self.h, int(nranks), int(rank), ctypes.create_string_buffer(uid, 128)))
