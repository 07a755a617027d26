"""ctypes face of the CPU ORACLE (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: a CPU restatement of the hot path of afranson/Lisp-MCMC
(/root/reference/mcmc-fitting.lisp; every function in orc_walker.cpp / orc_sexpr.cpp
cites the reference lines it follows).  Only tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py may import this module.  The product
package lisp_mcmc_b200 never does.

Parity pinning: the reference cannot run here (no Common Lisp in the image), so the
oracle is pinned on the reference's only known-answer vectors (example-covariance /
example-l-matrix, mcmc-fitting.lisp:729-751) and closed forms; for log-likelihood values
and accept/reject sequences the reference holds no golden vector ("parity unpinned"
beyond the restatement -- see DESIGN.md).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FAM_POLY, FAM_EXPO, FAM_SINUS, FAM_LORENTZ, FAM_GAUSS_LIN, FAM_NV_MULTI, FAM_SEXPR = 0, 1, 2, 3, 4, 5, 100
LL_NORMAL, LL_NORMAL_CUTOFF, LL_POISSON, LL_USER = 0, 1, 2, 3
AUTO_NONE, AUTO_PROB_SETTLE = 0, 1
OPT_COVARIANCE, OPT_BEST_VALUE = 0, 1
SUM_REFERENCE, SUM_COMPENSATED = 0, 1

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_long)


class AdaptCfg(C.Structure):
    _fields_ = [("n", C.c_long), ("temperature", C.c_double), ("auto_mode", C.c_int),
                ("sampling_optimization", C.c_int), ("max_walker_length", C.c_long), ("greedy", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in ("orc_sexpr.cpp", "orc_walker.cpp", "orc_internal.h", "mfit_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_log_normal.restype = C.c_double
        L.orc_log_normal.argtypes = [C.c_double] * 3
        L.orc_log_factorial.restype = C.c_double
        L.orc_log_factorial.argtypes = [C.c_double]
        L.orc_log_poisson.restype = C.c_double
        L.orc_log_poisson.argtypes = [C.c_double] * 2
        L.orc_lplist_covariance.argtypes = [_dp, C.c_int, C.c_int, _dp]
        L.orc_cholesky_decomp.argtypes = [_dp, C.c_int, _dp]
        L.orc_scale_array.argtypes = [C.c_double, _dp, C.c_int]
        L.orc_get_covariant_sample.argtypes = [_dp, _dp, C.c_int, _dp, _dp]
        L.orc_lplist_to_l_matrix.argtypes = [_dp, C.c_int, C.c_int, _dp]
        L.orc_model_builtin.restype = C.c_double
        L.orc_model_builtin.argtypes = [C.c_int, C.c_int, _dp, C.c_double]
        L.orc_fn_parse.restype = C.c_void_p
        L.orc_fn_parse.argtypes = [C.c_char_p, C.POINTER(C.c_char_p), C.c_int, C.c_char_p, C.c_int]
        L.orc_fn_free.argtypes = [C.c_void_p]
        L.orc_fn_call_model.restype = C.c_double
        L.orc_fn_call_model.argtypes = [C.c_void_p, C.c_double, _dp]
        L.orc_fn_call_model_v.restype = C.c_double
        L.orc_fn_call_model_v.argtypes = [C.c_void_p, _dp, C.c_int, _dp]
        L.orc_walker_dataset_xd.argtypes = [C.c_void_p, C.c_int, _dp, C.c_int, _dp, _dp, C.c_long]
        L.orc_fn_call_prior.restype = C.c_double
        L.orc_fn_call_prior.argtypes = [C.c_void_p, _dp]
        L.orc_temp_steps.restype = C.c_long
        L.orc_temp_steps.argtypes = [C.c_long, C.c_int]
        L.orc_temperature_schedule.argtypes = [C.c_long, C.c_int, C.c_double, _dp]
        L.orc_walker_new.restype = C.c_void_p
        L.orc_walker_new.argtypes = [C.c_int, C.c_int]
        L.orc_walker_free.argtypes = [C.c_void_p]
        L.orc_walker_dataset.argtypes = [C.c_void_p, C.c_int, _dp, _dp, _dp, C.c_long]
        L.orc_walker_model_builtin.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip, C.c_int]
        L.orc_walker_model_sexpr.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_char_p)]
        L.orc_walker_likelihood.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_walker_likelihood_sexpr.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
        L.orc_walker_prior_flat.argtypes = [C.c_void_p, C.c_int]
        L.orc_walker_prior_bounds.argtypes = [C.c_void_p, C.c_int, C.c_int, _ip, _dp, _dp]
        L.orc_walker_prior_sexpr.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.POINTER(C.c_char_p)]
        L.orc_walker_error.restype = C.c_char_p
        L.orc_walker_error.argtypes = [C.c_void_p]
        L.orc_walker_logp.argtypes = [C.c_void_p, _dp, C.c_int, _dp, _dp, _dp]
        L.orc_walker_init.argtypes = [C.c_void_p, _dp]
        L.orc_walker_adaptive_steps_full.argtypes = [C.c_void_p, C.POINTER(AdaptCfg), _dp, _dp, _dp, C.c_long, _lp]
        L.orc_walker_adaptive_steps_limited.argtypes = [C.c_void_p, C.POINTER(AdaptCfg), _dp, _dp, _dp, C.c_long, C.c_long, _lp]
        L.orc_walker_many_steps.argtypes = [C.c_void_p, C.c_long, _dp, _dp, _dp, C.c_long]
        L.orc_walker_sample_region.argtypes = [C.c_void_p, C.c_double, _dp, C.c_long]
        L.orc_walker_force_take_step.argtypes = [C.c_void_p]
        L.orc_walker_seed.argtypes = [C.c_void_p, C.c_ulonglong]
        L.orc_walker_length.restype = C.c_long
        L.orc_walker_length.argtypes = [C.c_void_p]
        L.orc_walker_age.restype = C.c_long
        L.orc_walker_age.argtypes = [C.c_void_p]
        L.orc_walker_trace.restype = C.c_long
        L.orc_walker_trace.argtypes = [C.c_void_p, C.c_long, _dp, _dp]
        L.orc_walker_most_likely.argtypes = [C.c_void_p, _dp, _dp]
        L.orc_walker_last_l_matrix.argtypes = [C.c_void_p, _dp]
        L.orc_walker_acceptance.argtypes = [C.c_void_p, C.c_long, _lp, _lp]
        L.orc_walker_keep_walks.argtypes = [C.c_void_p, C.c_long]
        L.orc_walker_modify.argtypes = [C.c_void_p, C.c_int, C.c_long]
        L.orc_walker_load_walk.argtypes = [C.c_void_p, _dp, _dp, C.c_long]
        L.orc_walker_data_and_fit.argtypes = [C.c_void_p, C.c_long, C.c_int, C.c_int, _dp, C.c_long, _dp, _dp, _dp, _dp]
        L.orc_bench_many.restype = C.c_double
        L.orc_bench_many.argtypes = [C.c_void_p, _dp, C.c_long, C.POINTER(AdaptCfg), C.c_int, C.c_ulonglong, _dp, C.c_long]
        _LIB = L
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def _keys(keys):
    arr = (C.c_char_p * len(keys))(*[k.encode() for k in keys])
    return arr


def log_normal(x, mu, sigma):
    return lib().orc_log_normal(x, mu, sigma)


def log_factorial(n):
    return lib().orc_log_factorial(float(n))


def log_poisson(lam, k):
    return lib().orc_log_poisson(lam, k)


def lplist_covariance(rows):
    rows, p = _d(rows)
    n, P = rows.shape
    out = np.zeros((P, P))
    lib().orc_lplist_covariance(p, n, P, out.ctypes.data_as(_dp))
    return out


def cholesky_decomp(cov):
    cov, p = _d(cov)
    P = cov.shape[0]
    out = np.zeros((P, P))
    lib().orc_cholesky_decomp(p, P, out.ctypes.data_as(_dp))
    return out


def get_covariant_sample(means, L, z):
    means, pm = _d(means)
    L, pl = _d(L)
    z, pz = _d(z)
    out = np.zeros(len(means))
    lib().orc_get_covariant_sample(pm, pl, len(means), pz, out.ctypes.data_as(_dp))
    return out


def lplist_to_l_matrix(rows):
    rows, p = _d(rows)
    n, P = rows.shape
    out = np.zeros((P, P))
    lib().orc_lplist_to_l_matrix(p, n, P, out.ctypes.data_as(_dp))
    return out


def model_builtin(family, degree, theta, x):
    theta, p = _d(theta)
    return lib().orc_model_builtin(family, degree, p, float(x))


def linspace(start, end, length=50):
    """linspace :235-247 with :len: exact rationals start + n (end - start)/(len - 1) -- the difference formed in double
    arithmetic first -- each coerced to double-float (one rounding per element)."""
    from fractions import Fraction
    step = Fraction(float(end) - float(start)) / (length - 1)
    s0 = Fraction(float(start))
    return np.array([float(s0 + n * step) for n in range(length)], dtype=np.float64)


def temperature_schedule(n, P, temperature):
    Tn = lib().orc_temp_steps(n, P)
    out = np.zeros(Tn)
    lib().orc_temperature_schedule(n, P, float(temperature), out.ctypes.data_as(_dp))
    return out


class Fn:
    """A parsed :function or :log-prior lambda."""

    def __init__(self, text, keys):
        err = C.create_string_buffer(512)
        self._keys = _keys(keys)
        self.P = len(keys)
        self.h = lib().orc_fn_parse(text.encode(), self._keys, len(keys), err, 512)
        if not self.h:
            raise ValueError("oracle s-expression: " + err.value.decode())

    def model(self, x, theta):
        theta, p = _d(theta)
        return lib().orc_fn_call_model(self.h, float(x), p)

    def model_v(self, x, theta):
        """(apply fn x params) with x a sequence of numbers"""
        theta, p = _d(theta)
        x, px = _d(x)
        return lib().orc_fn_call_model_v(self.h, px, len(x), p)

    def prior(self, theta):
        theta, p = _d(theta)
        return lib().orc_fn_call_prior(self.h, p)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fn_free(self.h)
            self.h = None


class Walker:
    """One reference walker (struct walker, mcmc-fitting.lisp:467-479) over D datasets and P parameters."""

    def __init__(self, keys, n_datasets=1):
        self.keys = list(keys)
        self.P = len(self.keys)
        self.D = n_datasets
        self._ck = _keys(self.keys)
        self.h = lib().orc_walker_new(self.D, self.P)

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: rc=%d %s" % (rc, lib().orc_walker_error(self.h).decode()))

    def dataset(self, d, x, y, sigma=None):
        if np.ndim(x) == 2:               # every x is a sequence: x[i] = (x1 x2 ...)
            return self._dataset_xd(d, x, y, sigma)
        x, px = _d(x)
        y, py = _d(y)
        if sigma is None:
            ps = None
        else:
            sigma = np.broadcast_to(np.asarray(sigma, dtype=np.float64), x.shape)
            sigma, ps = _d(sigma)
        self._chk(lib().orc_walker_dataset(self.h, d, px, py, ps, len(x)))
        if not hasattr(self, "_x"):
            self._x = {}
        self._x[d] = x

    def _dataset_xd(self, d, x, y, sigma):
        x, px = _d(x)
        y, py = _d(y)
        ps = None
        if sigma is not None:
            sigma, ps = _d(np.broadcast_to(np.asarray(sigma, dtype=np.float64), y.shape))
        self._chk(lib().orc_walker_dataset_xd(self.h, d, px, x.shape[1], py, ps, len(y)))

    def model_builtin(self, d, family, idx, degree=0):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self._chk(lib().orc_walker_model_builtin(self.h, d, family, degree, idx.ctypes.data_as(_ip), len(idx)))

    def model_sexpr(self, d, text):
        self._chk(lib().orc_walker_model_sexpr(self.h, d, text.encode(), self._ck))

    def likelihood(self, d, kind):
        self._chk(lib().orc_walker_likelihood(self.h, d, kind))

    def likelihood_sexpr(self, d, text):
        """:log-liklihood = (create-log-liklihood-function <text>), text = "(lambda (y model error) ...)" (:402-417)"""
        self._chk(lib().orc_walker_likelihood_sexpr(self.h, d, text.encode()))

    def prior_flat(self, d):
        self._chk(lib().orc_walker_prior_flat(self.h, d))

    def prior_bounds(self, d, idx, lo, hi):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        lo, pl = _d(lo)
        hi, ph = _d(hi)
        self._chk(lib().orc_walker_prior_bounds(self.h, d, len(idx), idx.ctypes.data_as(_ip), pl, ph))

    def prior_sexpr(self, d, text):
        self._chk(lib().orc_walker_prior_sexpr(self.h, d, text.encode(), self._ck))

    def logp(self, theta, sum_mode=SUM_REFERENCE):
        theta, p = _d(theta)
        prob = C.c_double()
        ll = np.zeros(self.D)
        lp = np.zeros(self.D)
        self._chk(lib().orc_walker_logp(self.h, p, sum_mode, C.byref(prob), ll.ctypes.data_as(_dp),
                                        lp.ctypes.data_as(_dp)))
        return prob.value, ll, lp

    def init(self, theta0):
        theta0, p = _d(theta0)
        self._chk(lib().orc_walker_init(self.h, p))

    def seed(self, s):
        lib().orc_walker_seed(self.h, s)

    def adaptive_steps_full(self, n=100000, temperature=1e3, auto=AUTO_NONE, sampling_optimization=OPT_COVARIANCE,
                            max_walker_length=0, l_matrix=None, z=None, u=None, greedy=False, max_steps=-1):
        cfg = AdaptCfg(int(n), float(temperature), auto, sampling_optimization, int(max_walker_length or 0),
                       int(bool(greedy)))
        pl = pz = pu = None
        nstream = 0
        if l_matrix is not None:
            l_matrix, pl = _d(l_matrix)
        if z is not None:
            z, pz = _d(z)
            nstream = z.shape[0]
        if u is not None:
            u, pu = _d(u)
        taken = C.c_long()
        self._chk(lib().orc_walker_adaptive_steps_limited(self.h, C.byref(cfg), pl, pz, pu, nstream, int(max_steps),
                                                          C.byref(taken)))
        return taken.value

    def adaptive_steps(self, n=30000, **kw):
        """walker-adaptive-steps, mcmc-fitting.lisp:946-947"""
        return self.adaptive_steps_full(n=n, temperature=10, auto=AUTO_PROB_SETTLE, **kw)

    def many_steps(self, n, l_matrix, z=None, u=None):
        l_matrix, pl = _d(l_matrix)
        pz = pu = None
        nstream = 0
        if z is not None:
            z, pz = _d(z)
            nstream = z.shape[0]
        if u is not None:
            u, pu = _d(u)
        self._chk(lib().orc_walker_many_steps(self.h, n, pl, pz, pu, nstream))

    def sample_region(self, initial_scale=1e-3, z=None):
        pz = None
        nstream = 0
        if z is not None:
            z, pz = _d(z)
            nstream = z.shape[0]
        self._chk(lib().orc_walker_sample_region(self.h, initial_scale, pz, nstream))

    def force_take_step(self):
        self._chk(lib().orc_walker_force_take_step(self.h))

    @property
    def length(self):
        return lib().orc_walker_length(self.h)

    @property
    def age(self):
        return lib().orc_walker_age(self.h)

    def trace(self, take=0):
        """(theta[take,P], prob[take]) newest first, like walker-get :steps :take."""
        n = self.length if not take else min(take, self.length)
        th = np.zeros((n, self.P))
        pr = np.zeros(n)
        r = lib().orc_walker_trace(self.h, take, th.ctypes.data_as(_dp), pr.ctypes.data_as(_dp))
        return th[:r], pr[:r]

    def most_likely(self):
        th = np.zeros(self.P)
        pr = C.c_double()
        lib().orc_walker_most_likely(self.h, th.ctypes.data_as(_dp), C.byref(pr))
        return th, pr.value

    def last_l_matrix(self):
        L = np.zeros((self.P, self.P))
        lib().orc_walker_last_l_matrix(self.h, L.ctypes.data_as(_dp))
        return L

    def acceptance(self, take):
        c, n = C.c_long(), C.c_long()
        lib().orc_walker_acceptance(self.h, take, C.byref(c), C.byref(n))
        return c.value, n.value

    def keep_walks(self, keep):
        lib().orc_walker_keep_walks(self.h, keep)

    def modify(self, modify, number=0):
        """walker-modify :burn-walks / :keep-walks / :reset / :reset-to-most-likely"""
        mode = {"burn-walks": 1, "keep-walks": 2, "reset": 3, "reset-to-most-likely": 4}[str(modify).lstrip(":")]
        self._chk(lib().orc_walker_modify(self.h, mode, int(number)))

    def load_walk(self, theta, prob):
        """replace the walk by (theta[n,P], prob[n]), newest first"""
        theta, pt = _d(theta)
        prob, pp = _d(prob)
        self._chk(lib().orc_walker_load_walk(self.h, pt, pp, len(prob)))

    def data_and_fit(self, take=1000, fn_number=0, which_solution="most-likely", n_fit=1000, x_shift=None, y_shift=None):
        """walker-get-data-and-fit :1231-1255 -> (x_fit, max_ys, min_ys, y_fit, params); x-data / y-data are the
        caller's own arrays and are not repeated here."""
        x = self._x[fn_number]
        xf = linspace(float(np.min(x)), float(np.max(x)), n_fit)
        yf, mx, mn, par = np.zeros(n_fit), np.zeros(n_fit), np.zeros(n_fit), np.zeros(self.P)
        which = 0 if str(which_solution).lstrip(":") == "most-likely" else 1
        self._chk(lib().orc_walker_data_and_fit(self.h, int(take or 0), fn_number, which, xf.ctypes.data_as(_dp), n_fit,
                                                yf.ctypes.data_as(_dp), mx.ctypes.data_as(_dp), mn.ctypes.data_as(_dp),
                                                par.ctypes.data_as(_dp)))
        xs = xf if x_shift is None else xf + x_shift
        ys = 0.0 if y_shift is None else y_shift
        return xs, mx + ys, mn + ys, yf + ys, par

    def bench_many(self, theta0, n, n_threads, seed=1, temperature=10.0, max_steps=-1):
        """CPU baseline: W independent copies of this walker run walker-adaptive-steps-full :n n (interrupted after
        max_steps take-step calls when >= 0), chains split across threads; returns (seconds, last probs)."""
        theta0, p = _d(theta0)
        W = theta0.shape[0]
        cfg = AdaptCfg(int(n), float(temperature), AUTO_NONE, OPT_COVARIANCE, 0, 0)
        last = np.zeros(W)
        secs = lib().orc_bench_many(self.h, p, W, C.byref(cfg), n_threads, seed, last.ctypes.data_as(_dp), int(max_steps))
        return secs, last

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_walker_free(self.h)
            self.h = None
