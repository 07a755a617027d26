"""Built-in model families.

The reference ships none (README.md:11 tells users to write their own; SURVEY.md 8(d) defines these).
Each family has (i) a hand-written FP64 device struct in csrc/mfit_models.cuh and (ii) the canonical
Lisp lambda below -- what a user of the reference would pass as :function.  Parity tests push (ii)
through the oracle's generic (apply fn x params) path (mcmc-fitting.lisp:400) and through the NVRTC
lowering, and compare both with (i)."""
from . import _lib

# name -> (family id, theta slot names)
_FIXED = {
    "expo": (_lib.FAM_EXPO, ["a", "b", "c"]),
    "sinus": (_lib.FAM_SINUS, ["a", "omega", "phi", "c"]),
    "lorentz": (_lib.FAM_LORENTZ, ["a", "x0", "gamma", "c"]),
    "gauss_lin": (_lib.FAM_GAUSS_LIN, ["a", "mu", "s", "b0", "b1"]),
}


def slots(name, degree=0):
    """theta slot names of a family, in the order mfit_model_builtin expects them"""
    if name == "poly":
        return ["c%d" % k for k in range(degree + 1)]
    if name == "nv_multi":
        return ["scale%d" % (k + 1) for k in range(degree)] + ["mu%d" % (k + 1) for k in range(degree)] + ["sigma", "bg0"]
    return list(_FIXED[name][1])


def family_id(name):
    if name == "poly":
        return _lib.FAM_POLY
    if name == "nv_multi":
        return _lib.FAM_NV_MULTI
    return _FIXED[name][0]


def lisp_text(name, degree=0, keys=None):
    """The family as a reference-style :function lambda over the given key names (default: the slot names)."""
    k = list(keys) if keys is not None else slots(name, degree)
    args = " ".join(k)
    if name == "poly":
        body = k[degree]
        for j in range(degree - 1, -1, -1):
            body = "(+ %s (* x %s))" % (k[j], body)
    elif name == "expo":
        body = "(+ (* %s (exp (* %s x))) %s)" % (k[0], k[1], k[2])
    elif name == "sinus":
        body = "(+ (* %s (sin (+ (* %s x) %s))) %s)" % (k[0], k[1], k[2], k[3])
    elif name == "lorentz":
        body = "(+ (/ (* %s (* %s %s)) (+ (* (- x %s) (- x %s)) (* %s %s))) %s)" % (k[0], k[2], k[2], k[1], k[1], k[2], k[2], k[3])
    elif name == "gauss_lin":
        body = "(+ (* %s (exp (* -1/2 (expt (/ (- x %s) %s) 2)))) %s (* %s x))" % (k[0], k[1], k[2], k[3], k[4])
    elif name == "nv_multi":
        K = degree
        g = k[2 * K]
        terms = " ".join("(/ (* %s (* %s %s)) (+ (* (- x %s) (- x %s)) (* %s %s)))" % (k[j], g, g, k[K + j], k[K + j], g, g)
                         for j in range(K))
        body = "(- %s %s)" % (k[2 * K + 1], terms)
    else:
        raise KeyError(name)
    return "(lambda (x &key %s &allow-other-keys) %s)" % (args, body)


def log_prior_nv_text(suffix=""):
    """The NV prior of nv-specific.lisp:21-34 (six bounds through prior-bounds-let plus three -1e9 ordering /
    separation / amplitude-ratio constraints, single-float literals as in the reference) over keys that may
    carry a per-dataset suffix for global fits (scale1-3, mu1-3, ...); `sigma` is the shared linewidth."""
    k = {n: n + suffix for n in ("scale1", "scale2", "mu1", "mu2", "bg0")}
    k["sigma"] = "sigma"
    return ("(lambda (params data) (declare (ignorable data)) "
            "(prior-bounds-let ((:%(scale1)s 1d-5 1d1) (:%(scale2)s 1d-5 1d1) (:%(mu1)s 2850 2870) (:%(mu2)s 2870 2890) "
            "(:%(sigma)s 9 20) (:%(bg0)s 0 1d-5)) "
            "(+ bounds-total (if (> %(mu1)s %(mu2)s) -1e9 0e0) (if (< (- %(mu2)s %(mu1)s) 6) -1e9 0e0) "
            "(if (not (< 0.9 (/ %(scale1)s %(scale2)s) 1.1)) -1e9 0e0))))") % k


LOG_PRIOR_NV = log_prior_nv_text()
