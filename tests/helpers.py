"""Shared by the parity tests: build the ORACLE walker that corresponds to a GPU workload dict."""
import numpy as np

from lisp_mcmc_b200.walker import Builtin, clean_data, clean_data_error

_LL = {"log-liklihood-normal": 0, "log-liklihood-normal-cutoff": 1, "log-poisson": 2}


def oracle_walker(orc, cfg, use_builtin_formula=False):
    """The reference-side walker for `cfg`: every :function goes through the oracle's generic s-expression path
    (apply fn x params), i.e. the family's canonical Lisp text -- never through GPU code."""
    fns = cfg["function"]
    D = len(fns)
    ow = orc.Walker(cfg["keys"], D)
    seq_x = np.ndim(cfg["data"][0][0]) == 2   # x made of sequences (:1137): only the list-of-sets form is accepted
    data = clean_data(cfg["data"] if D > 1 or seq_x else cfg["data"][0], D)
    err = clean_data_error(cfg["data_error"], data)
    lls = cfg["log_liklihood"] if isinstance(cfg["log_liklihood"], (list, tuple)) else [cfg["log_liklihood"]] * D
    lps = cfg["log_prior"]
    if not (isinstance(lps, (list, tuple)) and len(lps) == D and not (lps and isinstance(lps[0], tuple))):
        lps = [lps] * D
    for d in range(D):
        x, y = data[d]
        ow.dataset(d, x, y, err[d])
        f = fns[d]
        if isinstance(f, Builtin):
            if use_builtin_formula:
                ow.model_builtin(d, {"poly": 0, "expo": 1, "sinus": 2, "lorentz": 3, "gauss_lin": 4, "nv_multi": 5}[f.name],
                                 [cfg["keys"].index(k) for k in f.keys], f.degree)
            else:
                ow.model_sexpr(d, f.lisp_text())
        else:
            ow.model_sexpr(d, f)
        if isinstance(lls[d], str) and lls[d].lstrip().startswith("("):
            ow.likelihood_sexpr(d, lls[d])      # (create-log-liklihood-function (lambda (y model error) ...)) :402-417
        else:
            ow.likelihood(d, _LL[lls[d]])
        p = lps[d]
        if p in (None, "log-prior-flat"):
            ow.prior_flat(d)
        elif isinstance(p, str):
            ow.prior_sexpr(d, p)
        else:
            ow.prior_bounds(d, [cfg["keys"].index(k) if k in cfg["keys"] else -1 for k, _, _ in p],
                            [b[1] for b in p], [b[2] for b in p])
    return ow


def rel_err(a, b, scale=0.0):
    """max |a - b| / max(|b|, scale).  `scale` is the size of the terms that were summed (see ll_scale): a sum that
    cancels to nearly zero is only defined to 1e-12 of what went into it -- the reference's own sequential sum
    carries ~sqrt(N) eps of that."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.maximum(np.abs(b), scale), 1e-300))


def ll_scale(cfg):
    """sum_i |-ln(2 pi)/2 - ln sigma_i| over all data sets: the theta-independent half of log-liklihood-normal
    (mcmc-fitting.lisp:377), against which -sum r^2/2 cancels near the best fit."""
    D = len(cfg["function"])
    seq_x = np.ndim(cfg["data"][0][0]) == 2   # x made of sequences (:1137): only the list-of-sets form is accepted
    data = clean_data(cfg["data"] if D > 1 or seq_x else cfg["data"][0], D)
    err = clean_data_error(cfg["data_error"], data)
    return float(sum(np.sum(np.abs(-0.5 * np.log(2 * np.pi) - np.log(e))) for e in err))
