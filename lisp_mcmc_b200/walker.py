"""Host-side mirror of the reference's walker API for the GPU path.

Same names, argument meaning and error behaviour as mcmc-fitting.lisp: walker-create (:1132-1163),
walker-adaptive-steps(-full) (:862-947), walker-many-steps (:849-853), walker-sample-region (:949-969),
walker-force-take-step (:1124-1129), walker-get (:487-543), walker-modify :keep-walks (:568-569),
walker-save / walker-load (:971-1001, dead code in the reference), mcmc-fit (:1165-1175).  The SBCL
version of this file is lisp/mfit-cuda.lisp; both sit on the same C ABI (include/mfit_cuda.h).

One `Walker` here is a *walker set* in the reference's vocabulary (nv-specific.lisp:58-66): W independent
chains sharing :function / :data / :log-liklihood / :log-prior, advanced together on the GPU.  W = 1 is
exactly one reference walker.

Plugin slot values are descriptors instead of closures (nothing calls back into the host):
  :function       `builtin("gauss_lin", keys)` | Lisp lambda text (lowered to CUDA C, compiled by NVRTC)
  :log-liklihood  "log-liklihood-normal" | "log-liklihood-normal-cutoff" | "log-poisson"
                  | create_log_liklihood_function("(lambda (y model error) ...)") -- any per-point Lisp lambda (lowered, NVRTC)
  :log-prior      "log-prior-flat" | [(key, lo, hi), ...] (prior-bounds-let with body bounds-total)
                  | Lisp lambda text `(lambda (params data) ...)`
"""
import ctypes as C
import os
import struct

import numpy as np

from . import _lib, families

_dp = C.POINTER(C.c_double)


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Builtin:
    """:function slot value selecting a hand-written family; `keys` maps the family's theta slots to parameter keys."""

    def __init__(self, name, keys=None, degree=0):
        self.name, self.degree = name, int(degree)
        self.keys = [k.lstrip(":").lower() for k in (keys if keys is not None else families.slots(name, degree))]
        if len(self.keys) != len(families.slots(name, self.degree)):
            raise ValueError("family %s takes %d parameters" % (name, len(families.slots(name, self.degree))))

    def lisp_text(self):
        return families.lisp_text(self.name, self.degree, self.keys)

    def __repr__(self):
        return "(builtin %s %d (%s))" % (self.name, self.degree, " ".join(self.keys))


def builtin(name, keys=None, degree=0):
    return Builtin(name, keys, degree)


_LL = {"log-liklihood-normal": _lib.LL_NORMAL, "log-liklihood-normal-cutoff": _lib.LL_NORMAL_CUTOFF,
       "log-poisson": _lib.LL_POISSON, None: _lib.LL_NORMAL}
_AUTO = {None: _lib.AUTO_NONE, False: _lib.AUTO_NONE, "prob-settle": _lib.AUTO_PROB_SETTLE, ":prob-settle": _lib.AUTO_PROB_SETTLE}
_OPT = {None: _lib.OPT_COVARIANCE, "covariance": _lib.OPT_COVARIANCE, ":covariance": _lib.OPT_COVARIANCE,
        "best-value": _lib.OPT_BEST_VALUE, ":best-value": _lib.OPT_BEST_VALUE}


def linspace(start, end, length=50):
    """linspace :235-247 with :len -- exact rationals start + n (end - start)/(len - 1), the difference taken in double
    arithmetic first, every element rounded to a double once (numpy.linspace rounds differently)."""
    from fractions import Fraction
    step = Fraction(float(end) - float(start)) / (length - 1)
    first = Fraction(float(start))
    return np.array([float(first + n * step) for n in range(length)], dtype=np.float64)


# ------------------------------------------------------------------ input normalisation (walker-create :1142-1147)
def force_list(item):
    """force-list :755-759"""
    return list(item) if isinstance(item, (list, tuple)) else [item]


def get_depth(tree):
    """get-depth :761-771: nesting depth of a list tree (a flat list is 1)"""
    if isinstance(tree, np.ndarray):
        return tree.ndim
    if not isinstance(tree, (list, tuple)):
        return 0
    return 1 + (max([get_depth(t) for t in tree]) if len(tree) else 0)


def clean_data(data, n_functions):
    """clean-data :807-825: -> list of n_functions datasets, each (x, y) as float64 vectors.
    Accepts one dataset `(x y)` for a single function or a list of datasets; raises like the reference's
    `error` when the shape does not match the number of functions."""
    depth = get_depth(data)
    if depth == 2:
        sets = [data]                                  # :818
    elif depth >= 3:
        sets = list(data)                              # :820 (depth 4: sets whose x are sequences, :1137)
    else:
        raise ValueError("data must be (x y) or a list of (x y) sets, got depth %d" % depth)
    if len(sets) != n_functions:
        raise ValueError("number of data sets (%d) does not match number of functions (%d)" % (len(sets), n_functions))
    out = []
    for s in sets:
        if len(s) < 2:
            raise ValueError("each data set needs x and y")
        x, y = _f64(s[0]), _f64(s[1])
        # multi-dimensional inputs (:1137): every x may itself be a sequence (x1 x2 ...) -> x is n x K
        if x.ndim not in (1, 2) or y.ndim != 1 or x.shape[0] != y.shape[0]:
            raise ValueError("x and y must be of equal length (x: numbers, or sequences of equal length)")
        out.append((x, y))
    return out


def clean_data_error(err, data):
    """clean-data-error :774-805.  The stddev tree is kept when its structure equals the structure of the y data
    (one full list per data set); otherwise each data set takes the stddev list if it is flat and as long as that
    data set's y, and in every other case all its points get the FIRST number found in stddev.  nil means 1 (:1144)."""
    if err is None:
        err = 1.0

    def first_element(t):
        while isinstance(t, (list, tuple, np.ndarray)):
            if len(t) == 0:
                return None
            t = t[0]
        return float(t)

    def flat_numbers(t):
        return isinstance(t, np.ndarray) and t.ndim == 1 or \
            (isinstance(t, (list, tuple)) and all(np.ndim(v) == 0 for v in t))

    default = first_element(err)
    ys = [y for _, y in data]
    if isinstance(err, (list, tuple)) and len(err) == len(ys) and \
            all(not np.ndim(e) == 0 and flat_numbers(e) and len(e) == len(y) for e, y in zip(err, ys)):
        return [_f64(e) for e in err]                      # eq-structure with the whole y tree (:800)
    out = []
    for y in ys:
        if not np.ndim(err) == 0 and flat_numbers(err) and len(err) == len(y):
            out.append(_f64(err))                          # eq-structure with this data set's y
        else:
            out.append(np.full(y.shape, default))          # every number replaced by the first stddev
    return out


def plist(params):
    """(:a 1 :b 2) / dict / [(k, v)] -> (keys, values) with to-double-floats :833-835 applied"""
    if isinstance(params, dict):
        items = list(params.items())
    else:
        p = list(params)
        if p and isinstance(p[0], (tuple, list)):
            items = [(k, v) for k, v in p]
        else:
            if len(p) % 2:
                raise ValueError("params must be a property list")
            items = [(p[i], p[i + 1]) for i in range(0, len(p), 2)]
    keys = [str(k).lstrip(":").lower() for k, _ in items]
    if len(set(keys)) != len(keys):
        raise ValueError("duplicate parameter keys")
    if len(items) == 1 and np.ndim(items[0][1]) >= 1:
        # :single-item style (:8-14; walker-create :1153-1155 looks at (elt params 1)): ONE key whose value is a list, a
        # vector or an n x 1 array.  Flattened for the device: element k becomes parameter NAME[k], which the lowering
        # resolves from (elt NAME k) / (aref NAME k) / (aref NAME k 0) / (nth k NAME) ... in the :function text.
        v = np.asarray(items[0][1], dtype=np.float64)
        if v.ndim == 2 and v.shape[1] == 1:
            v = v[:, 0]
        if v.ndim != 1 or v.size == 0:
            raise ValueError("a single-item :params value must be a list, a vector or an n x 1 array of numbers")
        return ["%s[%d]" % (keys[0], k) for k in range(v.size)], v.copy()
    return keys, np.array([float(v) for _, v in items], dtype=np.float64)


def create_log_liklihood_function(log_liklihood_function):
    """create-log-liklihood-function :402-417: wraps a per-point (lambda (y model error) ...) into a data-set likelihood.
    Here the lambda is Lisp text and so is the result (a "(create-log-liklihood-function <lambda>)" wrapper is unwrapped):
    pass it as :log-liklihood.  `error` is the data set's whole stddev list, as in the reference (:416)."""
    t = log_liklihood_function.strip()
    low = t.lower()
    if low.startswith("(create-log-liklihood-function"):
        t = t[len("(create-log-liklihood-function"):].rstrip()
        if not t.endswith(")"):
            raise ValueError("unbalanced create-log-liklihood-function form")
        t = t[:-1].strip()
    if not t.lower().startswith("(lambda"):
        raise ValueError("create-log-liklihood-function takes a (lambda (y model error) ...)")
    return t


# ------------------------------------------------------------------ sharding plans (one process per GPU)
def shard_chains(W, rank, world):
    """Contiguous, balanced partition of W chains over `world` ranks -> (first global chain id, count)."""
    base, rem = divmod(W, world)
    lo = rank * base + min(rank, rem)
    return lo, base + (1 if rank < rem else 0)


def shard_datasets(D, world):
    """owner[d] for D datasets over `world` ranks, round-robin (C4: dataset d on GPU d)."""
    return [d % world for d in range(D)]


class _MultiABI:
    """The single-GPU entry points a Walker calls, answered by their mfit_multi_* twins (csrc/multi.cpp): the handle is an
    mfit_multi*, the walker set is sharded over the GPUs of this box inside the library (one host thread per GPU), per-chain
    results come back in global chain order.  Calls without a twin (sample-region, force-take-step, injected streams) raise."""

    def __init__(self, L):
        self._lib = L

    def __getattr__(self, name):
        L = self._lib
        if name == "mfit_ctx_destroy":
            return L.mfit_multi_destroy
        if name == "mfit_last_error":
            return L.mfit_multi_last_error
        if name in ("mfit_dataset_upload",):
            return lambda h, d, x, y, s, n: L.mfit_multi_dataset_upload(h, d, x, 0, y, s, n)
        if name == "mfit_dataset_upload_xd":
            return L.mfit_multi_dataset_upload
        if name == "mfit_chains_init":
            return lambda h, th, W, seed, base, cap: L.mfit_multi_chains_init(h, th, W, seed, cap)

        def no_streams(z, u):
            if z is not None or u is not None:
                raise NotImplementedError("injected z/u streams are a single-GPU test facility")
        if name == "mfit_run_adaptive":
            return lambda h, cfg, lm, z, u, ns, out: (no_streams(z, u), L.mfit_multi_run_adaptive(h, cfg, lm, out))[1]
        if name == "mfit_run_begin":
            return lambda h, cfg, lm, z, u, ns: (no_streams(z, u), L.mfit_multi_run_begin(h, cfg, lm))[1]
        if name == "mfit_run_many_steps":
            return lambda h, n, lm, z, u, ns: (no_streams(z, u), L.mfit_multi_run_many_steps(h, n, lm))[1]
        twin = "mfit_multi_" + name[len("mfit_"):]
        if name.startswith("mfit_") and twin in _lib.SIGNATURES:
            return getattr(L, twin)
        raise NotImplementedError("%s has no multi-GPU twin: use one GPU (device=) for it" % name)


class Walker:
    """W chains on one GPU -- or, with devices=[...], sharded over several GPUs of this box behind one handle -- behind
    the reference's walker interface."""

    def __init__(self, function, data, params, data_error=None, log_liklihood=None, log_prior=None,
                 n_walkers=None, theta0=None, seed=0, device=0, chain_id_base=0, trace_capacity=0,
                 nccl=None, dataset_owners=None, stream=None, per_walker=None, peers=None, devices=None, shard="chains"):
        L = _lib.lib()
        self.devices = list(devices) if devices is not None else None
        if self.devices is not None:
            if nccl is not None or peers is not None or dataset_owners is not None or per_walker is not None or chain_id_base:
                raise ValueError("devices=[...] shards inside the library: nccl / peers / dataset_owners / per_walker / chain_id_base do not apply")
            if shard not in ("chains", "datasets"):
                raise ValueError("shard must be 'chains' or 'datasets'")
            L = _MultiABI(L)
        self._L = L
        self.function = force_list(function)                        # :1142
        D = len(self.function)
        self.param_keys, p0 = plist(params)                         # :1145
        P = len(self.param_keys)
        self.P, self.D = P, D
        self.per_walker = per_walker
        if per_walker is not None:    # walker set with per-walker data (walker_set_create): one :function, W data sets
            if D != 1:
                raise ValueError("a walker set with per-walker data takes one :function")
            self.data, self.data_error = [(per_walker["x"], per_walker["y"])], [per_walker["sigma"]]
        else:
            self.data = clean_data(data, D)                             # :1143
            self.data_error = clean_data_error(data_error, self.data)   # :1144
        ll = log_liklihood if isinstance(log_liklihood, (list, tuple)) else [log_liklihood] * D
        self.log_liklihood = [x if x is not None else "log-liklihood-normal" for x in ll]
        if isinstance(log_prior, (list, tuple)) and log_prior and not isinstance(log_prior[0], (tuple, list)):
            lp = list(log_prior)
        elif isinstance(log_prior, (list, tuple)) and log_prior and isinstance(log_prior[0], (list, tuple)) \
                and log_prior[0] and isinstance(log_prior[0][0], (list, tuple)):
            lp = list(log_prior)            # list of bounds lists, one per dataset
        else:
            lp = [log_prior] * D
        self.log_prior = [x if x is not None else "log-prior-flat" for x in lp]
        if len(self.log_liklihood) != D or len(self.log_prior) != D:
            raise ValueError(":log-liklihood / :log-prior lists must match the :function list")
        self.param_style = "single-item" if len(self.param_keys) and self.param_keys[0].endswith("[0]") else "multiple-kwargs"
        self.device, self.seed, self.chain_id_base = device, int(seed), int(chain_id_base)
        self._keys_c = _lib.keys_array(self.param_keys)

        h = C.c_void_p()
        if self.devices is not None:
            ids = (C.c_int * len(self.devices))(*self.devices)
            rc = _lib.lib().mfit_multi_create(len(self.devices), ids, _lib.SHARD_DATASETS if shard == "datasets" else _lib.SHARD_CHAINS,
                                              C.byref(h))
        else:
            rc = L.mfit_ctx_create(device, stream, C.byref(h))
        if rc != 0:
            raise _lib.MfitError(rc, L.mfit_last_error(None).decode())
        self._h = h
        self._chk(L.mfit_problem_define(h, D, P))
        if nccl is not None:
            uid, rank, world = nccl
            self._chk(L.mfit_nccl_attach(h, uid, rank, world))
            self.rank, self.world = rank, world
        else:
            self.rank, self.world = 0, 1
        if peers is not None:
            # peer exchange for data sets sharded over the GPUs of one box (mfit_peer_setup / mfit_peer_attach):
            # peers = (rank, world, gather) with gather(bytes) -> the handles of all ranks concatenated in rank order
            prank, pworld, gather = peers
            n_chains = int(per_walker["y"].shape[0]) if per_walker is not None else \
                (int(np.shape(theta0)[0]) if theta0 is not None and np.ndim(theta0) == 2 else int(n_walkers or 1))
            handle = C.create_string_buffer(128)
            self._chk(L.mfit_peer_setup(h, prank, pworld, n_chains, -(-D // pworld), handle))
            allh = gather(handle.raw)
            if len(allh) != 128 * pworld:
                raise ValueError("peer handle exchange returned %d bytes for %d ranks" % (len(allh), pworld))
            self._chk(L.mfit_peer_attach(h, C.create_string_buffer(allh, len(allh))))
            self.rank, self.world = prank, pworld
        self.dataset_owners = None
        if dataset_owners is not None:
            own = np.ascontiguousarray(dataset_owners, dtype=np.int32)
            self._chk(L.mfit_dataset_owners(h, own.ctypes.data_as(C.POINTER(C.c_int32))))
            self.dataset_owners = list(own)
        for d in range(D):
            if per_walker is not None:
                pw = per_walker
                self._chk(L.mfit_dataset_upload_per_walker(h, d, _ptr(pw["x"]), _ptr(pw["y"]), _ptr(pw["sigma"]),
                                                           pw["y"].shape[1], pw["y"].shape[0], int(pw["x"].ndim == 1),
                                                           int(pw["sigma"].ndim == 1)))
            elif self.dataset_owners is None or self.dataset_owners[d] == self.rank:
                x, y = self.data[d]
                e = self.data_error[d]
                if x.ndim == 2:      # component-major for the device: one array per component of x
                    xt = np.ascontiguousarray(x.T)
                    self._chk(L.mfit_dataset_upload_xd(h, d, _ptr(xt), x.shape[1], _ptr(y), _ptr(e), len(y)))
                else:
                    self._chk(L.mfit_dataset_upload(h, d, _ptr(x), _ptr(y), _ptr(e), len(x)))
            self._set_function(d, self.function[d])
            ll_d = self.log_liklihood[d]
            if isinstance(ll_d, str) and ll_d.lstrip().startswith("("):
                # (create-log-liklihood-function (lambda (y model error) ...)) :402-417: the per-point lambda as text
                self._chk(L.mfit_likelihood_sexpr(h, d, create_log_liklihood_function(ll_d).encode()))
            else:
                kind = _LL.get(ll_d)
                if kind is None:
                    raise ValueError("unknown :log-liklihood %r" % (ll_d,))
                self._chk(L.mfit_likelihood_set(h, d, kind))
            self._set_prior(d, self.log_prior[d])

        if theta0 is None:
            W = int(per_walker["y"].shape[0] if per_walker is not None else (n_walkers or 1))
            theta0 = np.tile(p0, (W, 1))
        else:
            theta0 = _f64(theta0)
            if theta0.ndim == 1:
                theta0 = theta0[None, :]
            if theta0.shape[1] != P:
                raise ValueError("theta0 must be W x P")
        self.W, self.P, self.D = theta0.shape[0], P, D
        self._chk(L.mfit_chains_init(h, _ptr(_f64(theta0)), self.W, self.seed, self.chain_id_base, int(trace_capacity)))
        self.last_steps_launched = 0

    # -------------------------------------------------------------- plumbing
    def _chk(self, rc):
        if rc != 0:
            raise _lib.MfitError(rc, self._L.mfit_last_error(self._h).decode())

    def _set_function(self, d, f):
        L, h = self._L, self._h
        if isinstance(f, Builtin):
            try:
                idx = np.array([self.param_keys.index(k) for k in f.keys], dtype=np.int32)
            except ValueError as e:
                raise ValueError("function %d uses a key that is not in :params: %s" % (d, e))
            self._chk(L.mfit_model_builtin(h, d, families.family_id(f.name), f.degree,
                                           idx.ctypes.data_as(C.POINTER(C.c_int32)), len(idx)))
        elif isinstance(f, str):
            self._chk(L.mfit_model_sexpr(h, d, f.encode(), self._keys_c, self.P))
        else:
            raise TypeError(":function must be builtin(...) or Lisp lambda text; host closures cannot run on the GPU")

    def _set_prior(self, d, p):
        L, h = self._L, self._h
        if p in ("log-prior-flat", None):
            self._chk(L.mfit_prior_flat(h, d))
        elif isinstance(p, str):
            self._chk(L.mfit_prior_sexpr(h, d, p.encode(), self._keys_c, self.P))
        else:
            idx = np.array([self.param_keys.index(str(k).lstrip(":").lower()) if str(k).lstrip(":").lower() in self.param_keys
                            else -1 for k, _, _ in p], dtype=np.int32)
            lo = _f64([b[1] for b in p])
            hi = _f64([b[2] for b in p])
            self._chk(L.mfit_prior_bounds(h, d, len(idx), idx.ctypes.data_as(C.POINTER(C.c_int32)), _ptr(lo), _ptr(hi)))

    def close(self):
        if getattr(self, "_h", None):
            self._L.mfit_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -------------------------------------------------------------- walker-make-step on a batch (:1067-1070)
    def make_step(self, theta):
        """prob, ll[W,D], lp[W,D] for a batch of parameter vectors (W x P)."""
        theta = _f64(theta)
        if theta.ndim == 1:
            theta = theta[None, :]
        W = theta.shape[0]
        prob, ll, lp = np.empty(W), np.empty((W, self.D)), np.empty((W, self.D))
        self._chk(self._L.mfit_eval_logp(self._h, _ptr(theta), W, _ptr(prob), _ptr(ll), _ptr(lp)))
        return prob, ll, lp

    # -------------------------------------------------------------- drivers
    def _streams(self, z, u):
        nstream = 0
        if z is not None:
            z = _f64(z)
            if z.ndim == 2:
                z = z[None, :, :]
            if z.shape[0] != self.W or z.shape[2] != self.P:
                raise ValueError("z must be W x nstream x P")
            nstream = z.shape[1]
        if u is not None:
            u = _f64(u)
            if u.ndim == 1:
                u = u[None, :]
            if z is None or u.shape != z.shape[:2]:
                raise ValueError("u must be W x nstream and come with z")
        return z, u, nstream

    def adaptive_steps_full(self, n=100000, temperature=1e3, auto="prob-settle", sampling_optimization=None,
                            max_walker_length=None, l_matrix=None, z=None, u=None, greedy=False, max_steps=None):
        """walker-adaptive-steps-full :862-942.  `auto` = "prob-settle" (the reference's default: `(or :prob-settle
        :slope-settle nil)` evaluates to :prob-settle) or None / False for nil; :slope-settle is not supported.
        max_steps (not in the reference): leave the loop after that many take-step calls, as a stop request would
        (mfit-walker-estop :904) -- used by tests and benchmarks to time or compare the first iterations of a long run."""
        if auto in ("slope-settle", ":slope-settle"):
            raise NotImplementedError(":slope-settle runs a nested mcmc-fit on the host (:886-887); use :prob-settle")
        cfg = _lib.AdaptCfg(int(np.floor(n)), float(temperature), _AUTO[auto], _OPT[sampling_optimization],
                            int(max_walker_length or 0), int(bool(greedy)), 0)
        z, u, nstream = self._streams(z, u)
        lm = _f64(l_matrix) if l_matrix is not None else None
        if lm is not None and lm.shape != (self.P, self.P):
            raise ValueError(":l-matrix must be P x P")
        launched = C.c_int64()
        if max_steps is None:
            self._chk(self._L.mfit_run_adaptive(self._h, C.byref(cfg), _ptr(lm), _ptr(z), _ptr(u), nstream, C.byref(launched)))
        else:
            fin = C.c_int()
            self._chk(self._L.mfit_run_begin(self._h, C.byref(cfg), _ptr(lm), _ptr(z), _ptr(u), nstream))
            self._chk(self._L.mfit_run_advance(self._h, int(max_steps), C.byref(launched), C.byref(fin)))
        self.last_steps_launched = launched.value
        return self

    def adaptive_steps(self, n=30000, **kw):
        """walker-adaptive-steps :946-947"""
        return self.adaptive_steps_full(n=n, temperature=10, auto="prob-settle", **kw)

    def many_steps(self, n, l_matrix, z=None, u=None):
        """walker-many-steps :849-853 (the :median-params default l-matrix is host post-processing: pass one)"""
        z, u, nstream = self._streams(z, u)
        lm = _f64(l_matrix)
        if lm.shape == (self.W, self.P, self.P):     # one l-matrix per chain
            self._chk(self._L.mfit_run_many_steps_per_chain(self._h, int(n), _ptr(lm), _ptr(z), _ptr(u), nstream))
        elif lm.shape == (self.P, self.P):
            self._chk(self._L.mfit_run_many_steps(self._h, int(n), _ptr(lm), _ptr(z), _ptr(u), nstream))
        else:
            raise ValueError("l-matrix must be P x P, or W x P x P with one per chain")
        return self

    def sample_region(self, initial_scale=1e-3, z=None):
        """walker-sample-region :949-969"""
        z, _, nstream = self._streams(z, None)
        self._chk(self._L.mfit_run_sample_region(self._h, float(initial_scale), _ptr(z), nstream))
        return self

    def force_take_step(self):
        """walker-force-take-step :1124-1129"""
        self._chk(self._L.mfit_force_take_step(self._h))
        return self

    def request_stop(self):
        """(setf mfit-walker-estop t) :860"""
        self._L.mfit_request_stop(self._h)

    def segment_geometry(self, d=0):
        """(tiles per segment, segments) the likelihood kernel uses for data set d (mfit_segment_geometry)"""
        t, n = C.c_int32(), C.c_int32()
        self._chk(self._L.mfit_segment_geometry(self._h, int(d), C.byref(t), C.byref(n)))
        return t.value, n.value

    def set_small_path(self, mode):
        """0 = automatic, 1 = never, 2 = whenever the problem qualifies: many steps per launch for small problems (mfit_set_small_path)"""
        self._chk(self._L.mfit_set_small_path(self._h, int(mode)))

    def set_groups(self, n_groups):
        """0 = automatic, 1 / 2 = advance the chains as one / two groups on separate streams (mfit_set_groups)"""
        self._chk(self._L.mfit_set_groups(self._h, int(n_groups)))

    def modify(self, modify, burn_number=None, keep_number=None):
        """walker-modify :546-580 for every chain: :burn-walks, :keep-walks, :reset, :reset-to-most-likely"""
        m = str(modify).lstrip(":").lower()
        mode = {"burn-walks": 1, "keep-walks": 2, "reset": 3, "reset-to-most-likely": 4}[m]
        number = burn_number if mode == 1 else keep_number if mode == 2 else 0
        self._chk(self._L.mfit_walker_modify(self._h, mode, int(number)))
        return self

    def keep_walks(self, keep):
        """walker-modify :keep-walks :568-569"""
        self._chk(self._L.mfit_keep_walks(self._h, int(keep)))

    # -------------------------------------------------------------- readers
    def stats(self):
        arr = (_lib.ChainStats * self.W)()
        self._chk(self._L.mfit_stats_read(self._h, arr))
        return np.frombuffer(arr, dtype=[("length", "i8"), ("age", "i8"), ("steps_this_run", "i8"), ("last_prob", "f8"),
                                         ("best_prob", "f8"), ("temperature", "f8"), ("accept_200", "f8"),
                                         ("shutting_down", "i4"), ("l_updates", "i4")]).copy()

    def trace(self, take):
        """newest-first (theta[W,take,P], prob[W,take], n_valid[W]) like walker-get :steps :take"""
        take = int(take)
        th = np.empty((self.W, take, self.P))
        pr = np.empty((self.W, take))
        nv = np.empty(self.W, dtype=np.int64)
        self._chk(self._L.mfit_trace_read(self._h, take, _ptr(th), _ptr(pr), nv.ctypes.data_as(C.POINTER(C.c_int64))))
        return th, pr, nv

    def most_likely(self):
        th, pr = np.empty((self.W, self.P)), np.empty(self.W)
        self._chk(self._L.mfit_best_read(self._h, _ptr(th), _ptr(pr)))
        return th, pr

    def last_step(self):
        th, pr = np.empty((self.W, self.P)), np.empty(self.W)
        self._chk(self._L.mfit_last_read(self._h, _ptr(th), _ptr(pr)))
        return th, pr

    def l_matrix(self):
        out = np.empty((self.W, self.P, self.P))
        self._chk(self._L.mfit_l_matrix_read(self._h, _ptr(out)))
        return out

    def analytics(self, take=None, which=("most-likely-step", "median-params", "acceptance", "covariance-matrix", "l-matrix")):
        """walker-get's trace analytics for every chain at once, computed on the device (mfit_trace_analytics):
        dict with any of most_likely_theta[W,P], most_likely_prob[W], median[W,P], acceptance[W], covariance[W,P,P],
        l_matrix[W,P,P], stddev[W,P] (= diagonal of l_matrix, :stddev-params)."""
        if take is None:
            take = int(self.stats()["length"].max())
        W, P = self.W, self.P
        want = set(str(k).lstrip(":").lower() for k in which)
        bufs = {"most_likely_theta": np.empty((W, P)) if "most-likely-step" in want else None,
                "most_likely_prob": np.empty(W) if "most-likely-step" in want else None,
                "median": np.empty((W, P)) if "median-params" in want else None,
                "acceptance": np.empty(W) if "acceptance" in want else None,
                "covariance": np.empty((W, P, P)) if "covariance-matrix" in want else None,
                "l_matrix": np.empty((W, P, P)) if want & {"l-matrix", "stddev-params"} else None}
        self._chk(self._L.mfit_trace_analytics(self._h, int(take), _ptr(bufs["most_likely_theta"]), _ptr(bufs["most_likely_prob"]),
                                               _ptr(bufs["median"]), _ptr(bufs["acceptance"]), _ptr(bufs["covariance"]),
                                               _ptr(bufs["l_matrix"])))
        out = {k: v for k, v in bufs.items() if v is not None}
        if "l_matrix" in out:
            out["stddev"] = np.einsum("wii->wi", out["l_matrix"]).copy()
        return out

    def philox_stream(self, first_step, nsteps):
        """The normals z[W,nsteps,P] and uniforms u[W,nsteps] the device draws for take-step calls made at walker
        age first_step, first_step+1, ... (for replaying a Philox run through an injected-stream implementation)."""
        z, u = np.empty((self.W, nsteps, self.P)), np.empty((self.W, nsteps))
        self._chk(self._L.mfit_philox_stream(self._h, self.seed, self.chain_id_base, self.W, int(first_step), int(nsteps),
                                             self.P, _ptr(z), _ptr(u)))
        return z, u

    def get(self, get, take=None, param=None, chain=None):
        """walker-get :487-543 for the modes that read the trace (the list-walking analytics are host post-processing).
        Results carry a leading chain axis unless `chain` selects one."""
        g = str(get).lstrip(":").lower()
        st = self.stats()
        n = int(st["length"].max()) if take is None else int(take)
        sel = (lambda a: a if chain is None else a[chain])
        if g == "most-likely-params":   # :511-515 reads the walker's most-likely-step slot
            th, pr = self.most_likely()
            return sel(th)
        if g == "most-likely-step":     # :503-505 searches the newest `take` steps
            a = self.analytics(n, ["most-likely-step"])
            return sel(a["most_likely_theta"]), sel(a["most_likely_prob"])
        if g == "last-step":
            th, pr = self.last_step()
            return sel(th), sel(pr)
        if g == "median-params":      # :516-523, on the device
            return sel(self.analytics(n, ["median-params"])["median"])
        if g == "acceptance":         # :506-508
            return sel(self.analytics(n, ["acceptance"])["acceptance"])
        if g == "covariance-matrix":  # :541
            return sel(self.analytics(n, ["covariance-matrix"])["covariance"])
        if g == "l-matrix":           # :543
            return sel(self.analytics(n, ["l-matrix"])["l_matrix"])
        if g == "stddev-params":      # :526-539 (zeros for walks shorter than 10 steps)
            sd = self.analytics(n, ["stddev-params"])["stddev"]
            sd[st["length"] < 10] = 0.0
            return sel(sd)
        th, pr, nv = self.trace(max(n, 1))
        if g == "steps":
            return sel(th), sel(pr), sel(nv)
        if g == "params":
            return sel(th)
        if g == "log-liklihoods":
            return sel(pr)
        if g == "param":
            return sel(th[:, :, self.param_keys.index(str(param).lstrip(":").lower())])
        raise NotImplementedError("walker-get :%s" % g)

    def data_and_fit(self, take=1000, x_column=0, y_column=1, fn_number=0, which_solution="most-likely", x_shift=None,
                     y_shift=None, n_fit=1000):
        """walker-get-data-and-fit :1231-1255 for every chain at once (mfit_fit_envelope) ->
        (x_fit[n_fit], max_ys[W,n_fit], min_ys[W,n_fit], y_fit[W,n_fit], x_data, y_data, params[W,P]): the model at the
        most likely / median parameters of the newest `take` steps and, at every abscissa, the largest and smallest
        model value over the ceil(0.66 take) most probable entries of each chain's whole walk."""
        if self.per_walker:
            raise NotImplementedError("walker sets with per-walker data: one x range per walker")
        st = self.stats()
        take = int(st["length"].max()) if take is None else int(take)
        cols = self.data[fn_number]
        x_data, y_data = np.asarray(cols[x_column]), np.asarray(cols[y_column])
        x_fit = linspace(float(np.min(x_data)), float(np.max(x_data)), n_fit)   # :1243
        W, P = self.W, self.P
        y_fit, y_max, y_min, par = np.empty((W, n_fit)), np.empty((W, n_fit)), np.empty((W, n_fit)), np.empty((W, P))
        which = _lib.FIT_MOST_LIKELY if str(which_solution).lstrip(":").lower() == "most-likely" else _lib.FIT_MEDIAN
        self._chk(self._L.mfit_fit_envelope(self._h, int(fn_number), take, which, _ptr(x_fit), n_fit, _ptr(y_fit),
                                            _ptr(y_max), _ptr(y_min), _ptr(par), None))
        xs = 0.0 if x_shift is None else x_shift
        ys = 0.0 if y_shift is None else y_shift
        return x_fit + xs, y_max + ys, y_min + ys, y_fit + ys, x_data + xs, y_data + ys, par

    # -------------------------------------------------------------- walker-save / walker-load (:971-1001)
    def save(self, path, take=None):
        """s-expression header in the layout of the reference's (commented-out) walker-save -- :fn :param-keys
        :log-liklihood :log-prior as *recommendations* for walker-load, like the reference prints them -- followed by
        binary FP64 blobs: the walk, W x rows x (P+1) doubles (prob first, newest step first), then every chain's most
        likely step, W x (P+1).  With the ages and the Philox key in the header this is a checkpoint: walker_load
        continues the chains exactly where they stood."""
        st = self.stats()
        rows = int(st["length"].max()) if take is None else int(take)
        th, pr, nv = self.trace(rows)
        bth, bpr = self.most_likely()
        fn_names = " ".join('"%s"' % (repr(f) if isinstance(f, Builtin) else f).replace('"', '\\"') for f in self.function)
        hdr = ("(:mfit-walker 2 :fn (%s) :param-keys (%s) :log-liklihood (%s) :log-prior (%s) :n-walkers %d :rows %d "
               ":n-params %d :seed %d :chain-id-base %d :lengths (%s) :ages (%s))\n") % (
            fn_names, " ".join(":" + k for k in self.param_keys),
            " ".join('"%s"' % x for x in self.log_liklihood),
            " ".join('"%s"' % (x if isinstance(x, str) else repr(list(x))).replace('"', '\\"') for x in self.log_prior),
            self.W, rows, self.P, self.seed, self.chain_id_base, " ".join(str(int(v)) for v in nv),
            " ".join(str(int(v)) for v in st["age"]))
        blob = np.concatenate([pr[:, :, None], th], axis=2).astype("<f8")
        best = np.concatenate([bpr[:, None], bth], axis=1).astype("<f8")
        hb = hdr.encode()
        with open(path, "wb") as f:
            # lambda texts are usually multi-line Lisp source: the header is length-prefixed by a Lisp comment line
            f.write((";; mfit-walker header-bytes=%d\n" % len(hb)).encode())
            f.write(hb)
            f.write(struct.pack("<q", blob.size))
            f.write(blob.tobytes())
            f.write(struct.pack("<q", best.size))
            f.write(best.tobytes())
        return path

    def restore(self, theta, prob, n_valid, age=None, best_theta=None, best_prob=None):
        """walker-modify :add-walks as walker-load uses it (:989-1001), for all chains: replace every chain's walk by
        theta[W,rows,P] / prob[W,rows] (newest first, n_valid[w] rows each) -- mfit_chains_restore."""
        theta, prob = _f64(theta), _f64(prob)
        nv = np.ascontiguousarray(n_valid, dtype=np.int64)
        ag = None if age is None else np.ascontiguousarray(age, dtype=np.int64)
        if theta.shape != (self.W, prob.shape[1], self.P) or prob.shape[0] != self.W or nv.shape != (self.W,):
            raise ValueError("theta must be W x rows x P, prob W x rows, n_valid W")
        lp = C.POINTER(C.c_int64)
        self._chk(self._L.mfit_chains_restore(
            self._h, prob.shape[1], _ptr(theta), _ptr(prob), nv.ctypes.data_as(lp), None if ag is None else ag.ctypes.data_as(lp),
            _ptr(None if best_theta is None else _f64(best_theta)), _ptr(None if best_prob is None else _f64(best_prob))))
        return self


def walker_load(path, function=None, data=None, data_error=None, log_liklihood=None, log_prior=None, **gpu):
    """walker-load :989-1001.  Without :function / :data the file is only read -> dict(header, param_keys, prob[W,rows],
    theta[W,rows,P], n_valid[W], age[W], best_theta, best_prob, seed, chain_id_base); the header names the functions the
    walker was saved with (the reference prints them as *Recommendations*).  With them, the walker set is rebuilt
    (walker-create with the newest step of every chain as its parameters, :999) and its walks restored (:add-walks),
    Philox key and ages included, so that it goes on exactly like the saved one."""
    import re
    with open(path, "rb") as f:
        first = f.readline().decode()
        m = re.match(r";; mfit-walker header-bytes=(\d+)", first)
        hdr = f.read(int(m.group(1))).decode() if m else first   # files written before the length prefix: one line
        (cnt,) = struct.unpack("<q", f.read(8))
        blob = np.frombuffer(f.read(cnt * 8), dtype="<f8")
        tail = f.read(8)
        best = None
        if len(tail) == 8:
            (bcnt,) = struct.unpack("<q", tail)
            best = np.frombuffer(f.read(bcnt * 8), dtype="<f8")
    W = int(re.search(r":n-walkers (\d+)", hdr).group(1))
    rows = int(re.search(r":rows (\d+)", hdr).group(1))
    P = int(re.search(r":n-params (\d+)", hdr).group(1))
    keys = [k.lstrip(":") for k in re.search(r":param-keys \(([^)]*)\)", hdr).group(1).split()]
    lens = np.array([int(v) for v in re.search(r":lengths \(([^)]*)\)", hdr).group(1).split()], dtype=np.int64)
    m = re.search(r":ages \(([^)]*)\)", hdr)
    ages = np.array([int(v) for v in m.group(1).split()], dtype=np.int64) if m else lens.copy()
    m = re.search(r":chain-id-base (\d+)", hdr)
    blob = blob.reshape(W, rows, P + 1)
    out = {"header": hdr, "param_keys": keys, "prob": blob[:, :, 0].copy(), "theta": blob[:, :, 1:].copy(), "n_valid": lens,
           "age": ages, "seed": int(re.search(r":seed (\d+)", hdr).group(1)), "chain_id_base": int(m.group(1)) if m else 0,
           "best_theta": None, "best_prob": None}
    if best is not None:
        best = best.reshape(W, P + 1)
        out["best_prob"], out["best_theta"] = best[:, 0].copy(), best[:, 1:].copy()
    if function is None or data is None:
        return out
    gpu.setdefault("seed", out["seed"])
    gpu.setdefault("chain_id_base", out["chain_id_base"])
    theta0 = out["theta"][:, 0, :]
    w = Walker(function, data, [(k, v) for k, v in zip(keys, theta0[0])], data_error, log_liklihood, log_prior,
               theta0=theta0, **gpu)
    return w.restore(out["theta"], out["prob"], out["n_valid"], out["age"], out["best_theta"], out["best_prob"])


# ------------------------------------------------------------------ the reference's function names
def walker_create(function=None, data=None, params=None, data_error=None, log_liklihood=None, log_prior=None,
                  param_bounds=None, **gpu):
    """walker-create :1132-1163 (param-bounds is ignorable in the reference too, :1141).  Extra keywords select the
    GPU-side batch: n_walkers / theta0, seed, device, chain_id_base, trace_capacity, nccl, dataset_owners."""
    return Walker(function, data, params, data_error, log_liklihood, log_prior, **gpu)


def walker_set_create(function=None, data=None, params=None, data_error=None, log_liklihood=None, log_prior=None, **gpu):
    """A walker set with per-walker data, the reference's own many-walker workload: file->nv-walkers makes one walker per
    spectrum -- (x y_k) pairs with a common x -- and advances them one after another with mapc (nv-specific.lisp:50-66).
    `data` is a list of W (x y) sets of equal length, or (x, Y) with Y of shape W x n sharing one x; `data_error` a
    number, one n-vector for all walkers, or W x n.  :function / :log-liklihood / :log-prior are shared; `params` is the
    common starting point unless theta0 (W x P) is given.  All W walkers advance together, one warp per walker."""
    if isinstance(data, (list, tuple)) and len(data) == 2 and np.ndim(data[1]) == 2 and np.ndim(data[0]) == 1:
        x, Y = _f64(data[0]), _f64(data[1])
    else:
        sets = [(_f64(s_[0]), _f64(s_[1])) for s_ in data]
        n = len(sets[0][0])
        if any(len(a) != n or len(b) != n for a, b in sets):
            raise ValueError("every walker's data set must have the same number of points")
        Y = np.stack([b for _, b in sets])
        x = sets[0][0] if all(np.array_equal(a, sets[0][0]) for a, _ in sets) else np.stack([a for a, _ in sets])
    W, n = Y.shape
    if x.shape not in ((n,), (W, n)):
        raise ValueError("x must have n points, or W x n")
    if data_error is None:
        data_error = 1.0                                            # (if data-error data-error 1) :1144
    e = _f64(data_error)
    if e.ndim == 0 or (e.ndim == 1 and len(e) != n):                # clean-data-error :794-805: the first number, broadcast
        e = np.full(n, float(e.ravel()[0]))
    if e.shape not in ((n,), (W, n)):
        raise ValueError("data_error must be a number, n values, or W x n")
    pw = {"x": np.ascontiguousarray(x), "y": np.ascontiguousarray(Y), "sigma": np.ascontiguousarray(e)}
    return Walker(function, None, params, None, log_liklihood, log_prior, per_walker=pw, **gpu)


def walker_adaptive_steps_full(walker, **kw):
    return walker.adaptive_steps_full(**kw)


def walker_adaptive_steps(walker, n=30000):
    return walker.adaptive_steps(n)


def walker_many_steps(walker, n, l_matrix=None):
    if l_matrix is None:   # :851 (diagonal-covariance (scale-plist 1e-2 median-params)); 1e-2 is a single-float literal
        med = walker.get("median-params")             # per chain, as the reference makes it per walker
        l_matrix = np.float64(np.float32(1e-2)) * med[:, :, None] * np.eye(walker.P)[None, :, :]
    return walker.many_steps(n, l_matrix)


def walker_sample_region(walker, initial_scale=1e-3):
    return walker.sample_region(initial_scale)


def walker_force_take_step(walker):
    return walker.force_take_step()


def walker_get(walker, get, take=None, param=None, chain=None):
    return walker.get(get, take=take, param=param, chain=chain)


def walker_save(walker, path, take=None):
    return walker.save(path, take)


# "Functionality for a set of similar walker instances" :1003-1036.  A Walker here IS a set -- W chains advanced
# together -- so the set forms are the plain ones with the leading chain axis kept.
def walker_set_get(the_walker_set, get, take=None, param=None):
    """walker-set-get :1029-1030: (mapcar (lambda (x) (walker-get x ...)) the-walker-set) -> results with a chain axis"""
    return the_walker_set.get(get, take=take, param=param)


def walker_set_save(the_walker_set, path, take=None):
    """walker-set-save :1004-1010 (dead code in the reference): every chain's walk in one file"""
    return the_walker_set.save(path, take)


def walker_set_load(path, function, data, data_error=None, log_liklihood=None, log_prior=None, **gpu):
    """walker-set-load :1012-1027 (dead code in the reference): rebuild the set and restore its walks"""
    return walker_load(path, function=function, data=data, data_error=data_error, log_liklihood=log_liklihood,
                       log_prior=log_prior, **gpu)


def walker_set_delete(the_walker_set):
    """walker-set-delete :1032-1033 / walker-modify :delete :579-580: frees the device memory of the set"""
    the_walker_set.close()


def mcmc_fit(**kw):
    """mcmc-fit :1165-1175"""
    w = walker_create(**kw)
    walker_adaptive_steps(w)
    return w
