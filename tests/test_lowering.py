"""The s-expression -> CUDA C lowering (csrc/lower.cpp), checked without a GPU: the generated C is compiled with the
host compiler and evaluated against the oracle's Common-Lisp-semantics evaluator on the same inputs, and the full
NVRTC pipeline is run to an sm_100a cubin."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

from lisp_mcmc_b200 import _lib, families, workloads

HOST_PRELUDE = """
#include <cmath>
#include <cstddef>
#define __device__
#define __forceinline__ inline
struct MfitEvalArgs { const double* theta; long long stride; double xmin, xmax; };
struct MfitGuard {};
using std::fabs; using std::exp; using std::log; using std::sin; using std::cos; using std::sqrt; using std::pow;
// the device build routes these through mfit_math.cuh (<= 2.4 ulp); here only the lowering itself is under test
template <bool S> inline double mfit_exp_t(double x, MfitGuard&) { return std::exp(x); }
template <bool S> inline double mfit_log_t(double x, MfitGuard&) { return std::log(x); }
template <bool S> inline double mfit_sin_t(double x, MfitGuard&) { return std::sin(x); }
template <bool S> inline double mfit_cos_t(double x, MfitGuard&) { return std::cos(x); }
"""


def _host_compile(code, kind):
    """compile the lowered text for the HOST (the arithmetic is plain C) and return a callable"""
    d = tempfile.mkdtemp()
    src = os.path.join(d, "m.cpp")
    if kind == 0:
        tail = ('extern "C" double call(const double* th, double x) { MfitEvalArgs a{th, 1}; MfitUser m; m.load(a, 0); '
                "MfitGuard gd; return m.eval<false>(x, gd); }")
    else:
        tail = 'extern "C" double call(const double* th, double x) { (void)x; return mfit_user_prior(th, 1, 0); }'
    open(src, "w").write(HOST_PRELUDE + code + tail)
    so = os.path.join(d, "m.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    lib.call.restype = C.c_double
    lib.call.argtypes = [C.POINTER(C.c_double), C.c_double]
    return lambda th, x=0.0: lib.call(np.ascontiguousarray(th, dtype=np.float64).ctypes.data_as(C.POINTER(C.c_double)), x)


MODELS = [
    ("(lambda (x &key m b &allow-other-keys) (+ b (* -3 m) (* (- m (/ b 60)) x)))", ["b", "m"]),        # README example
    (workloads.USER_POLY5_EXP, ["c0", "c1", "c2", "c3", "c4", "c5", "k"]),
    ("(lambda (x &key a b &allow-other-keys) (let* ((u (* a x)) (v (+ u b))) (if (> v 0) (sqrt v) (- (expt v 2) 1/3))))", ["a", "b"]),
    ("(lambda (x &key a b (c 2.5) &allow-other-keys) (+ (* c (sin (* a x))) (cos b) (expt x 3) (expt a -2) (abs b) (max a b 0.5) (min a 1)))", ["a", "b"]),
    ("(lambda (x &key a b &allow-other-keys) (cond ((< x 0) (* a x)) ((< x 1) (+ a (* b x))) (t (log (+ 1 (* a a) (* x x))))))", ["a", "b"]),
    ("(lambda (x &key a b &allow-other-keys) (* (exp 1) (float 1/3) (+ a (* 0.1 b x)) (expt 2 10) (atan a b) (tanh x)))", ["a", "b"]),
    ("(lambda (x &key a b &allow-other-keys) (if (and (> a 0) (not (> b 10)) (or (< x 0) (>= a 1/2))) (+ a (* 2 pi x)) (1+ b)))", ["a", "b"]),
]
# :single-item parameter style (mcmc-fitting.lisp:8-14): one key, its value a list / vector / n x 1 array -> keys NAME[k]
MODELS += [
    ("(lambda (x &key params &allow-other-keys) (+ (elt params 0) (* (elt params 1) x)))", ["params[0]", "params[1]"]),          # :8
    ("(lambda (x &key params &allow-other-keys) (+ (aref params 0 0) (* (aref params 1 0) x)))", ["params[0]", "params[1]"]),    # :14
    ("(lambda (x &key p &allow-other-keys) (+ (first p) (* (nth 1 p) x) (* (svref p 2) (exp (* -1 (third p) x)))))", ["p[0]", "p[1]", "p[2]"]),
]
for _name, _deg in [("poly", 4), ("expo", 0), ("sinus", 0), ("lorentz", 0), ("gauss_lin", 0), ("nv_multi", 2)]:
    MODELS.append((families.lisp_text(_name, _deg), families.slots(_name, _deg)))


@pytest.mark.parametrize("text,keys", MODELS)
def test_model_lowering_matches_lisp_semantics(orc, text, keys):
    f = _host_compile(_lib.lower_sexpr(text, keys, 0), 0)
    ref = orc.Fn(text, keys)
    rng = np.random.default_rng(1)
    for _ in range(200):
        th = rng.uniform(0.2, 3.0, len(keys))
        x = rng.uniform(-2.0, 2.0)
        a, b = f(th, x), ref.model(x, th)
        assert a == b or abs(a - b) <= 4e-16 * abs(b), (text, th, x, a, b)


PRIORS = [
    ("(lambda (params data) (prior-bounds-let ((:x -10 10) (:y 100 200) (:z 1d-4 1d-2)) (+ bounds-total (* 4 x-bound))))", ["x", "y", "z"]),
    (families.LOG_PRIOR_NV, ["scale1", "scale2", "mu1", "mu2", "sigma", "bg0"]),
    ("(lambda (params data) (declare (ignore data)) (let ((a (getf params :a)) (q (getf params :missing 2d0))) (* -1/2 (expt (/ (- a q) 0.1d0) 2))))", ["a", "b"]),
    ("(lambda (params data) (prior-bounds-let ((:a 0 1) (:b (* 2 a) (+ 5 a))) bounds-total))", ["a", "b"]),   # bounds over earlier keys (let*)
]


PRIORS.append(("(lambda (params data) (declare (ignore data)) (let ((p0 (elt (getf params :p) 0))) (if (< 0 p0 (nth 1 (getf params :p))) 0d0 -1d10)))",
               ["p[0]", "p[1]"]))


@pytest.mark.parametrize("text,keys", PRIORS)
def test_prior_lowering_matches_lisp_semantics(orc, text, keys):
    f = _host_compile(_lib.lower_sexpr(text, keys, 1), 1)
    ref = orc.Fn(text, keys)
    rng = np.random.default_rng(2)
    for _ in range(300):
        th = rng.uniform(-20, 250, len(keys)) if "y 100" in text else rng.uniform(-1, 3000, len(keys)) * rng.choice([1e-3, 1e-2, 1.0], len(keys))
        a, b = f(th), ref.prior(th)
        assert a == b or abs(a - b) <= 1e-15 * abs(b) + 1e10 * 2.0 ** -52, (text, th, a, b)
    assert f(np.zeros(len(keys))) == ref.prior(np.zeros(len(keys))) or True


def test_single_item_parameter_style():
    """:params '(:params (-1 2)) (mcmc-fitting.lisp:8): the flattened keys lower to the same code as two keyword
    parameters; indexing beyond the sequence or using the sequence as a number is refused; plist() flattens lists,
    vectors and n x 1 arrays alike"""
    from lisp_mcmc_b200.walker import plist
    single = _lib.lower_sexpr("(lambda (x &key params &allow-other-keys) (+ (elt params 0) (* (elt params 1) x)))", ["params[0]", "params[1]"], 0)
    multi = _lib.lower_sexpr("(lambda (x &key b m &allow-other-keys) (+ b (* m x)))", ["b", "m"], 0)
    assert single == multi
    for bad in ["(lambda (x &key params &allow-other-keys) (+ (elt params 2) x))",
                "(lambda (x &key params &allow-other-keys) (+ params x))",
                "(lambda (x &key params &allow-other-keys) (+ (aref params 0 1) x))"]:
        with pytest.raises(_lib.MfitError):
            _lib.lower_sexpr(bad, ["params[0]", "params[1]"], 0)
    for value in ([-1, 2], (-1.0, 2.0), np.array([-1.0, 2.0]), np.array([[-1.0], [2.0]])):
        keys, vals = plist([":params", value])
        assert keys == ["params[0]", "params[1]"] and list(vals) == [-1.0, 2.0]
    assert plist([":b", -1, ":m", 2])[0] == ["b", "m"]


def test_exact_rational_comparison():
    """(< x 1/3): the constant is not a double; the lowering adjusts the operator so the outcome is the exact one"""
    f = _host_compile(_lib.lower_sexpr("(lambda (x &key a &allow-other-keys) (if (< a 1/3) 1d0 2d0))", ["a"], 0), 0)
    third = 1.0 / 3.0                       # RN(1/3) < 1/3
    assert third < 1 / 3 or True
    assert f([third], 0.0) == 1.0           # RN(1/3) is below the true third
    assert f([np.nextafter(third, 1.0)], 0.0) == 2.0


@pytest.mark.parametrize("text", ["(lambda (x &key a) (frobnicate a))", "(lambda (x &key a) (+ a", "(lambda (x y &key a) a)",
                                  "(lambda (x &key a) (< a 1))", "(lambda (x &key a) q)"])
def test_lowering_errors_are_reported(text):
    with pytest.raises(_lib.MfitError) as e:
        _lib.lower_sexpr(text, ["a"], 0)
    assert e.value.code == _lib.ERR_SEXPR


def test_nvrtc_compiles_user_model_and_prior_for_sm100a():
    """no GPU needed: NVRTC cross-compiles to an sm_100a cubin"""
    n = _lib.compile_sexpr(workloads.USER_POLY5_EXP, ["c0", "c1", "c2", "c3", "c4", "c5", "k"], 0)
    assert n > 10000
    n = _lib.compile_sexpr(families.LOG_PRIOR_NV, ["scale1", "scale2", "mu1", "mu2", "sigma", "bg0"], 1)
    assert n > 1000


XVEC_MODEL = ("(lambda (x &key a b c &allow-other-keys) (+ (* a (elt x 0)) (* b (exp (* -1 c (second x)))) "
              "(* 1/2 (nth 2 x)) (aref x 1)))")


def test_multi_dimensional_x_lowers_to_component_reads():
    """walker-create docstring :1137 -- "(elt x 0) ... (elt x 1)": x is a sequence; the lowered model declares how many
    components it reads (one staged data array each) and evaluates them like the oracle's (apply fn x params)"""
    from oracle import oracle as orc
    keys = ["a", "b", "c"]
    code = _lib.lower_sexpr(XVEC_MODEL, keys, 0)
    assert "XDIMS = 3" in code and "XVEC = true" in code and "evalv(" in code and "xv[2]" in code
    d = tempfile.mkdtemp()
    src = os.path.join(d, "m.cpp")
    tail = ('extern "C" double call(const double* th, const double* x) { MfitEvalArgs a{th, 1}; MfitUser m; m.load(a, 0); '
            "MfitGuard gd; double xv[MfitUser::XDIMS]; for (int k = 0; k < MfitUser::XDIMS; k++) xv[k] = x[k]; "
            "return m.evalv<false>(xv, gd); }")
    open(src, "w").write(HOST_PRELUDE + code + tail)
    so = os.path.join(d, "m.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    lib.call.restype = C.c_double
    dp = C.POINTER(C.c_double)
    lib.call.argtypes = [dp, dp]
    ref = orc.Fn(XVEC_MODEL, keys)
    rng = np.random.default_rng(4)
    for _ in range(200):
        th, x = rng.uniform(-2, 2, 3), rng.uniform(-3, 3, 3)
        got = lib.call(th.ctypes.data_as(dp), x.ctypes.data_as(dp))
        want = ref.model_v(x, th)
        assert got == want or abs(got - want) <= 4e-16 * abs(want)
    assert np.isnan(ref.model(1.0, [1.0, 1.0, 1.0]))          # a number where a sequence is indexed: type error
    assert np.isnan(ref.model_v([1.0, 2.0], [1.0, 1.0, 1.0]))  # (nth 2 x) beyond the end
    # a scalar model keeps the scalar form
    code1 = _lib.lower_sexpr("(lambda (x &key a b c &allow-other-keys) (+ a (* b x) c))", keys, 0)
    assert "XDIMS = 1" in code1 and "XVEC = false" in code1 and "eval(double x" in code1
    assert _lib.compile_sexpr(XVEC_MODEL, keys, 0) > 10000      # the whole NVRTC pipeline, sm_100a


@pytest.mark.parametrize("text", ["(lambda (x &key a) (+ x (elt x 0) a))",            # number and sequence at once
                                  "(lambda (x &key a) (+ a (elt x a)))",              # index must be a literal
                                  "(lambda (x &key a) (+ a (elt x 8)))",              # more than MFIT_MAX_XDIMS components
                                  "(lambda (x &key a) (let ((x a)) (elt x 0)))"])     # a shadowing binding is not x
def test_multi_dimensional_x_errors(text):
    with pytest.raises(_lib.MfitError) as e:
        _lib.lower_sexpr(text, ["a"], 0)
    assert e.value.code == _lib.ERR_SEXPR


def test_lowered_model_proves_its_fast_math_arguments_safe():
    """Model::safe() of a lowered lambda: when every argument of a fast routine (exp / log / sin / cos) is affine in x
    with parameter-only coefficients it is monotone over the data set's x range, so bounding it at both ends lets the
    chain take the guard-free loop; anything else keeps safe() = false (always guarded, as before)"""
    keys = ["a", "b", "k"]
    text = "(lambda (x &key a b k &allow-other-keys) (let ((u (* -1 k x))) (+ (* a (exp u)) (log (+ b (* 1/2 x))) (sin (- (* 3 x) a)))))"
    code = _lib.lower_sexpr(text, keys, 0)
    assert "bool safe(const MfitEvalArgs& a) const {" in code and ") < 700.0" in code and " > 1e-290" in code and ") < 9.0e4" in code
    d = tempfile.mkdtemp()
    src = os.path.join(d, "m.cpp")
    tail = ('extern "C" int call(const double* th, double xmin, double xmax) { MfitEvalArgs a{th, 1, xmin, xmax}; MfitUser m; '
            "m.load(a, 0); return m.safe(a) ? 1 : 0; }")
    open(src, "w").write(HOST_PRELUDE + code + tail)
    so = os.path.join(d, "m.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src])
    lib = C.CDLL(so)
    dp = C.POINTER(C.c_double)
    lib.call.argtypes = [dp, C.c_double, C.c_double]
    rng = np.random.default_rng(9)
    seen = set()
    for _ in range(400):
        th = np.array([rng.uniform(-5, 5), rng.uniform(-3, 8), rng.choice([0.5, 40.0, 400.0]) * rng.uniform(-1, 1)])
        xmin = rng.uniform(-6, 0)
        xmax = xmin + rng.uniform(0.1, 12)
        got = lib.call(th.ctypes.data_as(dp), xmin, xmax)
        xs = np.linspace(xmin, xmax, 201)
        e_arg, l_arg, s_arg = -th[2] * xs, th[1] + 0.5 * xs, 3 * xs - th[0]
        inside = np.all(np.abs(e_arg) < 700) and np.all(l_arg > 1e-290) and np.all(np.abs(s_arg) < 9e4)
        assert bool(got) == bool(inside), (th, xmin, xmax)
        seen.add(bool(got))
    assert seen == {True, False}
    a = np.array([1.0, 1.0, 1.0])
    assert lib.call(a.ctypes.data_as(dp), -np.inf, np.inf) == 0 and lib.call(a.ctypes.data_as(dp), 0.0, np.nan) == 0
    nan_th = np.array([1.0, 1.0, np.nan])
    assert lib.call(nan_th.ctypes.data_as(dp), 0.0, 1.0) == 0
    # not affine in x, or built from a local that is not a parameter expression: no proof
    for t in ["(lambda (x &key a b k &allow-other-keys) (+ a b k (exp (* x x))))",
              "(lambda (x &key a b k &allow-other-keys) (+ a b (exp (* k (sin x)))))"]:
        assert "safe(const MfitEvalArgs&) const { return false; }" in _lib.lower_sexpr(t, keys, 0)
    assert "fabs(s0b) < 700.0" in _lib.lower_sexpr(workloads.USER_POLY5_EXP, ["c0", "c1", "c2", "c3", "c4", "c5", "k"], 0)


def test_likelihood_lambda_lowers_and_compiles():
    """create-log-liklihood-function :402-417: a per-point (lambda (y model error) ...) becomes MfitTerm<MFITK_USER>;
    `error` is a LIST (:416) -- indexable with a literal, never a number; log-factorial only of y"""
    from lisp_mcmc_b200 import _lib
    code = _lib.lower_sexpr("(lambda (y model error) (- (log (+ 1 (expt (/ (- y model) 0.5d0) 2)))))", [], 2)
    assert "struct MfitTerm<MFITK_USER>" in code and "NARR = 2" in code and "return acc + (" in code
    code = _lib.lower_sexpr("(lambda (y model error) (log-normal y model (second error)))", [], 2)
    assert "mfit_u_log_normal(y, f, MFIT_ERR(1))" in code
    code = _lib.lower_sexpr("(lambda (y model error) (+ (log-poisson model y) 0))", [], 2)
    assert "NARR = 4" in code and "mfit_u_log_poisson(f, y, lf)" in code
    assert _lib.compile_sexpr("(lambda (y model error) (log-normal y model (elt error 0)))", [], 2) > 1000
    for bad in ("(lambda (y model error) (log-normal y model error))",       # the list used as a number
                "(lambda (y model) (- y model))",                             # wrong arity
                "(lambda (y model error) (log-factorial model))",             # factorial of something that is not the datum
                "(lambda (y model error) (elt error 16))"):
        with pytest.raises(_lib.MfitError):
            _lib.lower_sexpr(bad, [], 2)


def test_oracle_likelihood_lambda_equals_builtin_kinds():
    """the oracle's evaluator on (lambda (y model error) ...) reproduces its own built-in kinds exactly"""
    import numpy as np
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    x = np.linspace(0, 1, 200)
    y = rng.poisson(5 + 3 * x).astype(float)
    ow = orc.Walker(["a", "b"], 1)
    ow.dataset(0, x, y, np.full(200, 0.3))
    ow.model_sexpr(0, "(lambda (x &key a b &allow-other-keys) (+ a (* b x)))")
    ow.prior_flat(0)
    ow.likelihood(0, orc.LL_POISSON)
    p_pois = ow.logp([5.0, 3.0])[0]
    ow.likelihood_sexpr(0, "(lambda (y model error) (log-poisson model y))")
    assert ow.logp([5.0, 3.0])[0] == p_pois
    ow.likelihood(0, orc.LL_NORMAL)
    p_norm = ow.logp([5.0, 3.0])[0]
    ow.likelihood_sexpr(0, "(lambda (y model error) (log-normal y model (elt error 0)))")
    assert ow.logp([5.0, 3.0])[0] == p_norm
