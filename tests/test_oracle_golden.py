"""Pins the CPU oracle on every known answer the reference carries for the path and on closed forms.

The reference (Common Lisp) cannot be executed in this image, so these are the only anchors; log-likelihood values
and accept/reject sequences are pinned only by the restatement itself ("parity unpinned" beyond it)."""
import json
import math
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_example_covariance_and_l_matrix(orc):
    """mcmc-fitting.lisp:729-751: example-lplist -> example-covariance, example-l-matrix (values in the comments).
    lplist-to-l-matrix takes the plists as they are (no differencing)."""
    g = json.load(open(os.path.join(HERE, "golden", "reference_known_answers.json")))
    rows = np.array(g["example_lplist"])
    cov = orc.lplist_covariance(rows)
    assert np.array_equal(cov, np.array(g["example_covariance"]))
    L = orc.cholesky_decomp(cov)
    assert np.array_equal(L, np.array(g["example_l_matrix"]))   # every printed digit


def test_log_normal_closed_form(orc):
    import mpmath as mp
    mp.mp.dps = 40
    for x, mu, s in [(1.0, 0.5, 0.2), (-3.0, 2.0, 1e-7), (5e-6, 4.9e-6, 1e-7), (0.0, 0.0, 3.0)]:
        ref = -mp.log(2 * mp.pi) / 2 - mp.log(s) - ((mp.mpf(x) - mu) / s) ** 2 / 2
        got = orc.log_normal(x, mu, s)
        assert abs(got - float(ref)) <= 4e-16 * max(1.0, abs(float(ref)))


def test_log_factorial_is_single_float_accumulation(orc):
    """log-factorial :379-380: (log <integer>) is a single-float in CL, so the sum runs in single precision"""
    assert orc.log_factorial(0) == 0.0 and orc.log_factorial(1) == 0.0
    acc = np.float32(0.0)
    for j in range(1, 31):
        acc = np.float32(acc + np.float32(math.log(j)))
    assert orc.log_factorial(30) == float(acc)
    assert orc.log_factorial(30) != math.lgamma(31)            # differs from the exact value ...
    assert abs(orc.log_factorial(30) - math.lgamma(31)) < 1e-4   # ... by single-precision rounding only


def test_log_poisson_formula(orc):
    lam, k = 7.5, 4.0
    assert orc.log_poisson(lam, k) == (k * math.log(lam) - lam) - orc.log_factorial(k)


def test_line_fit_header_example(orc):
    """header example :5-17: data ((-4 -1 2 5 10) (0 2 5 9 13)), sigma 0.2, params (:b -1 :m 2), a line"""
    ow = orc.Walker(["b", "m"], 1)
    x, y = [-4.0, -1, 2, 5, 10], [0.0, 2, 5, 9, 13]
    ow.dataset(0, x, y, 0.2)
    ow.model_sexpr(0, "(lambda (x &key m b &allow-other-keys) (+ b (* m x)))")
    ow.likelihood(0, orc.LL_NORMAL)
    prob, ll, lp = ow.logp([-1.0, 2.0])
    ref = 0.0
    for i, (xi, yi) in enumerate(zip(x, y)):
        t = (-0.5 * math.log(2 * math.pi) + -1.0 * math.log(0.2)) + -0.5 * ((yi - (-1.0 + 2.0 * xi)) / 0.2) ** 2.0
        ref = t if i == 0 else ref + t
    assert prob == ref and lp[0] == 0.0


def test_prior_bounds_let_expansion(orc):
    f = orc.Fn("(lambda (params data) (prior-bounds-let ((:x -10 10) (:y 100 200) (:z 1d-4 1d-2)) "
               "(+ bounds-total (* 4 x-bound))))", ["x", "y", "z"])
    assert f.prior([0.0, 150.0, 1e-3]) == 0.0
    d = 2.0
    pen = -1e10 * (math.exp(d * 1e-5) - 1)
    assert f.prior([12.0, 150.0, 1e-3]) == pen + 4 * pen           # x-bound counted in bounds-total and 4x more
    assert f.prior([10.0, 150.0, 1e-3]) == 0.0 * 5                  # on the bound: strict < fails, distance 0 -> -0
    assert f.prior([0.0, 99.0, 1e-3]) == -1e10 * (math.exp(1.0 * 1e-5) - 1)


def test_log_prior_nv_known_answers(orc):
    """nv-specific.lisp:21-34 at points whose value follows by hand from the six bounds and three constraints"""
    from lisp_mcmc_b200 import families
    keys = ["scale1", "scale2", "mu1", "mu2", "sigma", "bg0"]
    f = orc.Fn(families.LOG_PRIOR_NV, keys)
    assert f.prior([1e-2, 1e-2, 2860.0, 2880.0, 12.0, 5e-6]) == 0.0
    assert f.prior([1e-2, 1e-2, 2869.0, 2872.0, 12.0, 5e-6]) == -1e9          # separation < 6
    assert f.prior([1.2e-2, 1e-2, 2860.0, 2880.0, 12.0, 5e-6]) == -1e9        # amplitude ratio outside (0.9, 1.1)
    v = f.prior([1e-2, 1e-2, 2860.0, 2880.0, 25.0, 5e-6])                      # sigma 5 above its bound
    assert v == -1e10 * (math.exp(5.0 * 1e-5) - 1)


def test_get_covariant_sample_uses_full_row(orc):
    L = np.array([[1.0, 2.0], [3.0, 4.0]])     # not triangular on purpose: :690-697 sums the whole row
    out = orc.get_covariant_sample([10.0, 20.0], L, [0.5, -1.0])
    assert np.array_equal(out, [(0.0 + 1.0 * 0.5 + 2.0 * -1.0) + 10.0, (0.0 + 3.0 * 0.5 + 4.0 * -1.0) + 20.0])


def test_temperature_schedule(orc):
    """:875-878 for n = 30000, temperature 10: Tn = 30000, factor 1 + 2*6 = 13 -> period 4*Tn/13 ~ 9231 steps"""
    t = orc.temperature_schedule(30000, 5, 10.0)
    assert len(t) == 30000 and t[0] == 10.0 and t.min() == 1.0
    ref = np.maximum(1.0, np.cos(np.arange(30000) * math.pi * 13 / 60000.0) * 10.0)
    assert np.max(np.abs(t - ref)) < 1e-12
    assert len(orc.temperature_schedule(100, 5, 10.0)) == 5000       # (max n (* 10 steps-to-settle))


def test_small_n_runs_the_full_tail(orc):
    """n below the 2000-step tail jumps i to n - 2000 (:906, :917): the loop still takes 2000 steps"""
    ow = orc.Walker(["b", "m"], 1)
    ow.dataset(0, [-4.0, -1, 2, 5, 10], [0.0, 2, 5, 9, 13], 0.2)
    ow.model_sexpr(0, "(lambda (x &key m b &allow-other-keys) (+ b (* m x)))")
    ow.init([-1.0, 2.0])
    assert ow.adaptive_steps_full(n=100, temperature=10) == 2000 and ow.length == 2001
    ow.init([-1.0, 2.0])
    assert ow.adaptive_steps_full(n=3000, temperature=10) == 3000


def test_acceptance_counts_prob_changes_not_flags(orc):
    """walker-get :acceptance :506-508 = runs of distinct consecutive probs / window"""
    ow = orc.Walker(["b", "m"], 1)
    ow.dataset(0, [0.0, 1.0], [0.0, 1.0], 1.0)
    ow.model_sexpr(0, "(lambda (x &key m b &allow-other-keys) (+ b (* m x)))")
    ow.init([0.0, 1.0])
    z = np.zeros((5, 2))                  # zero proposals: every step is "accepted" (equal prob -> (p1-p0)/T = 0 > ln u)
    ow.many_steps(5, np.eye(2), z=z, u=np.full(5, 0.5))
    assert ow.acceptance(6) == (1, 6)     # ... yet acceptance reads 1/6
    assert ow.length == 6 and ow.age == 6


def test_oracle_families_match_their_lisp_text(orc):
    """orc_model_builtin (used by the CPU baseline) == the family's canonical lambda through the generic evaluator"""
    from lisp_mcmc_b200 import families
    rng = np.random.default_rng(0)
    for name, deg in [("poly", 5), ("expo", 0), ("sinus", 0), ("lorentz", 0), ("gauss_lin", 0), ("nv_multi", 2)]:
        keys = families.slots(name, deg)
        fn = orc.Fn(families.lisp_text(name, deg), keys)
        for _ in range(50):
            th = rng.uniform(0.5, 2.0, len(keys))
            x = rng.uniform(-2, 2)
            assert fn.model(x, th) == orc.model_builtin(families.family_id(name), deg, th, x)


def test_linspace_is_the_exact_rational_one(orc):
    """linspace :235-247: start + n (end - start)/(len - 1) in exact rationals, rounded once per element -- not the
    floating-point recurrence numpy uses; the docstring case (linspace 0 50 :len 100) of :78 ends exactly on 50"""
    from fractions import Fraction
    xs = orc.linspace(0.0, 50.0, 100)
    assert xs[0] == 0.0 and xs[-1] == 50.0 and len(xs) == 100
    assert xs[1] == float(Fraction(50, 99)) and xs[37] == float(Fraction(50 * 37, 99))
    a, b = 0.1, 0.7                                   # (- end start) is rounded in double arithmetic first
    ys = orc.linspace(a, b, 7)
    assert ys[3] == float(Fraction(a) + 3 * Fraction(b - a) / 6)
    from lisp_mcmc_b200.walker import linspace as host_linspace
    assert np.array_equal(host_linspace(-3.7, 11.9, 1000), orc.linspace(-3.7, 11.9, 1000))


def test_data_and_fit_follows_the_list_definition(orc):
    """walker-get-data-and-fit :1231-1255 restated in C++ == the same definition spelled out on the step list in Python:
    sort the whole walk by prob (descending), keep ceil(0.66f take) entries, max / min of the model over them"""
    rng = np.random.default_rng(3)
    x = np.linspace(-4.0, 10.0, 25)
    y = 1.0 + 2.0 * x + 0.2 * rng.standard_normal(25)
    keys = ["b", "m"]
    text = "(lambda (x &key m b &allow-other-keys) (+ b (* m x)))"
    ow = orc.Walker(keys, 1)
    ow.dataset(0, x, y, 0.2)
    ow.model_sexpr(0, text)
    ow.init([-1.0, 2.0])
    ow.adaptive_steps_full(n=700, temperature=10)
    th, pr = ow.trace(0)                               # newest first
    fn = orc.Fn(text, keys)
    for take, which in [(100, "most-likely"), (150, "median"), (333, "median"), (10**6, "most-likely")]:
        xf, mx, mn, yf, par = ow.data_and_fit(take=take, which_solution=which, n_fit=40)
        tk = min(take, len(pr))
        if which == "most-likely":
            k = tk - 1 - int(np.argmax(pr[:tk][::-1]))   # of equal maxima the oldest
            ref_par = th[k]
        else:
            s = np.sort(th[:tk], axis=0)
            pos, rem = divmod(50 * (tk - 1), 100)
            ref_par = s[pos] if rem == 0 else (s[pos] + s[pos + 1]) / 2
        assert np.array_equal(par, ref_par)
        M = int(np.ceil(np.float32(0.66) * np.float32(tk)))
        order = np.argsort(-pr, kind="stable")[:M]
        ys = np.array([[fn.model(xx, th[j]) for j in order] for xx in xf])
        assert np.array_equal(mx, ys.max(axis=1)) and np.array_equal(mn, ys.min(axis=1))
        assert np.array_equal(yf, [fn.model(xx, ref_par) for xx in xf])
        assert np.array_equal(xf, orc.linspace(-4.0, 10.0, 40))
    # (* 0.66 take) is a single-float product: take = 150 gives 99.00001 -> 100 steps, where 0.66d0 would give 99
    assert int(np.ceil(np.float32(0.66) * np.float32(150))) == 100


def test_walker_modify_list_surgery_on_the_oracle(orc):
    """walker-modify :546-580 restated: the walk is newest-first, so :burn-walks (subseq walk 0 (- length n)) drops the
    OLDEST n steps, :keep-walks keeps the newest, :reset keeps (last walk) = the OLDEST step and makes it the last
    step, :reset-to-most-likely restarts from the best step; none of them touches walker-age"""
    ow = orc.Walker(["b", "m"], 1)
    ow.dataset(0, [-4.0, -1, 2, 5, 10], [0.0, 2, 5, 9, 13], 0.2)
    ow.model_sexpr(0, "(lambda (x &key m b &allow-other-keys) (+ b (* m x)))")

    def fresh():
        ow.init([-1.0, 2.0])
        rng = np.random.default_rng(1)
        ow.many_steps(40, 0.05 * np.eye(2), z=rng.standard_normal((40, 2)), u=rng.uniform(size=40))
        return ow.trace(0)

    th, pr = fresh()
    assert len(pr) == 41 and ow.age == 41
    ow.modify("burn-walks", 10)
    t2, p2 = ow.trace(0)
    assert ow.length == 31 and ow.age == 41 and np.array_equal(t2, th[:31]) and np.array_equal(p2, pr[:31])
    th, pr = fresh()
    ow.modify("keep-walks", 7)
    assert ow.length == 7 and np.array_equal(ow.trace(0)[0], th[:7])
    th, pr = fresh()
    ow.modify("reset")
    t2, p2 = ow.trace(0)
    assert ow.length == 1 and ow.age == 41 and np.array_equal(t2[0], th[-1]) and p2[0] == pr[-1]   # the FIRST step ever taken
    th, pr = fresh()
    best = ow.most_likely()
    ow.modify("reset-to-most-likely")
    t2, p2 = ow.trace(0)
    assert ow.length == 1 and np.array_equal(t2[0], best[0]) and p2[0] == best[1] == pr.max()
