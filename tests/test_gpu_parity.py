"""GPU parity tests proper: the CUDA path, called through the C ABI (walker.py -> libmfit_cuda.so), against the
CPU oracle on the same seeded inputs.  Tolerances are BASELINE.json's: log-likelihoods within 1e-12 relative
(FP64), identical accept/reject sequences under the same injected normal/uniform streams."""
import numpy as np
import pytest

from lisp_mcmc_b200 import builtin, families, walker_create, workloads
from lisp_mcmc_b200 import _lib
from helpers import ll_scale, oracle_walker, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-12  # relative, north_star


def _family_cfg(name, N, W, ll="log-liklihood-normal", degree=0, seed=0):
    rng = np.random.default_rng(100 + seed)
    if name == "poly":
        keys = families.slots("poly", degree)
        th = rng.uniform(-1, 1, degree + 1)
        x = np.linspace(-1.5, 1.5, N)
    elif name == "expo":
        keys, th, x = families.slots(name), np.array([2.0, -0.7, 0.5]), np.linspace(0, 4, N)
    elif name == "sinus":
        keys, th, x = families.slots(name), np.array([20.0, 0.7, 0.3, 50.0]), np.linspace(0, 100, N)
    elif name == "lorentz":
        keys, th, x = families.slots(name), np.array([5.0, 2500.0, 60.0, -1.0]), 2000.0 + np.linspace(0, 1000, N)
    elif name == "gauss_lin":
        keys, th, x = families.slots(name), np.array([10.0, 0.5, 1.5, 2.0, 0.1]), np.linspace(-10, 10, N)
    else:
        keys = families.slots("nv_multi", degree)
        th = np.array([1e-2 * (1 + 0.1 * k) for k in range(degree)] + [2860.0 + 15 * k for k in range(degree)] + [12.0, 5e-3])
        x = np.linspace(2800, 2940, N)
    fam = builtin(name, keys, degree)
    from oracle import oracle as orc
    f = np.array([orc.model_builtin(families.family_id(name), degree, th, xi) for xi in x]) if N else np.zeros(0)
    if ll == "log-poisson":
        y = rng.poisson(np.maximum(f, 0.1)).astype(np.float64)
        sig = 1.0
    else:
        sig = rng.uniform(0.05, 0.15, N) * (np.abs(th[0]) if name != "nv_multi" else 1e-3)
        y = f + sig * rng.standard_normal(N)
    theta0 = th[None, :] * (1 + 0.01 * rng.standard_normal((W, len(th))))
    return dict(function=[fam], data=[[x, y]], data_error=[sig] if np.ndim(sig) else sig, keys=keys, theta_star=th,
                theta0=theta0, log_liklihood=ll, log_prior="log-prior-flat")


FAMILIES = [("poly", 0), ("poly", 1), ("poly", 3), ("poly", 5), ("poly", 7), ("expo", 0), ("sinus", 0), ("lorentz", 0),
            ("gauss_lin", 0), ("nv_multi", 1), ("nv_multi", 2), ("nv_multi", 4)]


@pytest.mark.parametrize("name,degree", FAMILIES)
@pytest.mark.parametrize("ll", ["log-liklihood-normal", "log-liklihood-normal-cutoff"])
def test_builtin_family_loglik_matches_oracle(orc, name, degree, ll):
    cfg = _family_cfg(name, 1537, 70, ll, degree)   # ragged: 3 tiles + 1 point, 70 chains (not a multiple of 128)
    w = workloads.make_walker(cfg)
    prob, llv, lpv = w.make_step(cfg["theta0"])
    ow = oracle_walker(orc, cfg)
    ref = np.array([ow.logp(t)[0] for t in cfg["theta0"]])
    assert rel_err(prob, ref) < TOL
    assert rel_err(llv[:, 0], ref) < TOL
    assert np.all(lpv == 0.0)
    w.close()


@pytest.mark.parametrize("name", ["sinus", "expo", "gauss_lin"])
def test_log_poisson_matches_oracle(orc, name):
    cfg = _family_cfg(name, 2100, 33, "log-poisson")
    if name != "sinus":   # keep lambda positive
        cfg["theta0"] = np.abs(cfg["theta0"])
        cfg["theta0"][:, -1 if name == "expo" else 3] += 3.0
    w = workloads.make_walker(cfg)
    prob, _, _ = w.make_step(cfg["theta0"])
    ow = oracle_walker(orc, cfg)
    ref = np.array([ow.logp(t)[0] for t in cfg["theta0"]])
    assert rel_err(prob, ref) < TOL
    w.close()


def test_log_poisson_negative_rate_is_nan_without_disturbing_neighbours(orc):
    """a rate that dips below zero makes (log lambda) complex in the reference / NaN here (:383): that chain's sum is
    NaN -- decided from the guard, without a libdevice redo -- while the chains sharing its thread, warp and CTA
    (chains w and w + 128 share a thread) keep their exact values; lambda = 0 with k > 0 still takes the redo (-inf)"""
    cfg = _family_cfg("sinus", 1300, 300, "log-poisson")
    th = cfg["theta0"]
    bad = [3, 131, 200, 299]
    for b in bad:
        th[b, 3] = 0.5 * th[b, 0]          # c < a: lambda < 0 around the troughs
    w = workloads.make_walker(cfg)
    prob, _, _ = w.make_step(th)
    ow = oracle_walker(orc, cfg)
    ref = np.array([ow.logp(t)[0] for t in th])
    assert np.all(np.isnan(ref[bad])) and np.all(np.isnan(prob[bad]))
    good = np.setdiff1d(np.arange(300), bad)
    assert np.all(np.isfinite(prob[good])) and rel_err(prob[good], ref[good]) < TOL
    w.close()


@pytest.mark.parametrize("N", [0, 1, 2, 127, 128, 129, 511, 512, 513, 1024, 4097])
@pytest.mark.parametrize("W", [1, 128, 129])
def test_ragged_sizes(orc, N, W):
    cfg = _family_cfg("gauss_lin", N, W, seed=N + W)
    w = workloads.make_walker(cfg)
    prob, _, _ = w.make_step(cfg["theta0"])
    ow = oracle_walker(orc, cfg)
    ref = np.array([ow.logp(t)[0] for t in cfg["theta0"]])
    if N == 0:
        assert np.all(prob == 0.0) and np.all(ref == 0.0)   # (reduce #'+ nil) = 0
    else:
        assert rel_err(prob, ref) < TOL
    w.close()


def test_gauss_far_from_origin(orc):
    """Data far from x = 0 with peaks that are narrow next to that offset (|mu / s| from 570 to 9500): the Gaussian
    argument must be formed as (x - mu) first -- hoisting -(mu / s) into one FMA, fma(x, 1/s, -mu/s), saves an FP64
    instruction but loses these cases (1.4e-11 on this very input, measured)"""
    rng = np.random.default_rng(77)
    N, W = 3001, 260
    keys = families.slots("gauss_lin")
    th = np.array([10.0, 2869.3, 5.0, 2.0, 0.01])
    x = np.linspace(2800.0, 2940.0, N)
    from oracle import oracle as orc_mod
    f = np.array([orc_mod.model_builtin(families.family_id("gauss_lin"), 0, th, xi) for xi in x])
    sig = rng.uniform(0.05, 0.15, N)
    y = f + sig * rng.standard_normal(N)
    theta0 = th[None, :] * (1 + 0.002 * rng.standard_normal((W, 5)))
    theta0[::7, 1] += 6.0            # peaks a good sigma off the data: large residuals, full sensitivity to the argument
    narrow = [5, 128 + 5, 200, 259]
    theta0[narrow, 2] = [2.0, 1.0, 0.3, 2.2]
    cfg = dict(function=[builtin("gauss_lin", keys, 0)], data=[[x, y]], data_error=[sig], keys=keys, theta_star=th,
               theta0=theta0, log_liklihood="log-liklihood-normal", log_prior="log-prior-flat")
    w = workloads.make_walker(cfg)
    prob, _, _ = w.make_step(theta0)
    ow = oracle_walker(orc, cfg)
    ref = np.array([ow.logp(t)[0] for t in theta0])
    assert np.all(np.isfinite(prob))
    assert rel_err(prob, ref) < TOL
    w.close()


def test_cutoff_clamps_like_reference(orc):
    cfg = _family_cfg("lorentz", 700, 9, "log-liklihood-normal-cutoff")
    cfg["theta0"][:, 0] *= 50.0   # far off: most terms hit the -5000 floor (:426)
    w = workloads.make_walker(cfg)
    prob, _, _ = w.make_step(cfg["theta0"])
    ow = oracle_walker(orc, cfg)
    ref = np.array([ow.logp(t)[0] for t in cfg["theta0"]])
    assert np.all(ref <= -5000.0 * 100)   # > 100 points sit on the -5000 floor
    assert rel_err(prob, ref) < TOL
    w.close()


def test_nan_and_inf_are_returned_not_raised(orc):
    cfg = _family_cfg("expo", 600, 4)
    cfg["theta0"][1, 1] = 400.0       # exp overflow -> inf -> ll = -inf or nan
    cfg["theta0"][2, 0] = np.nan
    w = workloads.make_walker(cfg)
    prob, _, _ = w.make_step(cfg["theta0"])
    ow = oracle_walker(orc, cfg)
    ref = np.array([ow.logp(t)[0] for t in cfg["theta0"]])
    assert np.isfinite(prob[0]) and rel_err(prob[0], ref[0]) < TOL
    assert not np.isfinite(prob[1]) and not np.isfinite(ref[1])
    assert np.isnan(prob[2]) and np.isnan(ref[2])
    w.close()


def test_user_sexpr_model_matches_builtin_and_oracle(orc):
    cfg = _family_cfg("gauss_lin", 2000, 40)
    w1 = workloads.make_walker(cfg)
    p1, _, _ = w1.make_step(cfg["theta0"])
    cfg2 = dict(cfg)
    cfg2["function"] = [cfg["function"][0].lisp_text()]      # same family through lower.cpp + NVRTC
    w2 = workloads.make_walker(cfg2)
    p2, _, _ = w2.make_step(cfg["theta0"])
    ow = oracle_walker(orc, cfg2)
    ref = np.array([ow.logp(t)[0] for t in cfg["theta0"]])
    assert rel_err(p2, ref) < TOL and rel_err(p1, ref) < TOL
    w1.close(); w2.close()


def test_single_item_parameter_style_equals_keyword_parameters(orc):
    """:params '(:params (-1 2)) with (lambda (x &key params ...) (+ (elt params 0) (* (elt params 1) x))) -- the reference's
    "single param list fit" (mcmc-fitting.lisp:8, :1078-1088, :1153-1155) -- runs the same chains, bit for bit, as the
    same line written with two keyword parameters, and its log-probs agree with the oracle's evaluation of the
    single-item lambda; a vector and an n x 1 array start value behave like the list"""
    rng = np.random.default_rng(8)
    x = np.linspace(-4.0, 10.0, 700)
    y = 1.0 + 0.9 * x + 0.2 * rng.standard_normal(x.size)
    single = "(lambda (x &key params &allow-other-keys) (+ (elt params 0) (* (elt params 1) x)))"
    array = "(lambda (x &key params &allow-other-keys) (+ (aref params 0 0) (* (aref params 1 0) x)))"
    multi = "(lambda (x &key b m &allow-other-keys) (+ b (* m x)))"
    kw = dict(data=[x, y], data_error=0.2, n_walkers=5, seed=3)
    w_s = walker_create(function=single, params=[":params", [-1.0, 2.0]], **kw)
    w_v = walker_create(function=single, params=[":params", np.array([-1.0, 2.0])], **kw)
    w_a = walker_create(function=array, params=[":params", np.array([[-1.0], [2.0]])], **kw)
    w_m = walker_create(function=multi, params=[":b", -1.0, ":m", 2.0], **kw)
    assert w_s.param_style == "single-item" and w_m.param_style == "multiple-kwargs" and w_s.param_keys == ["params[0]", "params[1]"]
    for w in (w_s, w_v, w_a, w_m):
        w.adaptive_steps_full(n=600, temperature=10, auto=None)
    ref = w_m.trace(700)
    for w in (w_s, w_v, w_a):
        for a, b in zip(w.trace(700), ref):
            assert np.array_equal(a, b, equal_nan=True)
    th, pr, nv = w_s.trace(50)
    ow = orc.Walker(["params[0]", "params[1]"], 1)
    ow.dataset(0, x, y, np.full(x.size, 0.2))
    ow.model_sexpr(0, single)
    ow.likelihood(0, 0)
    ow.prior_flat(0)
    want = np.array([ow.logp(t)[0] for t in th[0, :20]])
    assert rel_err(pr[0, :20], want) < TOL
    for w in (w_s, w_v, w_a, w_m):
        w.close()


def test_c5_user_model_poly5_times_exp(orc):
    cfg = workloads.c5(W=50, N=3000)
    w = workloads.make_walker(cfg)
    prob, _, _ = w.make_step(cfg["theta0"])
    ow = oracle_walker(orc, cfg)
    ref = np.array([ow.logp(t)[0] for t in cfg["theta0"]])
    assert rel_err(prob, ref) < TOL
    w.close()


def test_prior_bounds_let_constant_bounds(orc):
    cfg = workloads.c1(W=6)
    th = cfg["theta0"].copy()
    th[1, 0] = -3e-6          # a below its lower bound
    th[2, 1] = 3000.5         # x0 above
    th[3, 2] = 1.0            # gamma exactly on the bound: strict < fails (:358)
    th[4, :] = [2e-4, 1999.0, 600.0, 2e-5]
    cfg["theta0"] = th
    w = workloads.make_walker(cfg)
    prob, ll, lp = w.make_step(th)
    ow = oracle_walker(orc, cfg)
    for c in range(6):
        rp, rll, rlp = ow.logp(th[c])
        assert rel_err(ll[c], rll) < TOL
        # ill-conditioned when active: -1e10 (exp(d 1e-5) - 1) turns one ulp of exp into 2e-6
        assert abs(lp[c, 0] - rlp[0]) <= 1e10 * 2.0 ** -50
        assert (lp[c, 0] == 0.0) == (rlp[0] == 0.0)
    w.close()


def test_nv_prior_sexpr_and_global_fit(orc):
    cfg = workloads.c4(W=20, D=3, N=1200)
    th = cfg["theta0"].copy()
    th[3, cfg["keys"].index("mu1-1")] = 2889.0     # violates mu1 < mu2 - 6 and the mu1 bound
    th[4, cfg["keys"].index("scale1-2")] *= 1.3     # ratio constraint
    th[5, 0] = 25.0                                 # shared sigma out of (9, 20): penalised by every dataset's prior
    cfg["theta0"] = th
    w = workloads.make_walker(cfg)
    prob, ll, lp = w.make_step(th)
    ow = oracle_walker(orc, cfg)
    for c in range(th.shape[0]):
        rp, rll, rlp = ow.logp(th[c])
        assert rel_err(ll[c], rll) < TOL
        assert np.all(np.abs(lp[c] - rlp) <= 1e10 * 2.0 ** -50 + 1e-12 * np.abs(rlp))
        assert abs(prob[c] - rp) <= 1e-12 * abs(rp) + 3e10 * 2.0 ** -50
    assert lp[3, 1] < -1e9 and lp[5, 0] < 0 and lp[5, 2] < 0 and lp[0].max() == 0.0
    w.close()


def _compare_chain(orc, cfg, w, z, u, run_oracle, chains, bit_exact=True):
    """Same injected streams through the GPU and through the oracle: the parameter traces must be bit-identical and the
    log-probabilities within TOL -- until, at most, a decision whose |delta log p| is itself below the tolerance
    (north_star: "identical accept/reject sequences ... except where |delta log p| falls below that tolerance"), after
    which the two chains legitimately part ways (the greedy sampler of walker-sample-region shrinks its steps until
    every proposal is such a coin toss)."""
    th, pr, nv = w.trace(int(w.stats()["length"].max()))
    Ls = w.l_matrix()
    ow = oracle_walker(orc, cfg)
    scale = ll_scale(cfg)
    compared = 0
    for c in chains:
        ow.init(cfg["theta0"][c])
        run_oracle(ow, z[c], u[c] if u is not None else None)
        oth, opr = ow.trace(0)
        n = len(opr)
        assert nv[c] == n, "chain %d took %d steps, oracle %d" % (c, nv[c], n)
        g_th, g_pr = th[c, :n][::-1], pr[c, :n][::-1]      # oldest first
        o_th, o_pr = oth[::-1], opr[::-1]
        same = np.all(g_th == o_th, axis=1)
        first = n if same.all() else int(np.argmin(same))
        assert rel_err(g_pr[:first], o_pr[:first], scale) < TOL
        compared += first
        if first == n:
            if bit_exact:
                assert np.array_equal(Ls[c], ow.last_l_matrix()), "chain %d final l-matrix differs" % c
            continue
        # the step at which they part: one side accepted the (bit-identical) proposal, the other kept the old state
        assert first >= 1
        prev_same_g = np.array_equal(g_th[first], g_th[first - 1])
        prev_same_o = np.array_equal(o_th[first], o_th[first - 1])
        assert prev_same_g != prev_same_o, "chain %d step %d: traces differ by more than one decision" % (c, first)
        proposal = o_th[first] if prev_same_g else g_th[first]
        p1 = ow.logp(proposal)[0]
        p0 = o_pr[first - 1]
        assert abs(p1 - p0) <= 4 * TOL * max(abs(p0), scale), \
            "chain %d step %d: decisions differ although |delta log p| = %g is above the tolerance" % (c, first, abs(p1 - p0))
    return compared


@pytest.fixture(params=["many-steps-per-launch", "three-launch"])
def step_path(request, monkeypatch):
    """Small problems advance many steps per launch, one CTA per chain (k_run_small); MFIT_SMALL_PATH=0 keeps them on the
    three-launch step of the large path.  The tests that take this fixture hold both against the oracle."""
    if request.param == "three-launch":
        monkeypatch.setenv("MFIT_SMALL_PATH", "0")
    return request.param


def test_many_steps_injected_stream_identical_to_oracle(orc, step_path):
    cfg = workloads.c2(W=40, N=1500)
    w = workloads.make_walker(cfg)
    n = 400
    rng = np.random.default_rng(11)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    Lm = np.diag(0.004 * np.abs(cfg["theta_star"])) + 1e-4 * np.tril(np.ones((w.P, w.P)), -1)
    w.many_steps(n, Lm, z=z, u=u)
    done = _compare_chain(orc, cfg, w, z, u, lambda ow, zz, uu: ow.many_steps(n, Lm, z=zz, u=uu), range(0, 40, 3))
    assert done == 14 * (n + 1)      # nothing borderline here: every step of every compared chain matched
    w.close()


def test_many_steps_with_one_l_matrix_per_chain(orc, step_path):
    """walker-many-steps makes its default l-matrix per walker (:851: from that walker's :median-params); on a walker set
    every chain therefore steps with its own (mfit_run_many_steps_per_chain, W x P x P) -- each chain against an oracle
    walker that was handed the same matrix, injected streams, bit-identical traces.  The host's default reproduces :851."""
    from lisp_mcmc_b200 import walker_many_steps
    cfg = workloads.c2(W=7, N=900)
    w = workloads.make_walker(cfg)
    n = 300
    rng = np.random.default_rng(12)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    Ls = np.stack([np.diag((0.002 + 0.001 * c) * np.abs(cfg["theta0"][c])) + 1e-4 * c * np.tril(np.ones((w.P, w.P)), -1)
                   for c in range(w.W)])
    w.many_steps(n, Ls, z=z, u=u)
    th, pr, nv = w.trace(n + 1)
    scale = ll_scale(cfg)
    for c in range(w.W):
        ow = oracle_walker(orc, cfg)
        ow.init(cfg["theta0"][c])
        ow.many_steps(n, Ls[c], z=z[c], u=u[c])
        oth, opr = ow.trace(0)
        assert nv[c] == len(opr) and np.array_equal(th[c, :nv[c]], oth)
        assert rel_err(pr[c, :nv[c]], opr, scale) < TOL
    med = w.get("median-params")
    before = w.stats()["length"].copy()
    walker_many_steps(w, 50)                       # default l-matrix: diag(1e-2 (single-float) x this chain's medians)
    assert np.array_equal(w.stats()["length"], before + 50)
    want = np.float64(np.float32(1e-2)) * med[:, :, None] * np.eye(w.P)[None]
    assert np.array_equal(w.l_matrix(), want)
    with pytest.raises(ValueError):
        w.many_steps(5, np.zeros((3, w.P, w.P)))
    w.close()


@pytest.mark.parametrize("auto", [None, "prob-settle"])
def test_adaptive_steps_full_injected_stream_identical_to_oracle(orc, auto, step_path):
    """walker-adaptive-steps-full end to end: temperature schedule, 200-step acceptance windows, x0.1 / x1.9 rescales,
    forward-step covariance + Cholesky re-estimates, shutdown tail and (auto) the :prob-settle stop."""
    cfg = workloads.c2(W=12, N=400)
    w = workloads.make_walker(cfg)
    n = 9000
    rng = np.random.default_rng(12)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    w.adaptive_steps_full(n=n, temperature=10, auto=auto, z=z, u=u)
    st = w.stats()
    assert st["l_updates"].min() > 3
    a = orc.AUTO_PROB_SETTLE if auto else orc.AUTO_NONE
    _compare_chain(orc, cfg, w, z, u,
                   lambda ow, zz, uu: ow.adaptive_steps_full(n=n, temperature=10, auto=a, z=zz, u=uu), range(12))
    w.close()


def test_adaptive_with_bounds_prior_lorentzian_c1(orc, step_path):
    cfg = workloads.c1(W=3)
    cfg["theta0"] = cfg["theta0"] * np.array([[1.0], [1.02], [0.97]])
    w = workloads.make_walker(cfg)
    n = 5000
    rng = np.random.default_rng(13)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    w.adaptive_steps_full(n=n, temperature=10, auto=None, z=z, u=u)
    _compare_chain(orc, cfg, w, z, u,
                   lambda ow, zz, uu: ow.adaptive_steps_full(n=n, temperature=10, auto=orc.AUTO_NONE, z=zz, u=uu), range(3))
    w.close()


def test_sample_region_and_force_take_step(orc, step_path):
    cfg = workloads.c2(W=5, N=300)
    w = workloads.make_walker(cfg)
    rng = np.random.default_rng(14)
    z = rng.standard_normal((w.W, 2999, w.P))
    w.sample_region(1e-3, z=z)
    w.force_take_step()

    def run(ow, zz, uu):
        ow.sample_region(1e-3, z=zz)
        ow.force_take_step()
    done = _compare_chain(orc, cfg, w, z, None, run, range(5))
    assert done > 5 * 200   # the chains agree bit for bit until the sampler has shrunk its steps into the noise
    st = w.stats()
    assert np.all(st["length"] == 1 + 2999 + 1) and np.all(st["age"] == st["length"])
    w.close()


def test_resume_reestimates_l_matrix_from_existing_walk(orc, step_path):
    """calling walker-adaptive-steps again on a long walk derives the l-matrix from it (:896-901)"""
    cfg = workloads.c2(W=4, N=300)
    w = workloads.make_walker(cfg)
    n = 3000
    rng = np.random.default_rng(15)
    z1, u1 = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    z2, u2 = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    w.adaptive_steps_full(n=n, temperature=10, auto=None, z=z1, u=u1)
    w.adaptive_steps_full(n=n, temperature=10, auto=None, z=z2, u=u2)
    th, pr, nv = w.trace(2 * n + 1)   # walker-create's first step + n take-steps per run
    ow = oracle_walker(orc, cfg)
    for c in range(4):
        ow.init(cfg["theta0"][c])
        ow.adaptive_steps_full(n=n, temperature=10, auto=orc.AUTO_NONE, z=z1[c], u=u1[c])
        ow.adaptive_steps_full(n=n, temperature=10, auto=orc.AUTO_NONE, z=z2[c], u=u2[c])
        oth, opr = ow.trace(0)
        assert nv[c] == len(opr)
        assert np.array_equal(th[c, :nv[c]], oth)
    w.close()


def test_philox_run_replays_through_oracle(orc, step_path):
    """on-device Philox4x32-10 proposals: export the stream the device drew and replay it through the oracle"""
    cfg = workloads.c2(W=6, N=500)
    w = workloads.make_walker(cfg, seed=1234, chain_id_base=77)
    n = 1500
    w.adaptive_steps_full(n=n, temperature=10, auto=None)
    z, u = w.philox_stream(1, 2000)   # the first take-step happens at walker age 1; n <= 2000 runs the full 2000-step tail (:906)
    assert abs(z.mean()) < 0.05 and abs(z.std() - 1) < 0.05 and 0.45 < u.mean() < 0.55 and u.min() > 0 and u.max() < 1
    _compare_chain(orc, cfg, w, z, u,
                   lambda ow, zz, uu: ow.adaptive_steps_full(n=n, temperature=10, auto=orc.AUTO_NONE, z=zz, u=uu), range(6))
    w.close()


def test_chain_sharding_is_bit_identical(orc):
    """W chains in one context == the same chains split over two contexts keyed by global chain id"""
    cfg = workloads.c2(W=10, N=700)
    n = 800
    w = workloads.make_walker(cfg, seed=99)
    w.adaptive_steps_full(n=n, temperature=10, auto=None)
    th, pr, nv = w.trace(n)
    parts = []
    from lisp_mcmc_b200 import shard_chains
    for r in range(2):
        lo, cnt = shard_chains(10, r, 2)
        wr = workloads.make_walker(cfg, theta0=cfg["theta0"][lo:lo + cnt], seed=99, chain_id_base=lo)
        wr.adaptive_steps_full(n=n, temperature=10, auto=None)
        parts.append(wr.trace(n))
        wr.close()
    th2 = np.concatenate([p[0] for p in parts]); pr2 = np.concatenate([p[1] for p in parts])
    assert np.array_equal(th, th2) and np.array_equal(pr, pr2)
    w.close()


def test_full_size_c2_properties():
    """BASELINE configs[1] at full size (100k points x 4096 chains): properties that need no oracle pass --
    chain independence (a permutation of chains permutes results bit for bit), run-to-run determinism, and
    additivity over a split of the data into two datasets of a global fit."""
    cfg = workloads.c2()
    w = workloads.make_walker(cfg)
    th = cfg["theta0"]
    p1, _, _ = w.make_step(th)
    p1b, _, _ = w.make_step(th)
    assert np.array_equal(p1, p1b)
    perm = np.random.default_rng(3).permutation(len(th))
    p2, _, _ = w.make_step(th[perm])
    assert np.array_equal(p2, p1[perm])
    w.close()
    x, y = cfg["data"][0]
    s = cfg["data_error"][0]
    h = 51200   # a whole number of 128-point tiles, so the halves cover the same segments
    cfg2 = dict(cfg)
    cfg2["function"] = [cfg["function"][0], cfg["function"][0]]
    cfg2["data"] = [[x[:h], y[:h]], [x[h:], y[h:]]]
    cfg2["data_error"] = [s[:h], s[h:]]
    w2 = workloads.make_walker(cfg2)
    q, ll, _ = w2.make_step(th[:512])
    assert rel_err(q, p1[:512], ll_scale(cfg)) < TOL   # C_d - sum/2 cancels for chains near the best fit (helpers.rel_err)
    w2.close()


def test_errors_are_codes_not_crashes():
    cfg = workloads.c2(W=4, N=100)
    with pytest.raises(_lib.MfitError):
        bad = dict(cfg); bad["function"] = ["(lambda (x &key a) (frobnicate a x))"]
        workloads.make_walker(bad)
    with pytest.raises(ValueError):
        walker_create(function=cfg["function"], data=[[1.0, 2.0], [1.0]], params=[("a", 1.0)])
    w = workloads.make_walker(cfg)
    with pytest.raises(_lib.MfitError) as e:
        w.adaptive_steps_full(n=500, auto=None, z=np.zeros((4, 10, 5)), u=np.zeros((4, 10)))
    assert e.value.code == _lib.ERR_STREAM
    w.close()


def test_multi_gpu_sharding_is_bit_identical():
    """chains and data sets sharded over every GPU of the box (tests/mgpu/mgpu_check.py under torchrun) reproduce the
    single-GPU traces bit for bit; skipped on a one-GPU box"""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = min(n, 8)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29731",
                        os.path.join(root, "tests", "mgpu", "mgpu_check.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MGPU_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_one_handle_multi_gpu_is_bit_identical():
    """mfit_multi_* (several GPUs behind one handle, one host thread per GPU inside the library; what the Lisp host's
    :devices uses): chains / data sets sharded over every GPU of the box equal the one-GPU traces bit for bit, :prob-settle
    and an in-flight stop request end on the same step everywhere (tests/mgpu/multi_check.py); skipped on a one-GPU box"""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "mgpu", "multi_check.py"), str(min(n, 8))],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "MULTI_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_one_handle_on_one_gpu_equals_the_plain_context():
    """the same fan-out with a single device listed (worker thread, offsets, twins of every call) == mfit_ctx calls"""
    cfg = workloads.c2(W=11, N=700)
    wm = workloads.make_walker(cfg, seed=4, devices=[0], shard="chains")
    w0 = workloads.make_walker(cfg, seed=4)
    for w in (wm, w0):
        w.adaptive_steps_full(n=250, temperature=10, auto=None)
    for a, b in zip(wm.trace(300), w0.trace(300)):
        assert np.array_equal(a, b, equal_nan=True)
    assert np.array_equal(wm.l_matrix(), w0.l_matrix())
    sm, s0 = wm.stats(), w0.stats()
    assert all(np.array_equal(sm[k], s0[k]) for k in sm.dtype.names)
    am, a0 = wm.analytics(200), w0.analytics(200)
    assert all(np.array_equal(am[k], a0[k], equal_nan=True) for k in a0)
    for u, v in zip(wm.data_and_fit(take=150), w0.data_and_fit(take=150)):
        assert np.array_equal(u, v, equal_nan=True)
    with pytest.raises(NotImplementedError):
        wm.sample_region()
    wm.modify("burn-walks", burn_number=40); w0.modify("burn-walks", burn_number=40)
    assert np.array_equal(wm.stats()["length"], w0.stats()["length"])
    g = workloads.c4(W=6, D=3, N=500)
    gm = workloads.make_walker(g, seed=2, devices=[0], shard="datasets")
    g0 = workloads.make_walker(g, seed=2)
    for a, b in zip(gm.make_step(g["theta0"]), g0.make_step(g["theta0"])):
        assert np.array_equal(a, b, equal_nan=True)
    for w in (wm, w0, gm, g0):
        w.close()


def test_one_handle_error_paths():
    """mfit_multi_*: bad device lists and call order are status codes with a message, not crashes; a failing call on one GPU
    names that GPU"""
    import ctypes as C
    L = _lib.lib()
    h = C.c_void_p()
    assert L.mfit_multi_create(0, None, 0, C.byref(h)) == _lib.ERR_ARG
    assert L.mfit_multi_create(2, (C.c_int * 2)(0, 0), 0, C.byref(h)) == _lib.ERR_ARG and b"twice" in L.mfit_multi_last_error(None)
    assert L.mfit_multi_create(1, (C.c_int * 1)(99), 0, C.byref(h)) == _lib.ERR_ARG and b"GPU 99" in L.mfit_multi_last_error(None)
    assert L.mfit_multi_create(1, (C.c_int * 1)(0), 7, C.byref(h)) == _lib.ERR_ARG
    assert L.mfit_multi_create(1, (C.c_int * 1)(0), _lib.SHARD_CHAINS, C.byref(h)) == 0 and h.value
    assert L.mfit_multi_size(h) == 1 and L.mfit_multi_ctx(h, 0) and not L.mfit_multi_ctx(h, 1)
    done, fin = C.c_int64(), C.c_int()
    assert L.mfit_multi_run_advance(h, 10, C.byref(done), C.byref(fin)) == _lib.ERR_STATE     # no chains yet
    assert L.mfit_multi_trace_read(h, 10, None, None, None) == _lib.ERR_STATE
    assert L.mfit_multi_problem_define(h, 1, 2) == 0
    th = np.zeros((3, 2))
    rc = L.mfit_multi_chains_init(h, th.ctypes.data_as(C.POINTER(C.c_double)), 3, 1, 0)    # no data / model yet
    assert rc < 0 and b"GPU 0 (rank 0)" in L.mfit_multi_last_error(h)
    L.mfit_multi_destroy(h)


def _nth_percentile(n, seq):
    """nth-percentile :1493-1504 restated: position n (len-1)/100; a fractional position gives the MEAN of its two
    neighbours (not an interpolation)"""
    s = np.sort(np.asarray(seq, dtype=np.float64))
    num = n * (len(s) - 1)
    pos, rem = divmod(num, 100)
    return s[pos] if rem == 0 else (s[pos] + s[pos + 1]) / 2


@pytest.mark.parametrize("take", [37, 400, 1201])
def test_trace_analytics_match_reference_definitions(orc, take):
    """walker-get :most-likely-step / :median-params / :acceptance / :covariance-matrix / :l-matrix / :stddev-params
    computed on the device over W chains at once == the reference's list definitions (:487-543) applied to the trace
    (covariance and Cholesky through the oracle's restatement of :614-643 / :583-598: bit-identical)"""
    cfg = workloads.c2(W=9, N=400)
    w = workloads.make_walker(cfg, seed=3)
    w.adaptive_steps_full(n=1200, temperature=10, auto=None)
    a = w.analytics(take)
    th, pr, nv = w.trace(take)
    for c in range(w.W):
        m = int(nv[c])
        p, t = pr[c, :m], th[c, :m]              # newest first
        # :most-likely-step: reduce keeps x only when strictly greater, so of equal maxima the oldest wins
        k = m - 1 - int(np.argmax(p[::-1]))
        assert a["most_likely_prob"][c] == p[k] and np.array_equal(a["most_likely_theta"][c], t[k])
        for i in range(w.P):
            assert a["median"][c, i] == _nth_percentile(50, t[:, i])
        changes = p[:-1].view(np.int64) != p[1:].view(np.int64)
        assert a["acceptance"][c] == (1 + np.count_nonzero(changes)) / m
        uniq = t[np.append(changes, True)]
        assert np.array_equal(a["covariance"][c], orc.lplist_covariance(uniq))
        fwd = t[:-1][p[:-1] > p[1:]]
        if len(fwd) >= 2:
            L = orc.lplist_to_l_matrix(fwd)
            assert np.array_equal(a["l_matrix"][c], L, equal_nan=True)
            assert np.array_equal(a["stddev"][c], np.diag(L), equal_nan=True)
        else:
            assert np.all(np.isnan(a["l_matrix"][c]))
    # the same through walker-get
    assert np.array_equal(w.get("median-params", take=take), a["median"])
    assert np.array_equal(w.get("acceptance", take=take), a["acceptance"])
    assert np.array_equal(w.get("most-likely-step", take=take)[1], a["most_likely_prob"])
    w.close()


def _check_data_and_fit(orc, cfg, w, chains, take, which, fn_number=0, n_fit=1000):
    """walker-get-data-and-fit :1231-1255 on the device for all chains == the oracle's restatement applied to the very
    walk the GPU holds (loaded into an oracle walker): abscissae and parameters bit-identical, model values within TOL
    of the size of the curve"""
    xf, mx, mn, yf, xd, yd, par = w.data_and_fit(take=take, fn_number=fn_number, which_solution=which, n_fit=n_fit)
    th, pr, nv = w.trace(int(w.stats()["length"].max()))
    ow = oracle_walker(orc, cfg)
    for c in chains:
        m = int(nv[c])
        ow.load_walk(th[c, :m], pr[c, :m])
        oxf, omx, omn, oyf, opar = ow.data_and_fit(take=take, fn_number=fn_number, which_solution=which, n_fit=n_fit)
        assert np.array_equal(xf, oxf)
        assert np.array_equal(par[c], opar)
        scale = float(np.max(np.abs(oyf)))
        assert rel_err(yf[c], oyf, scale) < TOL
        assert rel_err(mx[c], omx, scale) < TOL and rel_err(mn[c], omn, scale) < TOL
        assert np.all(mn[c] <= mx[c])
        # the envelope may miss the curve itself (:most-likely searches the newest `take` steps, the envelope the best
        # of the whole walk), but it always brackets the walk's most likely step
        best = th[c, int(np.argmax(pr[c, :m]))]
        ybest = np.array([orc.Fn(_fn_text(cfg, fn_number), cfg["keys"]).model(x, best) for x in xf[::97]])
        assert np.all(ybest <= mx[c][::97] + TOL * scale) and np.all(ybest >= mn[c][::97] - TOL * scale)
    return xf, mx, mn, yf, xd, yd, par


def _fn_text(cfg, d):
    f = cfg["function"][d]
    return f if isinstance(f, str) else f.lisp_text()


@pytest.mark.parametrize("take,which", [(37, "most-likely"), (400, "median"), (1000, "most-likely"), (5000, "median")])
def test_data_and_fit_envelope_matches_reference_definition(orc, take, which):
    cfg = workloads.c2(W=131, N=300)
    w = workloads.make_walker(cfg, seed=5)
    w.adaptive_steps_full(n=900, temperature=10, auto=None)
    xf, mx, mn, yf, xd, yd, par = _check_data_and_fit(orc, cfg, w, [0, 64, 130], take, which, n_fit=1000 if take == 400 else 120)
    # :x-shift / :y-shift are plain additions (:1236-1247, :1251-1252)
    xs, mxs, mns, yfs, xds, yds, _ = w.data_and_fit(take=take, which_solution=which, n_fit=len(xf), x_shift=2.5, y_shift=-1.0)
    assert np.array_equal(xs, xf + 2.5) and np.array_equal(mxs, mx - 1.0) and np.array_equal(mns, mn - 1.0)
    assert np.array_equal(yfs, yf - 1.0) and np.array_equal(xds, xd + 2.5) and np.array_equal(yds, yd - 1.0)
    w.close()


def test_data_and_fit_global_fit_and_user_model(orc):
    """:fn-number picks the data set of a global fit (its own x range and parameter slots); an NVRTC-lowered model goes
    through the same kernel body; a walk cut by walker-modify :keep-walks is searched as it stands"""
    g4 = workloads.c4(W=40, D=2, N=500)
    w = workloads.make_walker(g4, seed=9)
    w.adaptive_steps_full(n=500, temperature=10, auto=None)
    _check_data_and_fit(orc, g4, w, [0, 39], 300, "most-likely", fn_number=1, n_fit=90)
    w.close()
    c5 = workloads.c5(W=33, N=400)
    w = workloads.make_walker(c5, seed=11)
    w.adaptive_steps_full(n=450, temperature=10, auto=None)
    w.keep_walks(200)
    _check_data_and_fit(orc, c5, w, [0, 32], 150, "median", n_fit=64)
    w.close()


def _xvec_cfg(N, W, ll="log-liklihood-normal", seed=0):
    """a model of a two-component x, the walker-create docstring's own shape (:1137): x = (t, v)"""
    rng = np.random.default_rng(500 + seed)
    keys = ["a", "b", "k"]
    text = "(lambda (x &key a b k &allow-other-keys) (+ (* a (elt x 0)) (* b (exp (* -1 k (elt x 1))))))"
    th = np.array([1.5, 4.0, 0.8])
    x = np.stack([np.linspace(0.0, 3.0, N), rng.uniform(0.0, 2.0, N)], axis=1)     # N x 2
    f = th[0] * x[:, 0] + th[1] * np.exp(-th[2] * x[:, 1])
    if ll == "log-poisson":
        y, sig = rng.poisson(10 * f).astype(np.float64), 1.0
        th = th * np.array([10.0, 10.0, 1.0])
    else:
        sig = rng.uniform(0.02, 0.06, N)
        y = f + sig * rng.standard_normal(N)
    theta0 = th[None, :] * (1 + 0.01 * rng.standard_normal((W, 3)))
    return dict(function=[text], data=[[x, y]], data_error=[sig] if np.ndim(sig) else sig, keys=keys, theta_star=th,
                theta0=theta0, log_liklihood=ll, log_prior="log-prior-flat")


def _make_xvec_walker(cfg, **kw):
    params = [(k, v) for k, v in zip(cfg["keys"], cfg["theta_star"])]
    # one function, x made of sequences: the reference wants the data as a list of sets, (list (list xs ys)) :818-820
    return walker_create(function=cfg["function"], data=cfg["data"], params=params, data_error=cfg["data_error"],
                         log_liklihood=cfg["log_liklihood"], log_prior=cfg["log_prior"], theta0=cfg["theta0"], **kw)


@pytest.mark.parametrize("ll,N,W", [("log-liklihood-normal", 1537, 70), ("log-liklihood-normal-cutoff", 300, 257),
                                    ("log-poisson", 1000, 33), ("log-liklihood-normal", 1, 1)])
def test_multi_dimensional_x_loglik_matches_oracle(orc, ll, N, W):
    """x is itself a sequence (walker-create docstring :1137); every component is one more TMA-staged data
    array of the likelihood kernel, read by the lowered model as xv[k]"""
    cfg = _xvec_cfg(N, W, ll, seed=N)
    w = _make_xvec_walker(cfg)
    prob, llv, _ = w.make_step(cfg["theta0"])
    ow = oracle_walker(orc, cfg)
    ref = np.array([ow.logp(t)[0] for t in cfg["theta0"]])
    assert np.all(np.isfinite(ref)) and rel_err(prob, ref, ll_scale(cfg)) < TOL
    w.close()


def test_multi_dimensional_x_run_identical_to_oracle_and_next_to_a_scalar_set(orc):
    """a global fit of one data set with a two-component x and one with a plain x, sharing parameters; injected streams"""
    cx = _xvec_cfg(700, 20, seed=3)
    rng = np.random.default_rng(8)
    x1 = np.linspace(0.0, 2.0, 450)
    s1 = 0.05
    y1 = cx["theta_star"][1] * np.exp(-cx["theta_star"][2] * x1) + s1 * rng.standard_normal(450)
    f1 = "(lambda (x &key b k &allow-other-keys) (* b (exp (* -1 k x))))"
    cfg = dict(function=[cx["function"][0], f1], data=[cx["data"][0], [x1, y1]], data_error=[cx["data_error"][0], np.full(450, s1)],
               keys=cx["keys"], theta_star=cx["theta_star"], theta0=cx["theta0"], log_liklihood="log-liklihood-normal",
               log_prior="log-prior-flat")
    w = workloads.make_walker(cfg)
    n = 2200                                  # (the loop never takes fewer than 2000 steps, :906 / :917)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    w.adaptive_steps_full(n=n, temperature=10, auto=None, z=z, u=u)
    _compare_chain(orc, cfg, w, z, u,
                   lambda ow, zz, uu: ow.adaptive_steps_full(n=n, temperature=10, auto=orc.AUTO_NONE, z=zz, u=uu), [0, 7, 19])
    w.close()


def test_multi_dimensional_x_type_errors():
    """what the reference would signal inside the model: a sequence used as a number, a number indexed, an index beyond
    the end of x, a built-in family (all take a number) on sequence data"""
    cfg = _xvec_cfg(64, 4)
    bad = dict(cfg)
    bad["function"] = ["(lambda (x &key a b k &allow-other-keys) (+ a (* b x) k))"]          # x as a number
    with pytest.raises(_lib.MfitError) as e:
        _make_xvec_walker(bad)
    assert e.value.code == _lib.ERR_ARG
    bad["function"] = ["(lambda (x &key a b k &allow-other-keys) (+ a (* b (elt x 2)) k))"]  # third of two components
    with pytest.raises(_lib.MfitError) as e:
        _make_xvec_walker(bad)
    assert e.value.code == _lib.ERR_ARG
    flat = dict(cfg)
    flat["data"] = [[cfg["data"][0][0][:, 0], cfg["data"][0][1]]]                              # plain x, indexing model
    with pytest.raises(_lib.MfitError) as e:
        _make_xvec_walker(flat)
    assert e.value.code == _lib.ERR_ARG
    fam = dict(cfg)
    fam["function"] = [builtin("poly", ["a", "b", "k"], 2)]
    with pytest.raises(_lib.MfitError) as e:
        _make_xvec_walker(fam)
    assert e.value.code == _lib.ERR_ARG


def test_data_and_fit_with_walks_of_different_lengths(orc):
    """after walker-adaptive-steps (:auto :prob-settle) the chains hold walks of different lengths: take is clamped per
    chain (:1233-1235) and the 66 % count follows it"""
    cfg = workloads.c2(W=70, N=250)
    w = workloads.make_walker(cfg, seed=13)
    w.adaptive_steps(n=6000)
    lens = w.stats()["length"]
    assert lens.min() < lens.max()
    order = np.argsort(lens)
    _check_data_and_fit(orc, cfg, w, [int(order[0]), int(order[len(order) // 2]), int(order[-1])], int(lens.min()) + 50,
                        "most-likely", n_fit=60)
    w.close()


def test_restore_into_a_smaller_fixed_ring_keeps_the_newest_steps():
    """mfit_chains_restore with a trace ring that cannot hold the whole saved walk (trace_capacity fixed by the caller):
    the oldest steps drop out, like :max-walker-length; last step, age and most likely step are still the saved ones"""
    cfg = workloads.c2(W=33, N=200)
    a = workloads.make_walker(cfg, seed=2)
    a.many_steps(1500, np.diag(0.002 * np.abs(cfg["theta_star"])))
    th, pr, nv = a.trace(2000)
    assert np.all(nv == 1501)
    b = workloads.make_walker(cfg, seed=2, trace_capacity=1100)
    b.restore(th[:, :1501], pr[:, :1501], nv, a.stats()["age"], *a.most_likely())
    th2, pr2, nv2 = b.trace(2000)
    assert np.all(nv2 == 1100)
    assert np.array_equal(th2[:, :1100], th[:, :1100]) and np.array_equal(pr2[:, :1100], pr[:, :1100])
    assert np.array_equal(b.last_step()[0], th[:, 0]) and np.array_equal(b.stats()["age"], a.stats()["age"])
    assert np.array_equal(b.most_likely()[1], a.most_likely()[1])
    a.close(); b.close()


def test_two_stream_chain_groups_do_not_change_results(orc):
    """a run advances the chains as two groups on two streams (the per-chain kernels of one overlap the likelihood
    kernel of the other): bit-identical to the one-group run, with built-in and NVRTC priors, Philox and injected
    streams, and still identical to the oracle"""
    cfg = workloads.c1(W=600)                      # 3 chain-blocks -> groups of 2 and 1; constant-bounds prior
    cfg["theta0"] = cfg["theta0"] * (1 + 0.01 * np.random.default_rng(1).standard_normal(cfg["theta0"].shape))
    n = 450
    out = []
    for g in (1, 2):
        w = workloads.make_walker(cfg, seed=17)
        w.set_groups(g)
        w.adaptive_steps_full(n=n, temperature=10, auto=None)
        out.append(w.trace(2001) + (w.l_matrix(),))
        w.close()
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(a, b, equal_nan=True)
    # a global fit with lowered (NVRTC) priors and injected streams, against the oracle
    g4 = workloads.c4(W=520, D=2, N=700)
    rng = np.random.default_rng(2)
    z, u = rng.standard_normal((520, 2000, len(g4["keys"]))), rng.uniform(size=(520, 2000))
    w = workloads.make_walker(g4)
    w.set_groups(2)
    w.adaptive_steps_full(n=300, temperature=10, auto=None, z=z, u=u)
    _compare_chain(orc, g4, w, z, u,
                   lambda ow, zz, uu: ow.adaptive_steps_full(n=300, temperature=10, auto=orc.AUTO_NONE, z=zz, u=uu),
                   [0, 255, 256, 519])
    w.close()


@pytest.mark.parametrize("opt", ["max-walker-length", "best-value", "l-matrix"])
def test_adaptive_driver_options_identical_to_oracle(orc, opt):
    """the remaining keyword arguments of walker-adaptive-steps-full (:862): :max-walker-length (the walk is cut to
    half of it at the reset index, :868 / :923-927), :sampling-optimization :best-value (:895) and a caller-supplied
    :l-matrix (:896)"""
    cfg = workloads.c2(W=6, N=300)
    w = workloads.make_walker(cfg)
    n = 12500 if opt == "max-walker-length" else 2600
    rng = np.random.default_rng(21)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    kw, okw = {}, {}
    if opt == "max-walker-length":
        kw = okw = dict(max_walker_length=3000)
    elif opt == "best-value":
        kw, okw = dict(sampling_optimization="best-value"), dict(sampling_optimization=orc.OPT_BEST_VALUE)
    else:
        Lm = np.diag(1e-3 * np.abs(cfg["theta_star"])) + 1e-4 * np.tril(np.ones((w.P, w.P)), -1)
        kw = okw = dict(l_matrix=Lm)
    w.adaptive_steps_full(n=n, temperature=10, auto=None, z=z, u=u, **kw)
    _compare_chain(orc, cfg, w, z, u,
                   lambda ow, zz, uu: ow.adaptive_steps_full(n=n, temperature=10, auto=orc.AUTO_NONE, z=zz, u=uu, **okw),
                   range(6))
    if opt == "max-walker-length":
        assert np.all(w.stats()["length"] < 3000 + 1600)   # the walk really was cut
    w.close()


def test_estop_slices_and_keep_walks(orc):
    """mfit-walker-estop (:860, :904): a run issued in slices and stopped after 700 steps == the oracle's loop interrupted
    after 700 take-steps; walker-modify :keep-walks (:568-569) then limits what walker-get :steps returns"""
    cfg = workloads.c2(W=4, N=300)
    w = workloads.make_walker(cfg)
    n = 3000
    rng = np.random.default_rng(22)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    import ctypes as C
    L, h = w._L, w._h
    cfgc = _lib.AdaptCfg(n, 10.0, _lib.AUTO_NONE, _lib.OPT_COVARIANCE, 0, 0, 0)
    zp, up = np.ascontiguousarray(z), np.ascontiguousarray(u)
    dp = C.POINTER(C.c_double)
    w._chk(L.mfit_run_begin(h, C.byref(cfgc), None, zp.ctypes.data_as(dp), up.ctypes.data_as(dp), n))
    done, fin, total = C.c_int64(), C.c_int(), 0
    for k in (100, 250, 350):
        w._chk(L.mfit_run_advance(h, k, C.byref(done), C.byref(fin)))
        total += done.value
        assert not fin.value
    w.request_stop()
    w._chk(L.mfit_run_advance(h, 500, C.byref(done), C.byref(fin)))
    assert done.value == 0 and fin.value == 1 and total == 700
    _compare_chain(orc, cfg, w, z, u,
                   lambda ow, zz, uu: ow.adaptive_steps_full(n=n, temperature=10, auto=orc.AUTO_NONE, z=zz, u=uu, max_steps=700),
                   range(4))
    w.keep_walks(50)
    th, pr, nv = w.trace(200)
    assert np.all(nv == 50) and np.all(w.stats()["length"] == 50)
    w.close()


@pytest.mark.parametrize("name,ll,shared_x", [("nv_multi", "log-liklihood-normal", True), ("gauss_lin", "log-liklihood-normal", False),
                                             ("sinus", "log-poisson", True), ("lorentz", "log-liklihood-normal-cutoff", True)])
def test_walker_set_with_per_walker_data_matches_oracle(orc, name, ll, shared_x):
    """the reference's own many-walker workload (nv-specific.lisp:50-66): one walker per spectrum, same model / prior,
    every walker its own y (and x, sigma).  GPU: one warp per walker; oracle: W independent walkers.  Likelihoods within
    1e-12, injected-stream runs bit-identical in the parameters."""
    from lisp_mcmc_b200 import walker_set_create
    W, n = 37, 333
    degree = 2 if name == "nv_multi" else 0
    base = _family_cfg(name, n, W, ll, degree, seed=5)
    rng = np.random.default_rng(31)
    th = base["theta_star"]
    fam = base["function"][0]
    from oracle import oracle as o
    xs = np.stack([base["data"][0][0] * (1.0 if shared_x else 1 + 0.01 * k) for k in range(W)])
    thk = th[None, :] * (1 + 0.02 * rng.standard_normal((W, len(th))))          # every spectrum its own truth
    f = np.array([[o.model_builtin(families.family_id(name), degree, thk[k], xi) for xi in xs[k]] for k in range(W)])
    if ll == "log-poisson":
        Y, sig = rng.poisson(np.maximum(f, 0.1)).astype(np.float64), 1.0
    else:
        sig = rng.uniform(0.05, 0.15, (W, n)) * (np.abs(th[0]) if name != "nv_multi" else 1e-3)
        Y = f + sig * rng.standard_normal((W, n))
    keys = base["keys"]
    params = [(k, v) for k, v in zip(keys, th)]
    prior = [(keys[0], -1e3, 1e3)]
    w = walker_set_create(function=fam, data=(xs[0], Y) if shared_x else [(xs[k], Y[k]) for k in range(W)], params=params,
                          data_error=sig, log_liklihood=ll, log_prior=prior, theta0=thk)
    prob, _, _ = w.make_step(thk)

    def oracle_for(k):
        cfgk = dict(function=[fam], data=[[xs[k], Y[k]]], data_error=[sig[k]] if np.ndim(sig) else sig, keys=keys,
                    log_liklihood=ll, log_prior=[prior])
        return oracle_walker(orc, cfgk), cfgk
    ref = np.array([oracle_for(k)[0].logp(thk[k])[0] for k in range(W)])
    scale = max(ll_scale(oracle_for(k)[1]) for k in range(W)) if ll != "log-poisson" else 0.0
    assert rel_err(prob, ref, scale) < TOL
    nsteps = 250
    z, u = rng.standard_normal((W, 2000, len(th))), rng.uniform(size=(W, 2000))
    w.adaptive_steps_full(n=nsteps, temperature=10, auto=None, z=z, u=u)
    th_g, pr_g, nv = w.trace(2001)
    for k in (0, 17, 36):
        ow, _ = oracle_for(k)
        ow.init(thk[k])
        ow.adaptive_steps_full(n=nsteps, temperature=10, auto=orc.AUTO_NONE, z=z[k], u=u[k])
        oth, opr = ow.trace(0)
        assert nv[k] == len(opr)
        same = np.all(th_g[k, :nv[k]] == oth, axis=1)
        assert same[::-1].cumprod().sum() > 0.9 * len(opr)   # bit-identical until, at most, a borderline decision
    w.close()


@pytest.mark.parametrize("W,n", [(300, 500), (20011, 96), (5, 1)])
def test_walker_set_tma_ring_matches_register_path_and_oracle(orc, W, n):
    """walker sets with a common x and sigma (the NV workload's shape): every warp streams its chains' y rows through
    its own TMA ring in shared memory (mfit_perchain_ring).  Bit-identical to the path that reads the rows straight into
    registers (MFIT_NO_PC_RING), also when a warp works through many chains (more chains than resident warps x ring
    stages) and when some chains have finished; within 1e-12 of the oracle"""
    import os
    ns = workloads.nvset(W=W, N=n)
    w = workloads.make_walker_set(ns, seed=3)
    th0 = ns["theta0"]
    p_ring = w.make_step(th0)[0]
    os.environ["MFIT_NO_PC_RING"] = "1"
    try:
        p_reg = w.make_step(th0)[0]
    finally:
        del os.environ["MFIT_NO_PC_RING"]
    assert np.array_equal(p_ring, p_reg)
    # rows of up to 512 points with a uniform sigma take the register form of the ring kernel (x in registers, lane sums of
    # eight chains combined together); MFIT_NO_PC_REG keeps them on the shared-memory form: the same bits
    os.environ["MFIT_NO_PC_REG"] = "1"
    try:
        p_smem = w.make_step(th0)[0]
    finally:
        del os.environ["MFIT_NO_PC_REG"]
    assert np.array_equal(p_ring, p_smem)
    if W == 300:   # a sigma VECTOR shared by the walkers (not uniform): shared-memory form, against the register-reading path
        ns2 = dict(ns)
        ns2["data_error"] = 1e-7 * (1.0 + 0.3 * np.sin(np.arange(n)))
        w3 = workloads.make_walker_set(ns2, seed=3)
        p_a = w3.make_step(th0)[0]
        os.environ["MFIT_NO_PC_RING"] = "1"
        try:
            p_b = w3.make_step(th0)[0]
        finally:
            del os.environ["MFIT_NO_PC_RING"]
        assert np.array_equal(p_a, p_b) and not np.array_equal(p_a, p_ring)
        w3.close()
    x, Y = ns["data"]
    for k in sorted({0, W // 2, W - 1}):
        cfgk = dict(function=ns["function"], data=[[x, Y[k]]], data_error=ns["data_error"], keys=ns["keys"],
                    log_liklihood=ns["log_liklihood"], log_prior=ns["log_prior"])
        ref = oracle_walker(orc, cfgk).logp(th0[k])[0]
        assert abs(p_ring[k] - ref) <= TOL * max(abs(ref), ll_scale(cfgk))
    if W == 300:   # a run: chains that stop early (:prob-settle) are skipped by the kernel while the others go on
        w.adaptive_steps(n=2500)
        a = w.trace(4000)
        w2 = workloads.make_walker_set(ns, seed=3)
        os.environ["MFIT_NO_PC_RING"] = "1"
        try:
            w2.adaptive_steps(n=2500)
        finally:
            del os.environ["MFIT_NO_PC_RING"]
        b = w2.trace(4000)
        for u, v in zip(a, b):
            assert np.array_equal(u, v, equal_nan=True)
        w2.close()
    w.close()


@pytest.mark.parametrize("modify,number", [("burn-walks", 120), ("reset", 0), ("reset-to-most-likely", 0)])
def test_walker_modify_list_surgery(orc, modify, number):
    """walker-modify :566-578 on every chain == the oracle's list surgery: :burn-walks drops the OLDEST steps, :reset keeps
    the OLDEST step ((last walk) of a newest-first list) and makes it the last step, :reset-to-most-likely restarts from
    the best step; a run taken up afterwards stays identical"""
    cfg = workloads.c2(W=130, N=300)
    w = workloads.make_walker(cfg)
    rng = np.random.default_rng(31)
    n = 300
    z, u = rng.standard_normal((w.W, 2 * n, w.P)), rng.uniform(size=(w.W, 2 * n))
    Lm = np.diag(0.003 * np.abs(cfg["theta_star"]))
    w.many_steps(n, Lm, z=z[:, :n], u=u[:, :n])
    w.modify(modify, burn_number=number)
    st = w.stats()
    th, pr, nv = w.trace(1000)
    lth, lpr = w.last_step()
    ows = {}
    for c in (0, 64, 129):
        ow = oracle_walker(orc, cfg)
        ow.init(cfg["theta0"][c])
        ow.many_steps(n, Lm, z=z[c, :n], u=u[c, :n])
        ow.modify(modify, number)
        oth, opr = ow.trace(0)
        assert st["length"][c] == ow.length == len(opr) == nv[c] and st["age"][c] == ow.age
        assert np.array_equal(th[c, :nv[c]], oth) and rel_err(pr[c, :nv[c]], opr, ll_scale(cfg)) < TOL
        assert np.array_equal(lth[c], oth[0])
        ows[c] = ow
    w.many_steps(n, Lm, z=z[:, n:], u=u[:, n:])
    th, pr, nv = w.trace(1000)
    for c, ow in ows.items():
        ow.many_steps(n, Lm, z=z[c, n:], u=u[c, n:])
        oth, opr = ow.trace(0)
        assert nv[c] == len(opr) and np.array_equal(th[c, :nv[c]], oth)
    w.close()


def test_walker_save_and_load_round_trip(tmp_path):
    """walker-save / walker-load (:971-1001, dead code in the reference, revived): an s-expression header in the layout
    of the reference's dump followed by the FP64 trace, newest step first"""
    from lisp_mcmc_b200 import walker_load, walker_save
    cfg = workloads.c2(W=5, N=200)
    w = workloads.make_walker(cfg, seed=8)
    w.adaptive_steps_full(n=300, temperature=10, auto=None)
    path = walker_save(w, str(tmp_path / "walk.mfit"))
    th, pr, nv = w.trace(int(w.stats()["length"].max()))
    got = walker_load(path)
    assert got["param_keys"] == cfg["keys"] and got["header"].startswith("(:mfit-walker 2 :fn (")
    assert np.array_equal(got["n_valid"], nv)
    assert np.array_equal(got["theta"], th, equal_nan=True) and np.array_equal(got["prob"], pr, equal_nan=True)
    assert ":log-liklihood (\"log-liklihood-normal\")" in got["header"] and ":n-walkers 5" in got["header"]
    assert np.array_equal(got["age"], w.stats()["age"]) and got["seed"] == 8
    bth, bpr = w.most_likely()
    assert np.array_equal(got["best_theta"], bth) and np.array_equal(got["best_prob"], bpr)
    w.close()


def test_walker_save_load_multiline_lambda(tmp_path):
    """:function and :log-prior given as multi-line Lisp source (the usual form) survive the header round trip"""
    from lisp_mcmc_b200 import walker_load, walker_save
    cfg = workloads.c2(W=5, N=200)
    fn = """(lambda (x &key a mu s b0 b1 &allow-other-keys)
  ;; gaussian on a line, "quoted" text in a comment
  (+ (* a (exp (* -1/2 (expt (/ (- x mu) s) 2))))
     b0
     (* b1 x)))"""
    prior = """(lambda (params data)
  (prior-bounds-let ((:a 0 100)
                     (:s 0 10))
    bounds-total))"""
    a = walker_create(function=fn, data=cfg["data"][0], params=list(zip(cfg["keys"], cfg["theta_star"])),
                      data_error=cfg["data_error"], log_prior=prior, theta0=cfg["theta0"], seed=3)
    a.adaptive_steps_full(n=300, temperature=10, auto=None)
    path = walker_save(a, str(tmp_path / "ml.mfit"))
    got = walker_load(path)
    assert "gaussian on a line" in got["header"] and "bounds-total" in got["header"]
    assert got["theta"].shape == (5, int(a.stats()["length"].max()), 5)
    b = walker_load(path, function=fn, data=cfg["data"][0], data_error=cfg["data_error"], log_prior=prior)
    assert np.array_equal(a.trace(301)[0], b.trace(301)[0]) and np.array_equal(a.stats()["age"], b.stats()["age"])
    a.close(); b.close()


def test_keep_walks_never_grows_a_walk():
    """(subseq walk 0 keep) on a shorter walk is an error in the reference; here the walk is left alone"""
    cfg = workloads.c2(W=3, N=100)
    w = workloads.make_walker(cfg)
    w.adaptive_steps_full(n=300, temperature=10, auto=None)
    w.modify("reset")
    assert np.all(w.stats()["length"] == 1)
    w.keep_walks(50)
    assert np.all(w.stats()["length"] == 1)
    w.close()


@pytest.mark.parametrize("keep", [None, 700])
def test_walker_load_resumes_exactly(tmp_path, keep):
    """a walker set rebuilt by walker-load (:989-1001: walker-create on the newest step, then :add-walks) continues bit
    for bit like the one that was saved -- same walk, ages (the Philox key), most likely steps -- through a fixed-L run
    and through another adaptive run (whose first l-matrix comes from the restored walk); also after :keep-walks, when
    the most likely step may have left the walk"""
    from lisp_mcmc_b200 import walker_load, walker_save
    cfg = workloads.c2(W=140, N=300)
    a = workloads.make_walker(cfg, seed=21, chain_id_base=1000)
    a.adaptive_steps_full(n=2100, temperature=10, auto=None)
    if keep:
        a.keep_walks(keep)
    path = walker_save(a, str(tmp_path / "ckpt.mfit"))
    b = walker_load(path, function=cfg["function"], data=cfg["data"][0], data_error=cfg["data_error"],
                    log_liklihood=cfg["log_liklihood"], log_prior=cfg["log_prior"])
    sa, sb = a.stats(), b.stats()
    for k in ("length", "age", "last_prob", "best_prob"):
        assert np.array_equal(sa[k], sb[k]), k
    assert np.array_equal(a.most_likely()[0], b.most_likely()[0])
    Lm = np.diag(0.002 * np.abs(cfg["theta_star"]))
    for w in (a, b):
        w.many_steps(150, Lm)
        w.adaptive_steps_full(n=2100, temperature=10, auto=None)
    ta, tb = a.trace(5000), b.trace(5000)
    for x, y in zip(ta, tb):
        assert np.array_equal(x, y, equal_nan=True)
    assert np.array_equal(a.l_matrix(), b.l_matrix())
    assert ta[2].min() >= 2100 + 150                   # every chain really ran on
    a.close(); b.close()


@pytest.mark.parametrize("text,strict", [
    # Cauchy / Student-t shaped term with a fixed scale: nothing of it exists as a built-in kind
    ("(lambda (y model error) (- (log (+ 1 (expt (/ (- y model) 0.5d0) 2)))))", True),
    # log-normal with the scale taken from the error LIST the way the reference hands it over (:416)
    ("(lambda (y model error) (log-normal y model (elt error 3)))", True),
    # log-poisson inside a larger expression: the per-point path with the log-factorial table of y
    ("(create-log-liklihood-function (lambda (y model error) (* 1 (log-poisson model y))))", True),
    ("(lambda (y model error) (- (* y (log model)) model (log-factorial y)))", True)])
@pytest.mark.parametrize("fn", ["builtin", "sexpr"])
def test_user_likelihood_lambda_matches_oracle(orc, text, strict, fn):
    """create-log-liklihood-function :402-417 with an arbitrary per-point lambda: lowered, compiled with the data set's
    model (built-in family or lowered :function) and summed by the same kernel -- against the oracle's evaluator"""
    from lisp_mcmc_b200 import create_log_liklihood_function
    cfg = _family_cfg("expo", 2777, 70, "log-poisson")
    cfg["theta0"] = np.abs(cfg["theta0"])
    cfg["theta0"][:, -1] += 3.0                      # keep the rate positive
    if fn == "sexpr":
        cfg["function"] = [cfg["function"][0].lisp_text()]
    cfg["data_error"] = [np.linspace(0.2, 0.9, 2777)]
    cfg["log_liklihood"] = text
    w = workloads.make_walker(cfg)
    prob, llv, _ = w.make_step(cfg["theta0"])
    cfg_o = dict(cfg)
    cfg_o["log_liklihood"] = create_log_liklihood_function(text)
    ow = oracle_walker(orc, cfg_o)
    ref = np.array([ow.logp(t)[0] for t in cfg["theta0"]])
    assert np.all(np.isfinite(ref))
    assert rel_err(prob, ref) < TOL
    # and through a short run: accept / reject sequences identical to the oracle's under injected streams
    n = 120
    rng = np.random.default_rng(31)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    Lm = np.diag(0.01 * np.abs(cfg["theta_star"]))
    w.many_steps(n, Lm, z=z, u=u)
    _compare_chain(orc, cfg_o, w, z, u, lambda o, zz, uu: o.many_steps(n, Lm, z=zz, u=uu), [0, 33, 69])
    w.close()


def test_log_poisson_spelled_as_a_lambda_is_the_builtin_kind(orc):
    """(create-log-liklihood-function (lambda (y model error) (log-poisson model y))) -- how the reference spells a
    Poisson likelihood -- must equal the built-in kind bit for bit"""
    cfg = _family_cfg("sinus", 1900, 40, "log-poisson")
    w1 = workloads.make_walker(cfg)
    p1 = w1.make_step(cfg["theta0"])[0]
    cfg2 = dict(cfg)
    cfg2["log_liklihood"] = "(lambda (count rate sigmas) (declare (ignorable sigmas)) (log-poisson rate count))"
    w2 = workloads.make_walker(cfg2)
    p2 = w2.make_step(cfg["theta0"])[0]
    assert np.array_equal(p1, p2)
    ow = oracle_walker(orc, cfg2)
    assert rel_err(p2, np.array([ow.logp(t)[0] for t in cfg["theta0"]])) < TOL
    w1.close(); w2.close()


@pytest.mark.parametrize("name,degree,ll", [("poly", 3, "log-liklihood-normal-cutoff"), ("sinus", 0, "log-poisson"),
                                            ("expo", 0, "log-liklihood-normal"), ("nv_multi", 2, "log-liklihood-normal"),
                                            ("lorentz", 0, "log-liklihood-normal")])
def test_small_path_every_family_identical_to_oracle(orc, name, degree, ll):
    """k_run_small (many steps per launch, one CTA per chain) for each built-in family and likelihood kind, bounds prior
    included: parameter traces bit-identical to the oracle under injected streams, log p within tolerance"""
    cfg = _family_cfg(name, 700, 9, ll, degree, seed=5)
    if name == "sinus":
        cfg["theta0"][:, 3] = np.abs(cfg["theta0"][:, 3]) + 25.0
    k0 = cfg["keys"][0]
    cfg["log_prior"] = [[(k0, -1e3, 1e3)]]
    w = workloads.make_walker(cfg)
    w.set_small_path(2)
    n = 700
    rng = np.random.default_rng(41)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    w.adaptive_steps_full(n=n, temperature=10, auto=None, z=z, u=u, max_steps=n)
    _compare_chain(orc, cfg, w, z, u,
                   lambda ow, zz, uu: ow.adaptive_steps_full(n=n, temperature=10, auto=orc.AUTO_NONE, z=zz, u=uu, max_steps=n),
                   range(9), bit_exact=False)
    assert np.all(w.stats()["steps_this_run"] == n)
    w.close()


def test_small_path_global_fit_and_auto_stop(orc):
    """two data sets sharing a parameter, :auto :prob-settle: the chains leave the loop on their own, at the oracle's step"""
    cfg = workloads.c4(W=6, D=2, N=500)
    cfg["log_prior"] = [[("sigma", 9.0, 20.0)], "log-prior-flat"]
    w = workloads.make_walker(cfg)
    w.set_small_path(2)
    n = 6000
    rng = np.random.default_rng(43)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    w.adaptive_steps_full(n=n, temperature=10, auto="prob-settle", z=z, u=u)
    _compare_chain(orc, cfg, w, z, u,
                   lambda ow, zz, uu: ow.adaptive_steps_full(n=n, temperature=10, auto=orc.AUTO_PROB_SETTLE, z=zz, u=uu), range(6))
    w.close()
