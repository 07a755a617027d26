"""Synthetic inputs of the BASELINE.json configurations: host-generated with numpy default_rng,
all FP64.  Each returns a dict that both the GPU path (walker.py) and the oracle harness (tests/, bench.py) can
consume: functions, data, data_error, keys, theta_star, theta0[W,P], log_liklihood, log_prior, steps.
fp64_instr_per_point = FP64-pipe instructions the likelihood kernel executes per point-eval for that workload (counted in
its SASS with tools/sass_mix.py; C5 is compiled at run time by NVRTC and counted by hand from the lowered code)."""
import numpy as np

from . import families
from .walker import builtin

USER_POLY5_EXP = ("(lambda (x &key c0 c1 c2 c3 c4 c5 k &allow-other-keys) (* (+ c0 (* x (+ c1 (* x (+ c2 (* x (+ c3 (* x "
                  "(+ c4 (* x c5)))))))))) (exp (* -1 k x))))")


def _jitter(theta_star, W, rel, seed):
    rng = np.random.default_rng(seed)
    return theta_star[None, :] * (1.0 + rel * rng.standard_normal((W, len(theta_star))))


def c1(W=1):
    """test.lisp-style single Lorentzian, 1k points, 4 params, log-normal"""
    rng = np.random.default_rng(1001)
    x = 2000.0 + np.arange(1000, dtype=np.float64)
    th = np.array([5e-6, 2500.0, 60.0, -1e-6])
    f = th[0] * th[2] ** 2 / ((x - th[1]) ** 2 + th[2] ** 2) + th[3]
    y = f + 1e-7 * rng.standard_normal(x.shape)
    keys = ["a", "x0", "gamma", "c"]
    theta0 = np.tile(np.array([1e-6, 2450.0, 100.0, 1e-7]), (W, 1))
    return dict(name="C1", function=[builtin("lorentz", keys)], data=[[x, y]], data_error=1e-7, keys=keys,
                theta_star=th, theta0=theta0, log_liklihood="log-liklihood-normal",
                log_prior=[[("a", 0.0, 1e-4), ("x0", 2000.0, 3000.0), ("gamma", 1.0, 500.0), ("c", -1e-5, 1e-5)]],
                steps=100000, flop_per_point=9, transc_per_point={}, fp64_instr_per_point=9)


def c2(W=4096, N=100000):
    """Gaussian + linear background, 100k points, 4096 walkers, log-normal"""
    rng = np.random.default_rng(1002)
    x = np.linspace(-10.0, 10.0, N)
    th = np.array([10.0, 0.5, 1.5, 2.0, 0.1])
    sig = rng.uniform(0.05, 0.15, N)
    f = th[0] * np.exp(-0.5 * ((x - th[1]) / th[2]) ** 2) + th[3] + th[4] * x
    y = f + sig * rng.standard_normal(N)
    keys = ["a", "mu", "s", "b0", "b1"]
    return dict(name="C2", function=[builtin("gauss_lin", keys)], data=[[x, y]], data_error=[sig], keys=keys,
                theta_star=th, theta0=_jitter(th, W, 0.05, 2002), log_liklihood="log-liklihood-normal",
                log_prior="log-prior-flat", steps=2000, flop_per_point=12, transc_per_point={"exp": 1}, fp64_instr_per_point=13,
                fp64_instr_by_mode=[13, 13, 13, 12])   # guard-free, clamp-only, guarded, fused argument (profiles/r02_sass_gauss_lin.txt)


def c3(W=16384, N=1000000):
    """sinusoid with log-poisson likelihood, 1M points, 16384 walkers"""
    rng = np.random.default_rng(1003)
    x = np.linspace(0.0, 100.0, N)
    th = np.array([20.0, 0.7, 0.3, 50.0])
    lam = th[3] + th[0] * np.sin(th[1] * x + th[2])
    y = rng.poisson(lam).astype(np.float64)
    keys = ["a", "omega", "phi", "c"]
    return dict(name="C3", function=[builtin("sinus", keys)], data=[[x, y]], data_error=1.0, keys=keys,
                theta_star=th, theta0=_jitter(th, W, 0.02, 2003), log_liklihood="log-poisson",
                log_prior="log-prior-flat", steps=200, flop_per_point=7, transc_per_point={"sin": 1, "log": 1}, fp64_instr_per_point=31)


def c4(W=1024, D=8, N=1000000):
    """nv-specific double-Lorentzian global fit: shared linewidth `sigma`, per-dataset scale/mu/bg0, per-dataset
    log-prior-nv (nv-specific.lisp:21-34)"""
    keys = ["sigma"]
    for d in range(D):
        keys += ["%s-%d" % (k, d) for k in ("scale1", "scale2", "mu1", "mu2", "bg0")]
    x = np.linspace(2800.0, 2940.0, N)
    th = np.zeros(len(keys))
    th[0] = 12.0
    fns, data, priors = [], [], []
    for d in range(D):
        rng = np.random.default_rng(1004 + d)
        xi = rng.uniform(-1.0, 1.0)
        s1, s2 = 1e-2 * (1 + 0.02 * xi), 1e-2 * (1 - 0.02 * xi)
        m1, m2, bg = 2860.0 - d, 2880.0 + d, 5e-6
        th[1 + 5 * d:6 + 5 * d] = [s1, s2, m1, m2, bg]
        g2 = th[0] ** 2
        f = bg - s1 * g2 / ((x - m1) ** 2 + g2) - s2 * g2 / ((x - m2) ** 2 + g2)
        y = f + 1e-7 * rng.standard_normal(N)
        sfx = "-%d" % d
        fns.append(builtin("nv_multi", ["scale1" + sfx, "scale2" + sfx, "mu1" + sfx, "mu2" + sfx, "sigma", "bg0" + sfx], degree=2))
        data.append([x, y])
        priors.append(families.log_prior_nv_text(sfx))
    return dict(name="C4", function=fns, data=data, data_error=1e-7, keys=keys, theta_star=th,
                theta0=_jitter(th, W, 1e-4, 2004), log_liklihood="log-liklihood-normal", log_prior=priors,
                steps=200, flop_per_point=14, transc_per_point={"div": 2}, fp64_instr_per_point=14)


def c5(W=65536, N=10000000):
    """NVRTC-lowered user model: 5th-order polynomial x exponential"""
    rng = np.random.default_rng(1005)
    x = np.linspace(0.0, 2.0, N)
    th = np.array([1.0, -0.5, 0.8, 0.3, -0.2, 0.05, 1.2])
    poly = th[0] + x * (th[1] + x * (th[2] + x * (th[3] + x * (th[4] + x * th[5]))))
    y = poly * np.exp(-th[6] * x) + 0.01 * rng.standard_normal(N)
    keys = ["c0", "c1", "c2", "c3", "c4", "c5", "k"]
    return dict(name="C5", function=[USER_POLY5_EXP], data=[[x, y]], data_error=0.01, keys=keys, theta_star=th,
                theta0=_jitter(th, W, 0.01, 2005), log_liklihood="log-liklihood-normal", log_prior="log-prior-flat",
                steps=20, flop_per_point=16, transc_per_point={"exp": 1}, fp64_instr_per_point=18)


def nvset(W=131072, N=500):
    """walker set with per-walker data: one double-Lorentzian (NV) spectrum of N points per walker, common x and sigma
    (the reference's own many-walker workload, nv-specific.lisp:50-66); returns a dict for make_walker_set"""
    rng = np.random.default_rng(1006)
    x = np.linspace(2800.0, 2940.0, N)
    keys = ["scale1", "scale2", "mu1", "mu2", "sigma", "bg0"]
    th = np.array([1e-2, 1e-2, 2860.0, 2880.0, 12.0, 5e-6])
    thk = th[None, :] * (1.0 + np.array([0.05, 0.05, 1e-3, 1e-3, 0.05, 0.05])[None, :] * rng.standard_normal((W, 6)))
    g2 = thk[:, 4:5] ** 2
    f = thk[:, 5:6] - thk[:, 0:1] * g2 / ((x[None, :] - thk[:, 2:3]) ** 2 + g2) - thk[:, 1:2] * g2 / ((x[None, :] - thk[:, 3:4]) ** 2 + g2)
    Y = f + 1e-7 * rng.standard_normal((W, N))
    return dict(name="NVSET", function=[builtin("nv_multi", keys, degree=2)], data=(x, Y), data_error=1e-7, keys=keys,
                theta_star=th, theta0=thk * (1.0 + 1e-4 * rng.standard_normal((W, 6))), log_liklihood="log-liklihood-normal",
                log_prior=families.log_prior_nv_text(), steps=200, fp64_instr_per_point=14, bytes_per_point=8, flop_per_point=14, transc_per_point={"div": 2})


def make_walker_set(cfg, **gpu):
    from .walker import walker_set_create
    params = [(k, v) for k, v in zip(cfg["keys"], cfg["theta_star"])]
    return walker_set_create(function=cfg["function"][0], data=cfg["data"], params=params, data_error=cfg["data_error"],
                             log_liklihood=cfg["log_liklihood"], log_prior=cfg["log_prior"], theta0=cfg["theta0"], **gpu)


CONFIGS = {"C1": c1, "C2": c2, "C3": c3, "C4": c4, "C5": c5}


def make_walker(cfg, theta0=None, **gpu):
    """Build the GPU walker set for a workload dict."""
    from .walker import walker_create
    th0 = cfg["theta0"] if theta0 is None else theta0
    params = [(k, v) for k, v in zip(cfg["keys"], cfg["theta_star"])]
    return walker_create(function=cfg["function"], data=cfg["data"] if len(cfg["function"]) > 1 else cfg["data"][0],
                         params=params, data_error=cfg["data_error"], log_liklihood=cfg["log_liklihood"],
                         log_prior=cfg["log_prior"], theta0=th0, **gpu)
