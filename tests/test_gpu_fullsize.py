"""GPU parity at the sizes bench.py runs (BASELINE.json configs[1..4]): the CUDA path through the C ABI against the CPU
oracle at the SHIPPED geometry -- several tiles per work item, hundreds to a thousand segments per data set, the fused
eight-data-set launch of a global fit, the MFIT_MAX_SEG clamp of C5 -- not only at the few-tile sizes of
test_gpu_parity.py.  The segment geometry is a function of the point count alone, so a handful of chains at full N
exercises exactly the sums the benchmark forms for thousands.

Every comparison reports two numbers (written to gpurun_out/fullsize_parity.json when that directory exists, copied to
profiles/ by hand): `strict` = max |gpu - ref| / |ref| (north_star's 1e-12, asserted), and `scaled` = the same relative
to the size of the terms that were summed (helpers.rel_err).  `ref_order_vs_exact` is how far the REFERENCE's own
left-to-right sum (mcmc-fitting.lisp:400, `reduce #'+`) lies from a compensated (Neumaier) sum of the same terms: the
yardstick for what 1e-12 means at N = 1e5 ... 1e7."""
import json
import os

import numpy as np
import pytest

from lisp_mcmc_b200 import _lib, workloads
from helpers import ll_scale, oracle_walker, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-12
_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_REPORT = {}


def _record(name, **kw):
    _REPORT[name] = {k: (float(v) if isinstance(v, (np.floating, float)) else v) for k, v in kw.items()}
    out = os.path.join(_ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "fullsize_parity.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)


def _strict(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def _check_logp(orc, cfg, name, chains, builtin_formula=False, compensated=True):
    """log p of the chains `chains` of cfg["theta0"]: GPU (all of theta0 in one batch) vs oracle (those chains)"""
    w = workloads.make_walker(cfg)
    prob, ll, lp = w.make_step(cfg["theta0"])
    seg = [int(v) for v in w.segment_geometry()]
    w.close()
    ow = oracle_walker(orc, cfg, use_builtin_formula=builtin_formula)
    ref = np.array([ow.logp(cfg["theta0"][c])[0] for c in chains])
    strict = _strict(prob[chains], ref)
    scaled = rel_err(prob[chains], ref, ll_scale(cfg))
    rec = dict(strict=strict, scaled=scaled, chains=len(chains), geometry=seg)
    if compensated:
        exact = np.array([ow.logp(cfg["theta0"][c], orc.SUM_COMPENSATED)[0] for c in chains[:2]])
        rec["ref_order_vs_exact"] = _strict(ref[:2], exact)
        rec["gpu_vs_exact"] = _strict(prob[chains[:2]], exact)
    _record(name, **rec)
    assert np.all(np.isfinite(ref) == np.isfinite(prob[chains]))
    assert strict < TOL, "%s: strict relative error %g" % (name, strict)
    return prob


def _c2_mixed(W=512):
    """C2 at full N with the chains a run actually meets: near the best fit (theta0 jitter), exactly theta*, and the
    wild proposals of the reference's start-up phase (its first l-matrix is diag(theta): sigma_i = |theta_i|)."""
    cfg = workloads.c2(W=W)
    rng = np.random.default_rng(42)
    th = cfg["theta0"]
    th[1] = cfg["theta_star"]
    wild = np.arange(2, W, 5)
    th[wild] = cfg["theta_star"][None, :] * (1.0 + rng.standard_normal((len(wild), 5)))
    th[7, 2] = 0.02           # a very narrow peak: |zs| far beyond the clamp for most points
    th[12, 2] = -1.5          # negative width (squared away, :exp of -z^2/2)
    th[17, 1] = 400.0         # peak far outside the data
    cfg["theta0"] = th
    return cfg, [0, 1, 2, 7, 12, 17, 22, 130, 257, 300, 384, 511]


def test_c2_full_size_matches_oracle(orc):
    cfg, chains = _c2_mixed()
    _check_logp(orc, cfg, "C2 N=100000 (5 tiles/item, 157 segments)", chains)


def test_c3_full_size_poisson_matches_oracle(orc):
    cfg = workloads.c3(W=300)
    _check_logp(orc, cfg, "C3 N=1e6 sinus + log-poisson", [0, 131, 299], builtin_formula=True)


def test_c4_full_size_fused_global_fit_matches_oracle(orc):
    cfg = workloads.c4(W=260)
    _check_logp(orc, cfg, "C4 D=8 x N=1e6 nv_multi_2, one fused launch", [0, 259], builtin_formula=True, compensated=False)


def test_c5_full_size_user_model_matches_oracle(orc):
    cfg = workloads.c5(W=130)
    _check_logp(orc, cfg, "C5 N=1e7 lowered poly5 x exp (77 tiles/item, 1015 segments)", [0, 129], compensated=False)


@pytest.mark.parametrize("N", [40000, 40001])
def test_segment_tile_sweep_matches_oracle(orc, N):
    """mfit_set_segment_tiles in {1, 2, 5, 8, 77}: every geometry within tolerance of the oracle; the automatic choice is
    bit-identical to the same value set explicitly (the geometry, not the knob, decides the bits)."""
    cfg = workloads.c2(W=200, N=N)
    ow = oracle_walker(orc, cfg)
    chains = [0, 64, 199]
    ref = np.array([ow.logp(cfg["theta0"][c])[0] for c in chains])
    res, geo = {}, {}
    for tps in (0, 1, 2, 5, 8, 77):
        w = workloads.make_walker(cfg)
        w._chk(_lib.lib().mfit_set_segment_tiles(w._h, tps))
        res[tps] = w.make_step(cfg["theta0"])[0]
        geo[tps] = w.segment_geometry()
        w.close()
        assert _strict(res[tps][chains], ref) < TOL, "tiles per segment %d" % tps
    auto = int(geo[0][0])
    w = workloads.make_walker(cfg)
    w._chk(_lib.lib().mfit_set_segment_tiles(w._h, auto))
    same = w.make_step(cfg["theta0"])[0]
    w.close()
    assert np.array_equal(same, res[0])
    _record("segment sweep N=%d" % N, auto_tiles_per_segment=auto,
            strict={str(k): _strict(v[chains], ref) for k, v in res.items()})


def test_c2_full_size_adaptive_run_identical_to_oracle(orc):
    """200 steps of walker-adaptive-steps-full at full C2 N under injected streams: parameter traces bit-identical,
    log p within 1e-12 (start-up phase: wild proposals, x0.1 rescale at i = 200 not yet reached -> pure accept logic)."""
    from test_gpu_parity import _compare_chain
    cfg = workloads.c2(W=300)
    w = workloads.make_walker(cfg)
    n = 260
    rng = np.random.default_rng(21)
    z, u = rng.standard_normal((w.W, n, w.P)), rng.uniform(size=(w.W, n))
    w.adaptive_steps_full(n=30000, temperature=10, auto=None, z=z, u=u, max_steps=n)
    chains = [0, 299]
    th, pr, nv = w.trace(n + 1)
    ow = oracle_walker(orc, cfg, use_builtin_formula=True)
    scale = ll_scale(cfg)
    worst = 0.0
    for c in chains:
        ow.init(cfg["theta0"][c])
        ow.adaptive_steps_full(n=30000, temperature=10, auto=orc.AUTO_NONE, z=z[c], u=u[c], max_steps=n)
        oth, opr = ow.trace(0)
        assert nv[c] == len(opr) == n + 1
        assert np.array_equal(th[c, :n + 1], oth), "chain %d: parameter trace differs from the oracle" % c
        worst = max(worst, _strict(pr[c, :n + 1], opr))
        assert rel_err(pr[c, :n + 1], opr, scale) < TOL
    _record("C2 N=100000 adaptive 260 steps, injected streams", strict_logp=worst, chains=len(chains))
    assert worst < TOL
    w.close()
