"""A small tour of every kernel of the library (shared-data and per-walker likelihood kernels, built-in, NVRTC and
sequence-x models, the per-chain MH / adaptation / re-estimate kernels, trace analytics, envelope, restore, modify) at
sizes that finish in seconds: every C-ABI call synchronises and checks cudaGetLastError, so a faulting kernel fails the
tour.  Written for compute-sanitizer; the sanitizer is closed on this GPU pool, so in round 1 the tour only
ran plain (gpurun_out/san_plain.log: "sanitize tour done") -- no memcheck / racecheck result exists.

  python tools/sanitize_run.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def adaptive_slice(w, steps):
    """the first `steps` MH steps of walker-adaptive-steps-full :n 30000 :temperature 10, then mfit-walker-estop"""
    import ctypes as C
    from lisp_mcmc_b200 import _lib
    cfg = _lib.AdaptCfg(30000, 10.0, _lib.AUTO_NONE, _lib.OPT_COVARIANCE, 0, 0, 0)
    done, fin = C.c_int64(), C.c_int()
    w._chk(w._L.mfit_run_begin(w._h, C.byref(cfg), None, None, None, 0))
    w._chk(w._L.mfit_run_advance(w._h, steps, C.byref(done), C.byref(fin)))
    w._chk(w._L.mfit_request_stop(w._h))
    w._chk(w._L.mfit_run_advance(w._h, 1, C.byref(done), C.byref(fin)))
    assert fin.value == 1


def main():
    from lisp_mcmc_b200 import workloads
    from lisp_mcmc_b200.walker import walker_create
    rng = np.random.default_rng(0)
    # shared-data kernel (TMA ring, work-item queue), per-chain MH / adaptation / re-estimate kernels
    cfg = workloads.c2(W=300, N=1500)
    w = workloads.make_walker(cfg, seed=3)
    w.make_step(cfg["theta0"])
    adaptive_slice(w, 420)                                  # two acceptance windows: rescales and re-estimates
    w.many_steps(20, np.diag(0.002 * np.abs(cfg["theta_star"])))
    w.sample_region(1e-3)
    w.force_take_step()
    # trace consumers
    w.analytics(700)
    w.data_and_fit(take=500, n_fit=300)
    w.data_and_fit(take=77, which_solution="median", n_fit=1000)
    th, pr, nv = w.trace(int(w.stats()["length"].max()))
    w.restore(th, pr, nv, w.stats()["age"], *w.most_likely())
    w.modify("burn-walks", burn_number=50)
    w.modify("reset-to-most-likely")
    w.many_steps(10, np.diag(0.002 * np.abs(cfg["theta_star"])))
    w.modify("reset")
    w.close()
    # global fit (fused multi-data-set launch, lowered priors), Poisson, NVRTC user model, sequence-valued x
    for c in (workloads.c4(W=260, D=2, N=700), workloads.c3(W=130, N=900), workloads.c5(W=70, N=600)):
        w = workloads.make_walker(c, seed=4)
        adaptive_slice(w, 40)
        w.close()
    x = np.stack([np.linspace(0, 3, 500), rng.uniform(0, 2, 500)], axis=1)
    y = 1.5 * x[:, 0] + 4 * np.exp(-0.8 * x[:, 1]) + 0.05 * rng.standard_normal(500)
    w = walker_create(function=["(lambda (x &key a b k &allow-other-keys) (+ (* a (elt x 0)) (* b (exp (* -1 k (elt x 1))))))"],
                      data=[[x, y]], params=[("a", 1.5), ("b", 4.0), ("k", 0.8)], data_error=0.05, n_walkers=140, seed=5)
    w.many_steps(30, np.diag([0.01, 0.02, 0.005]))
    w.close()
    # walker set with per-walker data (one warp per chain; staged and unstaged shapes)
    for n in (400, 2500):
        ns = workloads.nvset(W=300, N=n)
        w = workloads.make_walker_set(ns, seed=6)
        adaptive_slice(w, 40)
        w.close()
    print("sanitize tour done")


if __name__ == "__main__":
    main()
