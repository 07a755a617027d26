"""Timing of the off-step-path device work on a B200 (not a bench line): walker-get analytics, the fit curve + envelope of
walker-get-data-and-fit, trace read-back and checkpoint restore, for a C2-shaped walker set.  Wall clock around the
C-ABI calls (each one synchronises), host buffers included.

  python tools/abench.py [--W 4096] [--N 100000] [--steps 2000] [--take 1000]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--W", type=int, default=4096)
    ap.add_argument("--N", type=int, default=100000)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--take", type=int, default=1000)
    args = ap.parse_args()
    from lisp_mcmc_b200 import workloads
    cfg = workloads.c2(W=args.W, N=args.N)
    w = workloads.make_walker(cfg, seed=1)
    w.adaptive_steps_full(n=args.steps, temperature=10, auto=None)
    rows = int(w.stats()["length"].max())

    def timed(label, fn, reps=3):
        fn()
        t = []
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            t.append(time.perf_counter() - t0)
        print("%-58s %9.2f ms" % (label, 1e3 * min(t)), flush=True)
        return out

    print("C2-shaped set: W=%d chains, N=%d points, %d rows held per chain, take=%d" % (args.W, args.N, rows, args.take))
    timed("walker-get :most-likely-step/:acceptance/:covariance/:l-matrix", lambda: w.analytics(args.take, ["most-likely-step", "acceptance", "covariance-matrix", "l-matrix"]))
    timed("walker-get :median-params", lambda: w.analytics(args.take, ["median-params"]))
    out = timed("walker-get-data-and-fit (1000 abscissae, most likely)", lambda: w.data_and_fit(take=args.take))
    nd = np.empty(w.W, dtype=np.int64)
    import ctypes as C
    from lisp_mcmc_b200.walker import linspace
    xf = linspace(-10.0, 10.0, 1000)
    w._chk(w._L.mfit_fit_envelope(w._h, 0, args.take, 0, xf.ctypes.data_as(C.POINTER(C.c_double)), 1000, None, None, None, None,
                                  nd.ctypes.data_as(C.POINTER(C.c_int64))))
    m = int(np.ceil(np.float32(0.66) * np.float32(min(args.take, rows))))
    evals = float(nd.sum() + w.W) * 1000
    print("  selected entries per chain %d, distinct steps mean %.1f -> %.3g model evaluations" % (m, nd.mean(), evals))
    tr = timed("walker-get :steps (trace read-back, %d rows)" % rows, lambda: w.trace(rows), reps=2)
    st = w.stats()
    bth, bpr = w.most_likely()
    timed("walker-load restore (mfit_chains_restore, %d rows)" % rows, lambda: w.restore(tr[0], tr[1], tr[2], st["age"], bth, bpr), reps=2)
    w.close()


if __name__ == "__main__":
    main()
