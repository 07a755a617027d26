"""Timing of the one-handle multi-GPU path (mfit_multi_*, one process, one host thread per GPU inside the library) -- not a
bench line, a builder-run record for profiles/:  python tools/multi_bench.py [n_gpus]

  C4 (8 data sets x 1e6 points, 1024 chains): data sets sharded over the GPUs, per-step exchange of the per-data-set terms
     over NVLink peer memory between contexts of one process; against the same fit on one GPU.
  C2 (100 000 points x 4096 chains PER GPU): chains sharded, nothing crosses GPUs; against 4096 chains on one GPU.
Wall-clock around mfit_multi_run_advance of a run that is already warm (the handle has no device-side timer of its own:
the GPUs run concurrently under n host threads), best of 3 blocks."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed_steps(w, steps, multi):
    from lisp_mcmc_b200 import _lib
    L = _lib.lib()
    cfg = _lib.AdaptCfg(30000, 10.0, _lib.AUTO_NONE, _lib.OPT_COVARIANCE, 0, 0, 0)
    done, fin = C.c_int64(), C.c_int()
    if multi:
        w._chk(L.mfit_multi_run_begin(w._h, C.byref(cfg), None))
        adv = lambda k: w._chk(L.mfit_multi_run_advance(w._h, k, C.byref(done), C.byref(fin)))
    else:
        w._chk(L.mfit_run_begin(w._h, C.byref(cfg), None, None, None, 0))
        adv = lambda k: w._chk(L.mfit_run_advance(w._h, k, C.byref(done), C.byref(fin)))
    adv(steps)                      # warm-up block
    best = 1e30
    for _ in range(3):
        t = time.perf_counter()
        adv(steps)
        best = min(best, time.perf_counter() - t)
    return best / steps


def main():
    import torch
    from lisp_mcmc_b200 import workloads
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    devs = list(range(n))
    g = workloads.c4()
    pts = sum(len(d[0]) for d in g["data"])
    W = g["theta0"].shape[0]
    w1 = workloads.make_walker(g, seed=1, device=0)
    t1 = timed_steps(w1, 40, False)
    w1.close()
    wm = workloads.make_walker(g, seed=1, devices=devs, shard="datasets")
    tm = timed_steps(wm, 40, True)
    wm.close()
    print("C4 data sets over %d GPUs behind one handle: %.3f ms/step (%.3g point-evals/s) against %.3f ms/step on one GPU: %.2f x"
          % (n, tm * 1e3, W * pts / tm, t1 * 1e3, t1 / tm))
    c = workloads.c2(W=4096 * n)
    N = len(c["data"][0][0])
    c1 = workloads.c2(W=4096)
    w1 = workloads.make_walker(c1, seed=1, device=0)
    t1 = timed_steps(w1, 200, False)
    w1.close()
    wm = workloads.make_walker(c, seed=1, devices=devs, shard="chains")
    tm = timed_steps(wm, 200, True)
    wm.close()
    print("C2 %d chains over %d GPUs behind one handle: %.4f ms/step (%.4g point-evals/s) against %.4f ms/step for 4096 chains on one GPU: weak-scaling efficiency %.3f"
          % (4096 * n, n, tm * 1e3, 4096 * n * N / tm, t1 * 1e3, t1 / tm))


if __name__ == "__main__":
    main()
