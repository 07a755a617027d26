import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
from lisp_mcmc_b200 import workloads
cfg = workloads.c1(W=1)
for rep in range(2):
    w = workloads.make_walker(cfg, seed=11, trace_capacity=100010)
    w.adaptive_steps_full(n=3000, temperature=10, auto=None)
    t0 = time.perf_counter()
    w.adaptive_steps_full(n=100000, temperature=10, auto=None)
    dt = time.perf_counter() - t0
    st = w.stats()
    print("C1: %d steps in %.3f s -> %.0f MH steps/s (%.2f us/step), best prob %.6f" % (st["steps_this_run"][0], dt, st["steps_this_run"][0] / dt, dt / st["steps_this_run"][0] * 1e6, st["best_prob"][0]))
    w.close()
