"""Wall-clock phases of the end-to-end path on a B200 (not a bench line): python tools/runs/e2e_phases.py [steps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lisp_mcmc_b200 import workloads
K = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfg = workloads.c2()
for rep in range(4):
    if rep == 3:
        os.environ["MFIT_DEBUG_TIMING"] = "1"
    t = [time.perf_counter()]
    w = workloads.make_walker(cfg, seed=rep, trace_capacity=K + 1100); t.append(time.perf_counter())
    w.adaptive_steps_full(n=30000, temperature=10, auto=None, max_steps=K); t.append(time.perf_counter())
    out = w.trace(K + 1); t.append(time.perf_counter())
    w.close(); t.append(time.perf_counter())
    print("rep %d: create %.3f run %.3f trace %.3f close %.3f ms" % ((rep,) + tuple(1e3 * (b - a) for a, b in zip(t, t[1:]))), flush=True)
