"""Kernel-level timing on a B200 (not a bench line): likelihood-kernel device time per launch for a workload,
optionally for an alternative build of the library, and a phase breakdown of the end-to-end path.

  python tools/kbench.py --config C2 [--reps 20] [--lib path/to/libmfit_cuda.so] [--W 4096] [--e2e 200]
"""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--lib", default=None)
    ap.add_argument("--W", type=int, default=0)
    ap.add_argument("--N", type=int, default=0)
    ap.add_argument("--e2e", type=int, default=0)
    ap.add_argument("--seg-tiles", type=int, default=0)
    args = ap.parse_args()
    from lisp_mcmc_b200 import _lib, workloads
    if args.lib:
        _lib.LIB_PATH = os.path.abspath(args.lib)
    L = _lib.lib()
    kw = {}
    if args.W:
        kw["W"] = args.W
    if args.N:
        kw["N"] = args.N
    if args.config == "NVSET":
        cfg = workloads.nvset(**kw)
        N = cfg["data"][1].shape[1]
        W = cfg["theta0"].shape[0]
        w = workloads.make_walker_set(cfg)
    else:
        cfg = workloads.CONFIGS[args.config](**kw)
        N = sum(len(d[0]) for d in cfg["data"])
        W = cfg["theta0"].shape[0]
        w = workloads.make_walker(cfg)
    if args.seg_tiles:
        w._chk(L.mfit_set_segment_tiles(w._h, args.seg_tiles))
    w._chk(L.mfit_set_timing(w._h, 1))
    ksec, kl = C.c_double(), C.c_int64()
    best, tot = 1e30, 0.0
    for r in range(args.reps + 2):
        w.make_step(cfg["theta0"])
        L.mfit_eval_kernel_time(w._h, C.byref(ksec), C.byref(kl))
        if r >= 2:
            best = min(best, ksec.value)
            tot += ksec.value
    mean = tot / args.reps
    print("%s lib=%s seg_tiles=%d W=%d N=%d: eval kernels %.4f ms mean, %.4f ms best -> %.4g point-evals/s (mean), %.4g (best)"
          % (args.config, os.path.basename(_lib.LIB_PATH), args.seg_tiles, W, N, mean * 1e3, best * 1e3, W * N / mean, W * N / best))
    if args.config == "NVSET":
        by = cfg["bytes_per_point"] * W * N
        print("  per-walker data: %.1f MB of y/sigma per pass -> %.0f GB/s of algorithmic traffic (measured copy peak 6540 GB/s)"
              % (by / 1e6, by / mean / 1e9))
        cfgc = _lib.AdaptCfg(30000, 10.0, _lib.AUTO_NONE, _lib.OPT_COVARIANCE, 0, 0, 0)
        done, fin = C.c_int64(), C.c_int()
        w._chk(L.mfit_set_timing(w._h, 0))
        w._chk(L.mfit_run_begin(w._h, C.byref(cfgc), None, None, None, 0))
        w._chk(L.mfit_run_advance(w._h, 5, C.byref(done), C.byref(fin)))
        t0 = time.perf_counter()
        w._chk(L.mfit_run_advance(w._h, 50, C.byref(done), C.byref(fin)))
        dt = time.perf_counter() - t0
        print("  50 adaptive MH steps of %d walkers: %.2f ms/step -> %.3g MH steps/s, %.3g point-evals/s"
              % (W, dt / 50 * 1e3, W * 50 / dt, W * N * 50 / dt))
        for groups in (1, 2):
            w._chk(L.mfit_set_groups(w._h, groups))
            w._chk(L.mfit_set_timing(w._h, 1 if groups == 1 else 0))
            t0 = time.perf_counter()
            w._chk(L.mfit_run_advance(w._h, 50, C.byref(done), C.byref(fin)))
            dt = time.perf_counter() - t0
            L.mfit_eval_kernel_time(w._h, C.byref(ksec), C.byref(kl))
            print("  groups=%d: %.3f ms/step%s" % (groups, dt / 50 * 1e3,
                  ", eval kernel %.4f ms per launch inside the run" % (ksec.value / max(kl.value, 1) * 1e3) if groups == 1 else ""))
    w.close()
    if args.e2e:
        K = args.e2e
        cfgc = _lib.AdaptCfg(30000, 10.0, _lib.AUTO_NONE, _lib.OPT_COVARIANCE, 0, 0, 0)
        for rep in range(2):
            if rep == 1:   # where does walker_create spend its time?
                import cProfile, pstats
                pr_ = cProfile.Profile()
                pr_.enable()
                wp = workloads.make_walker(cfg, seed=1, trace_capacity=K + 1100)
                pr_.disable()
                wp.close()
                pstats.Stats(pr_).sort_stats("cumulative").print_stats(12)
            t = [time.perf_counter()]
            w2 = workloads.make_walker(cfg, seed=2025 + rep, trace_capacity=K + 1100)
            t.append(time.perf_counter())
            w2._chk(L.mfit_run_begin(w2._h, C.byref(cfgc), None, None, None, 0))
            t.append(time.perf_counter())
            done, fin = C.c_int64(), C.c_int()
            for lo in range(0, K, 50):
                w2._chk(L.mfit_run_advance(w2._h, min(50, K - lo), C.byref(done), C.byref(fin)))
                t.append(time.perf_counter())
            th, pr, nv = w2.trace(K + 1)
            t.append(time.perf_counter())
            w2.close()
            t.append(time.perf_counter())
            print("e2e K=%d rep %d: create %.1f ms, begin %.1f ms, advance(50s) %s ms, trace %.1f ms, close %.1f ms"
                  % (K, rep, (t[1] - t[0]) * 1e3, (t[2] - t[1]) * 1e3,
                     " ".join("%.1f" % ((b - a) * 1e3) for a, b in zip(t[2:-3], t[3:-2])), (t[-2] - t[-3]) * 1e3,
                     (t[-1] - t[-2]) * 1e3))


if __name__ == "__main__":
    main()
