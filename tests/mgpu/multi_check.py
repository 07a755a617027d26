"""Several GPUs behind ONE handle (mfit_multi_*, csrc/multi.cpp) -- the single-process twin of mgpu_check.py: no torchrun,
no torch.distributed, one Python process, the library runs one host thread per GPU.  What the SBCL host's :devices uses.
   python tests/mgpu/multi_check.py [n_gpus]
1. chains sharded over the GPUs: trace / stats / best / last step equal, bit for bit, those of all chains on GPU 0.
2. data sets of a global fit sharded over the GPUs (peer exchange between contexts of one process: raw pointers, no IPC):
   trace bit-identical to the one-GPU fit; walker-make-step over a batch equal too.
3. :auto :prob-settle on the sharded fit equals the one-GPU run; a stop request raised while the run is in flight ends
   it on every GPU at the same step (no hang, finished = 1).
Prints MULTI_OK."""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    from lisp_mcmc_b200 import _lib, workloads
    import torch
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    if n < 2:
        print("MULTI_SKIP: %d GPU" % n)
        return 0
    devs = list(range(n))
    ok = True

    # 1. chains
    W, steps = 37, 400
    cfg = workloads.c2(W=W, N=3000)
    wm = workloads.make_walker(cfg, seed=11, devices=devs, shard="chains")
    w0 = workloads.make_walker(cfg, seed=11, device=0)
    for w in (wm, w0):
        w.adaptive_steps_full(n=steps, temperature=10, auto=None)
    thm, prm, nvm = wm.trace(2001)
    th0, pr0, nv0 = w0.trace(2001)
    sm, s0 = wm.stats(), w0.stats()
    ok1 = np.array_equal(nvm, nv0) and np.array_equal(thm, th0, equal_nan=True) and np.array_equal(prm, pr0, equal_nan=True)
    ok1 = ok1 and all(np.array_equal(sm[k], s0[k]) for k in sm.dtype.names)
    ok1 = ok1 and all(np.array_equal(a, b) for a, b in zip(wm.most_likely(), w0.most_likely()))
    ok1 = ok1 and all(np.array_equal(a, b) for a, b in zip(wm.last_step(), w0.last_step()))
    ok1 = ok1 and np.array_equal(wm.l_matrix(), w0.l_matrix())
    am, a0 = wm.analytics(300), w0.analytics(300)          # walker-get's trace analytics, the fit envelope, list surgery
    ok1 = ok1 and all(np.array_equal(am[k], a0[k], equal_nan=True) for k in a0)
    ok1 = ok1 and all(np.array_equal(u, v, equal_nan=True) for u, v in zip(wm.data_and_fit(take=200), w0.data_and_fit(take=200)))
    wm.modify("burn-walks", burn_number=50); w0.modify("burn-walks", burn_number=50)
    ok1 = ok1 and all(np.array_equal(u, v, equal_nan=True) for u, v in zip(wm.trace(2001), w0.trace(2001)))
    theta = cfg["theta0"] * (1 + 0.01 * np.arange(W)[:, None])
    ok1 = ok1 and all(np.array_equal(a, b, equal_nan=True) for a, b in zip(wm.make_step(theta), w0.make_step(theta)))
    print("chains over %d GPUs behind one handle: %s (%d chains x %d rows)" % (n, "bit-identical" if ok1 else "MISMATCH", W, int(nv0.max())))
    wm.close(); w0.close()
    # fewer chains than GPUs: the empty shards idle
    cfg1 = workloads.c2(W=1, N=500)
    w1m = workloads.make_walker(cfg1, seed=3, devices=devs, shard="chains")
    w1 = workloads.make_walker(cfg1, seed=3, device=0)
    for w in (w1m, w1):
        w.adaptive_steps_full(n=300, temperature=10, auto=None)
    ok1b = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(w1m.trace(500), w1.trace(500)))
    print("one chain over %d GPUs: %s" % (n, ok1b))
    w1m.close(); w1.close()
    ok = ok and ok1 and ok1b

    # 2. data sets
    Dn = max(n, 3)
    g = workloads.c4(W=24, D=Dn, N=2000)
    gm = workloads.make_walker(g, seed=5, devices=devs, shard="datasets")
    g0 = workloads.make_walker(g, seed=5, device=0)
    th = g["theta0"] * (1 + 0.003 * np.arange(24)[:, None])
    ok2 = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(gm.make_step(th), g0.make_step(th)))
    for w in (gm, g0):
        w.adaptive_steps_full(n=300, temperature=10, auto=None)
    ok2 = ok2 and all(np.array_equal(a, b, equal_nan=True) for a, b in zip(gm.trace(2001), g0.trace(2001)))
    ag, a0 = gm.analytics(200), g0.analytics(200)
    ok2 = ok2 and all(np.array_equal(ag[k], a0[k], equal_nan=True) for k in a0)
    for d in (0, Dn - 1):      # the fit envelope of a data set is evaluated on the GPU that owns it
        ok2 = ok2 and all(np.array_equal(u, v, equal_nan=True)
                          for u, v in zip(gm.data_and_fit(take=150, fn_number=d), g0.data_and_fit(take=150, fn_number=d)))
    print("data sets (%d over %d GPUs) behind one handle: %s" % (Dn, n, "bit-identical" if ok2 else "MISMATCH"))
    gm.close(); g0.close()
    ok = ok and ok2

    # 3. auto stop; stop request in flight
    am = workloads.make_walker(g, seed=6, devices=devs, shard="datasets")
    a0 = workloads.make_walker(g, seed=6, device=0)
    for w in (am, a0):
        w.adaptive_steps(n=5000)
    ta, t0 = am.trace(7001), a0.trace(7001)
    ok3 = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(ta, t0))
    print("sharded :prob-settle run equals the one-GPU run: %s (steps per chain %d..%d)" % (ok3, t0[2].min(), t0[2].max()))
    am.close(); a0.close()
    L = _lib.lib()
    for mode in ("datasets", "chains"):
        ws = workloads.make_walker(g, seed=7, devices=devs, shard=mode)
        cfgc = _lib.AdaptCfg(3000000, 10.0, _lib.AUTO_NONE, _lib.OPT_COVARIANCE, 0, 0, 0)
        ws._chk(L.mfit_multi_run_begin(ws._h, C.byref(cfgc), None))
        done, fin = C.c_int64(), C.c_int()
        ws._chk(L.mfit_multi_run_advance(ws._h, 100, C.byref(done), C.byref(fin)))
        assert done.value == 100 and fin.value == 0
        timer = threading.Timer(0.3, lambda: L.mfit_multi_request_stop(ws._h))     # (setf mfit-walker-estop t) from another thread
        timer.start()
        t = time.time()
        ws._chk(L.mfit_multi_run_advance(ws._h, -1, C.byref(done), C.byref(fin)))
        timer.join()
        lens = ws.stats()["length"]
        ok4 = fin.value == 1 and 0 < done.value < 3000000 - 100 and time.time() - t < 60
        if mode == "datasets":     # every GPU holds all chains: the same step everywhere
            per_gpu = []
            for r in range(n):
                arr = (_lib.ChainStats * ws.W)()
                assert L.mfit_stats_read(L.mfit_multi_ctx(ws._h, r), arr) == 0
                per_gpu.append([a.steps_this_run for a in arr])
            ok4 = ok4 and all(p == per_gpu[0] for p in per_gpu)
        print("stop request in flight (%s): left after %d more steps, finished=%d, walk lengths %d..%d: %s"
              % (mode, done.value, fin.value, lens.min(), lens.max(), ok4))
        ws.close()
        ok3 = ok3 and ok4
    ok = ok and ok3
    print("MULTI_OK" if ok else "MULTI_FAIL")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
