"""Multi-GPU determinism check, launched with torchrun (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu/mgpu_check.py
1. chain sharding: W chains split over N ranks give, after the gather, bit for bit the trace rank 0 gets from all W
   chains on its own GPU (Philox keyed by global chain id).
2. dataset sharding: a D-data-set global fit with data sets spread over the ranks gives bit for bit the trace of the same
   fit on one GPU -- with the peer exchange (every rank's k_finalize stores its terms into the peers' buffers over NVLink,
   k_post waits on flags) and with the NCCL all-gather (MFIT_NO_PEERS), ordered sum either way.
3. the same fit with :auto :prob-settle (chains stop on their own, the host loop must end at the same step on every rank)
   and with a stop request raised on ONE rank only: no rank hangs, all leave the loop after the same number of steps.
Prints MGPU_OK on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from lisp_mcmc_b200 import distributed as D, workloads
    rank, local_rank, world = D.init_process_group("nccl")
    torch.cuda.set_device(local_rank)

    # 1. chains
    W, n = 37, 400
    cfg = workloads.c2(W=W, N=3000)
    w = D.walker_create_sharded(cfg, "chains", seed=11)
    w.adaptive_steps_full(n=n, temperature=10, auto=None)
    th, pr, nv = D.gather_trace(w, 2001, W)
    w.close()
    ok1 = True
    if rank == 0:
        w0 = workloads.make_walker(cfg, seed=11, device=local_rank)
        w0.adaptive_steps_full(n=n, temperature=10, auto=None)
        th0, pr0, nv0 = w0.trace(2001)
        w0.close()
        m = int(nv0.max())
        ok1 = np.array_equal(nv, nv0) and np.array_equal(th[:, :m], th0[:, :m]) and np.array_equal(pr[:, :m], pr0[:, :m])
        print("chain sharding over %d ranks: %s (%d chains x %d rows)" % (world, "bit-identical" if ok1 else "MISMATCH", W, m))

    # 2. datasets
    Dn = max(world, 3)
    g = workloads.c4(W=24, D=Dn, N=2000)
    wg = D.walker_create_sharded(g, "datasets", seed=5)
    wg.adaptive_steps_full(n=300, temperature=10, auto=None)
    thg, prg, nvg = wg.trace(2001)
    wg.close()
    ok2 = True
    if rank == 0:
        w1 = workloads.make_walker(g, seed=5, device=local_rank)
        w1.adaptive_steps_full(n=300, temperature=10, auto=None)
        th1, pr1, nv1 = w1.trace(2001)
        w1.close()
        m = int(nv1.max())
        ok2 = np.array_equal(nvg, nv1) and np.array_equal(thg[:, :m], th1[:, :m]) and np.array_equal(prg[:, :m], pr1[:, :m])
        print("dataset sharding (%d data sets over %d ranks): %s" % (Dn, world, "bit-identical" if ok2 else "MISMATCH"))
    # 2b. the same through the NCCL all-gather instead of the peer exchange
    os.environ["MFIT_NO_PEERS"] = "1"
    wn = D.walker_create_sharded(g, "datasets", seed=5)
    wn.adaptive_steps_full(n=300, temperature=10, auto=None)
    thn, prn, nvn = wn.trace(2001)
    wn.close()
    del os.environ["MFIT_NO_PEERS"]
    ok2b = np.array_equal(thn, thg) and np.array_equal(prn, prg, equal_nan=True)
    if rank == 0:
        print("NCCL all-gather and peer exchange agree: %s" % ok2b)
    ok2 = ok2 and ok2b

    # 3. auto stop and a one-sided stop request on the sharded fit
    import ctypes as C
    from lisp_mcmc_b200 import _lib
    wa = D.walker_create_sharded(g, "datasets", seed=6)
    wa.adaptive_steps(n=5000)                       # :auto :prob-settle, temperature 10
    tha, pra, nva = wa.trace(7001)
    wa.close()
    ok4 = True
    if rank == 0:
        w2 = workloads.make_walker(g, seed=6, device=local_rank)
        w2.adaptive_steps(n=5000)
        th2, pr2, nv2 = w2.trace(7001)
        w2.close()
        ok4 = np.array_equal(nva, nv2) and np.array_equal(tha, th2, equal_nan=True)
        print("sharded :prob-settle run equals the one-GPU run: %s (steps per chain %d..%d)" % (ok4, nva.min(), nva.max()))
    ws = D.walker_create_sharded(g, "datasets", seed=7)
    L = _lib.lib()
    cfgc = _lib.AdaptCfg(30000, 10.0, _lib.AUTO_NONE, _lib.OPT_COVARIANCE, 0, 0, 0)
    ws._chk(L.mfit_run_begin(ws._h, C.byref(cfgc), None, None, None, 0))
    done, fin = C.c_int64(), C.c_int()
    ws._chk(L.mfit_run_advance(ws._h, 100, C.byref(done), C.byref(fin)))
    if rank == world - 1:
        ws.request_stop()                           # (setf mfit-walker-estop t) on one rank only
    ws._chk(L.mfit_run_advance(ws._h, 5000, C.byref(done), C.byref(fin)))
    steps = torch.tensor([done.value, fin.value], dtype=torch.int64).cuda()
    allsteps = [torch.empty_like(steps) for _ in range(world)]
    dist.all_gather(allsteps, steps)
    ok5 = all(torch.equal(allsteps[0], p) for p in allsteps) and int(allsteps[0][0]) < 5000 and int(allsteps[0][1]) == 1
    ws.close()
    if rank == 0:
        print("one-sided stop request: every rank left after %d more steps, finished=%d: %s" % (int(allsteps[0][0]), int(allsteps[0][1]), ok5))
    ok2 = ok2 and ok4 and ok5

    # every rank must hold the same chains after a dataset-sharded run
    t = torch.from_numpy(prg[:, 0].copy()).cuda()
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    ok3 = all(torch.equal(parts[0], p) for p in parts)
    if rank == 0:
        print("all ranks agree on the chains of the global fit: %s" % ok3)
        print("MGPU_OK" if (ok1 and ok2 and ok3) else "MGPU_FAIL")
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not (ok1 and ok2 and ok3):
        sys.exit(1)


if __name__ == "__main__":
    main()
