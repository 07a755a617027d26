"""Multi-GPU determinism check, launched with torchrun (one rank per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/mgpu/mgpu_check.py
1. chain sharding: W chains split over N ranks give, after the gather, bit for bit the trace rank 0 gets from all W
   chains on its own GPU (Philox keyed by global chain id).
2. dataset sharding: a D-data-set global fit with data sets spread over the ranks (per-step NCCL all-gather inside
   libmfit_cuda.so, ordered sum) gives bit for bit the trace of the same fit on one GPU.
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
