"""world_size-2 `gloo` tests (CPU) of the N>1 host logic in lisp_mcmc_b200/distributed.py: shard plans, the NCCL unique-id
broadcast, gathers with uneven shards, and -- with the ORACLE standing in for the GPU as the per-rank compute -- that a
chain-sharded run reassembles bit for bit into the single-rank run and that a dataset-sharded walker-make-step sums to
the same bits as the unsharded one."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                          MASTER_PORT=str(port))
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import torch.distributed as dist
        from lisp_mcmc_b200 import distributed as D, workloads
        from oracle import oracle as orc
        from helpers import oracle_walker
        r, lr, w = D.init_process_group("gloo")
        assert (r, w) == (rank, world)
        out = {}

        # 1. the NCCL unique id made on rank 0 reaches every rank unchanged
        uid = D.broadcast_nccl_unique_id()
        ids = [None] * world
        dist.all_gather_object(ids, uid.raw)
        out["uid_same"] = all(i == ids[0] for i in ids) and len(ids[0]) == 128 and any(b != 0 for b in ids[0])

        # 2. chain sharding: W = 7 chains over 2 ranks (4 + 3), oracle walkers as the per-rank compute
        cfg = workloads.c2(W=7, N=200)
        W, n = 7, 300
        rng = np.random.default_rng(5)
        z, u = rng.standard_normal((W, 2000, 5)), rng.uniform(size=(W, 2000))   # streams indexed by GLOBAL chain id
        lo, cnt = D.chain_shard(W)
        out["shard"] = (lo, cnt)

        def run_chain(c):
            ow = oracle_walker(orc, cfg)
            ow.init(cfg["theta0"][c])
            ow.adaptive_steps_full(n=n, temperature=10, auto=orc.AUTO_NONE, z=z[c], u=u[c])
            th, pr = ow.trace(0)
            return th, pr

        local = [run_chain(c) for c in range(lo, lo + cnt)]
        th = D.gather_chains(np.stack([t for t, _ in local]), W)
        pr = D.gather_chains(np.stack([p for _, p in local]), W)
        nv = D.gather_chains(np.array([len(p) for _, p in local], dtype=np.int64), W)
        if rank == 0:
            full = [run_chain(c) for c in range(W)]
            out["chains_bitwise"] = bool(np.array_equal(th, np.stack([t for t, _ in full])) and
                                         np.array_equal(pr, np.stack([p for _, p in full])) and
                                         np.array_equal(nv, [len(p) for _, p in full]))

        # 3. dataset sharding: a 3-data-set global fit, data set d evaluated on rank owner[d], terms gathered and
        #    added in data-set order
        g = workloads.c4(W=5, D=3, N=150)
        owners = D.dataset_owners(3)
        out["owners"] = owners
        ow = oracle_walker(orc, g)
        ll = np.zeros((3, 5))
        lp = np.zeros((3, 5))
        for c in range(5):
            _, ll_d, lp_d = ow.logp(g["theta0"][c])
            for d in range(3):
                if owners[d] == rank:
                    ll[d, c], lp[d, c] = ll_d[d], lp_d[d]
        import torch
        t = torch.from_numpy(np.stack([ll, lp]))
        dist.all_reduce(t)   # every entry is owned by exactly one rank, the others contribute an exact 0
        prob = D.sum_in_dataset_order(t[0].numpy(), t[1].numpy())
        ref = np.array([ow.logp(g["theta0"][c])[0] for c in range(5)])
        out["datasets_bitwise"] = bool(np.array_equal(prob, ref))

        # 3b. per-chain statistics (what Walker.stats() returns) gathered field by field in global chain order
        sdt = [("length", "i8"), ("age", "i8"), ("accept_200", "f8"), ("shutting_down", "i4")]
        mine = np.zeros(cnt, dtype=sdt)
        mine["length"] = np.arange(lo, lo + cnt) + 100
        mine["age"] = mine["length"] + 1
        mine["accept_200"] = 0.25 + 0.01 * np.arange(lo, lo + cnt)
        mine["shutting_down"] = np.arange(lo, lo + cnt) % 2
        allst = D.gather_stats(mine, W)
        out["stats_ok"] = bool(np.array_equal(allst["length"], np.arange(W) + 100) and allst.dtype == np.dtype(sdt) and
                               np.array_equal(allst["accept_200"], 0.25 + 0.01 * np.arange(W)) and
                               np.array_equal(allst["shutting_down"], np.arange(W) % 2))

        # 4. max over ranks
        out["max"] = D.max_over_ranks(10.0 + rank)
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, out))
    except Exception as e:   # surface the failure in the parent
        import traceback
        q.put((rank, {"error": "%s\n%s" % (e, traceback.format_exc())}))


def test_two_rank_gloo_host_logic():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in procs:
        rank, out = q.get(timeout=240)
        res[rank] = out
    for p in procs:
        p.join(timeout=60)
    for r in (0, 1):
        assert "error" not in res[r], res[r].get("error")
        assert res[r]["uid_same"]
        assert res[r]["owners"] == [0, 1, 0]
        assert res[r]["datasets_bitwise"]
        assert res[r]["stats_ok"]
        assert res[r]["max"] == 11.0
    assert res[0]["shard"] == (0, 4) and res[1]["shard"] == (4, 3)
    assert res[0]["chains_bitwise"]


def test_shard_plans_cover_everything():
    from lisp_mcmc_b200 import shard_chains, shard_datasets
    for W in (1, 7, 4096, 65537):
        for world in (1, 2, 4, 8):
            parts = [shard_chains(W, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == W
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert shard_datasets(8, 8) == list(range(8)) and shard_datasets(8, 2) == [0, 1] * 4
