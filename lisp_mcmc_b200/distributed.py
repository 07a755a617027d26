"""One process per GPU: host-side plumbing for runs over several B200s of one box.

The reference has no notion of more than one worker (a walker set is advanced with `mapc`, nv-specific.lisp:58-66);
this is the sharding SURVEY.md 8(e) specifies:

  chains   W independent chains are split contiguously over the ranks, the data set is replicated, Philox streams are
           keyed by GLOBAL chain id (chain_id_base), so an N-rank run draws the same numbers as a 1-rank run.  Nothing
           crosses ranks during a run; traces and statistics are gathered afterwards.
  datasets global fits (walker-make-step's sum over data sets, mcmc-fitting.lisp:1068-1069): data set d lives on rank
           owner[d]; every rank holds all chains and replays identical proposals; per step the per-data-set ll / lp are
           all-gathered (NCCL, inside libmfit_cuda.so) and added in data-set order on every rank.

`torch.distributed` carries the rendezvous, the NCCL unique id and the result gathers: backend "nccl" on GPUs, "gloo"
on CPUs (tests/test_distributed_gloo.py runs this file's logic with world_size 2 on CPU)."""
import ctypes as C
import os

import numpy as np

from . import _lib
from .walker import shard_chains, shard_datasets


def env_ranks():
    """(rank, local_rank, world) from the variables torchrun / torch.distributed.run export"""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT.  Returns (rank, local_rank, world)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_ranks()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def _device():
    import torch
    dist = _dist()
    if dist is not None and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_nccl_unique_id():
    """The 128-byte ncclUniqueId every rank hands to mfit_nccl_attach: made on rank 0, broadcast to the others."""
    import torch
    dist = _dist()
    buf = C.create_string_buffer(128)
    if dist is None or dist.get_rank() == 0:
        rc = _lib.lib().mfit_nccl_unique_id(buf)
        if rc != 0:
            raise _lib.MfitError(rc, _lib.lib().mfit_last_error(None).decode())
    if dist is None:
        return buf
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone().to(_device())
    dist.broadcast(t, src=0)
    return C.create_string_buffer(bytes(t.cpu().numpy().tobytes()), 128)


_PROCESS_UID = None


def process_comm(rank=None, world=None):
    """(unique id, rank, world) for Walker(nccl=...): ONE ncclUniqueId per process and job -- made on rank 0 and broadcast
    the first time, reused afterwards -- so that every walker set of this process attaches to the same communicator
    (libmfit_cuda.so keeps one per id; ncclCommInitRank is a collective that costs tens of milliseconds)."""
    global _PROCESS_UID
    r, _, w = env_ranks()
    rank = r if rank is None else rank
    world = w if world is None else world
    if _PROCESS_UID is None:
        _PROCESS_UID = broadcast_nccl_unique_id()
    return _PROCESS_UID, rank, world


def gather_peer_handles(handle_bytes):
    """all ranks' 128-byte peer-exchange handles (mfit_peer_setup), concatenated in rank order -- the `gather` of
    Walker(peers=(rank, world, gather))"""
    import torch
    dist = _dist()
    if dist is None:
        return bytes(handle_bytes)
    mine = torch.frombuffer(bytearray(handle_bytes), dtype=torch.uint8).clone().to(_device())
    parts = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    return b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)


def peer_exchange(rank=None, world=None):
    """Walker(peers=...) argument for this job, or None when the peer exchange cannot be used (one rank, more than 8,
    MFIT_NO_PEERS set, or no GPUs): then the per-step exchange is the NCCL all-gather."""
    r, _, w = env_ranks()
    rank = r if rank is None else rank
    world = w if world is None else world
    dist = _dist()
    if world < 2 or world > 8 or os.environ.get("MFIT_NO_PEERS") or dist is None or dist.get_backend() != "nccl":
        return None
    return rank, world, gather_peer_handles


def chain_shard(W):
    """(first global chain id, count) of this rank"""
    dist = _dist()
    if dist is None:
        return 0, W
    return shard_chains(W, dist.get_rank(), dist.get_world_size())


def dataset_owners(D):
    """owner[d] for a global fit with D data sets over the ranks of this job"""
    dist = _dist()
    return shard_datasets(D, dist.get_world_size() if dist is not None else 1)


def gather_chains(local, W):
    """All-gather a per-chain array (leading axis = this rank's chains, in global chain order) into the full
    W-chain array on every rank.  Shards may be uneven (shard_chains); float64 / int64 arrays of any trailing shape."""
    import torch
    dist = _dist()
    local = np.ascontiguousarray(local)
    if dist is None:
        assert local.shape[0] == W
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, cnt = shard_chains(W, rank, world)
    if local.shape[0] != cnt:
        raise ValueError("rank %d holds %d chains, its shard of %d is %d" % (rank, local.shape[0], W, cnt))
    most = max(shard_chains(W, r, world)[1] for r in range(world))
    pad = np.zeros((most,) + local.shape[1:], dtype=local.dtype)
    pad[:cnt] = local
    dev = _device()
    mine = torch.from_numpy(pad).to(dev)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    out = np.empty((W,) + local.shape[1:], dtype=local.dtype)
    for r in range(world):
        rlo, rcnt = shard_chains(W, r, world)
        out[rlo:rlo + rcnt] = parts[r].cpu().numpy()[:rcnt]
    return out


def gather_trace(walker, take, W):
    """walker-get :steps of a chain-sharded walker set: (theta[W,take,P], prob[W,take], n_valid[W]) on every rank."""
    th, pr, nv = walker.trace(take)
    return gather_chains(th, W), gather_chains(pr, W), gather_chains(nv, W)


def gather_stats(stats, W):
    """The per-chain statistics of a chain-sharded walker set (Walker.stats(): length, age, acceptance over the last 200
    steps, best / last prob, temperature ...) on every rank, in global chain order: one gather per field."""
    stats = np.asarray(stats)
    out = np.empty(W, dtype=stats.dtype)
    for name in stats.dtype.names:
        col = np.ascontiguousarray(stats[name])
        wide = col.astype(np.int64) if col.dtype.kind in "iu" else col.astype(np.float64)
        out[name] = gather_chains(wide, W).astype(col.dtype)
    return out


def max_over_ranks(value):
    """device-timed durations are reported as the maximum over ranks"""
    import torch
    dist = _dist()
    if dist is None:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def sum_in_dataset_order(ll_by_dataset, lp_by_dataset):
    """walker-make-step :1067-1070 from per-data-set terms that may have been computed on different ranks:
    prob = (ll_0 + ll_1 + ...) + (lp_0 + lp_1 + ...), each sum left to right -- the order the device uses after its
    all-gather, which is what keeps accept decisions bit-identical on every rank (an all-reduce would not fix it)."""
    ll = np.asarray(ll_by_dataset, dtype=np.float64)
    lp = np.asarray(lp_by_dataset, dtype=np.float64)
    s_ll, s_lp = ll[0].copy(), lp[0].copy()
    for d in range(1, ll.shape[0]):
        s_ll = s_ll + ll[d]
        s_lp = s_lp + lp[d]
    return s_ll + s_lp


def walker_create_sharded(cfg, mode="chains", **gpu):
    """Build this rank's part of a workload dict (lisp_mcmc_b200.workloads): mode "chains" -> this rank's chains with
    chain_id_base set; mode "datasets" -> all chains, data sets distributed by dataset_owners, NCCL attached."""
    from . import workloads
    rank, local_rank, world = env_ranks()
    gpu.setdefault("device", local_rank)
    W = cfg["theta0"].shape[0]
    if mode == "chains":
        lo, cnt = chain_shard(W)
        return workloads.make_walker(cfg, theta0=np.ascontiguousarray(cfg["theta0"][lo:lo + cnt]), chain_id_base=lo, **gpu)
    if mode != "datasets":
        raise ValueError("mode must be 'chains' or 'datasets'")
    if world == 1:
        return workloads.make_walker(cfg, **gpu)
    return workloads.make_walker(cfg, nccl=process_comm(rank, world), peers=peer_exchange(rank, world),
                                 dataset_owners=dataset_owners(len(cfg["function"])), **gpu)
