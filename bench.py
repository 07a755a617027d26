#!/usr/bin/env python
"""bench.py -- likelihood point-evals/s and MH steps/s of the walker-adaptive-steps hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference] [--no-extras] [--no-cpu]

A "step" is one Metropolis-Hastings step of every chain: one pass of the likelihood kernel over the whole
dataset for all W proposals, then accept/reject/adapt on the device (walker-take-step, mcmc-fitting.lisp:1072-1095,
inside the loop of walker-adaptive-steps-full :902-942).  The N=1 workload is BASELINE.json configs[1]
(C2: Gaussian + linear background, 100k points, 4096 walkers, log-normal); with N>1 every rank holds the dataset
and W more chains (weak scaling, Philox streams keyed by global chain id, no collective on the data path).

value   = W_total * K * N_points / (device time of K steps), inputs resident in HBM.  The K-step block -- bracketed by a
          barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks -- is repeated on the
          same run until at least 0.5 s have been timed; `value` and `ms_per_step` are the MEDIAN block, `blocks` lists
          how many there were and the fastest / slowest.  (One 20-step block is 8 ms: a single sample of that says little.)
e2e     = the same quantity through the public API with host buffers: walker_create (H2D of x, y, sigma, theta0),
          K adaptive steps, read-back of the K-step trace (D2H), all inside the timed region; `breakdown_ms` splits it.
roofline= the likelihood kernel against the FP64 pipe: achieved = FP64 instructions the kernel executes per launch
          (counted in its SASS per loop variant, weighted by the tiles each variant ran -- mfit_eval_mode_tiles; 2 flop
          each) / its CUDA-event duration; peak = the DFMA issue rate measured live on this GPU (MEASURED_PEAKS.json
          has no FP64 figure).  SURVEY 8(d)'s algorithmic-flop accounting is reported beside it -- see DESIGN.md.
cpu_baseline / --impl reference: the CPU restatement of the reference algorithm (oracle/, C++ -O2, compiled
          model formula -- a conservative stand-in for SBCL, which is not installable here) on the host cores;
          cpu_baseline.cpu_1t is the same on one thread (the reference is single-threaded).
extra   = the other named configurations on this GPU (C1 single walker, C3, C4, C5 per-GPU share, the NV walker set),
          one short device-timed run each: ms_per_step, point-evals/s, roofline fraction.  N > 1: extra.c4_datasets is
          the global fit with its data sets spread over the ranks (per-step exchange of the per-data-set terms).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "likelihood_point_evals_per_sec"
UNIT = "point-evals/s"
MIN_TIMED_S = 0.5        # the K-step block is repeated until this much device time has been measured
MAX_BLOCKS = 400


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def ncu_traffic(kernel_regex="mfit_eval_kernel"):
    """dram bytes per launch of the likelihood kernel from the committed ncu capture (profiles/r02_eval_kernel_ncu_raw.csv:
    `ncu --set full`, page raw); None when the file is not there -- never a pasted literal"""
    import csv
    path = os.path.join(ROOT, "profiles", "r02_eval_kernel_ncu_raw.csv")
    try:
        rows = list(csv.reader(open(path)))
        h = rows[0]
        for r in rows[2:]:
            if kernel_regex in r[h.index("Kernel Name")]:
                def val(name):
                    v, u = float(r[h.index(name)]), rows[1][h.index(name)]
                    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                return int(val("dram__bytes_read.sum") + val("dram__bytes_write.sum")), os.path.relpath(path, ROOT)
    except Exception:
        pass
    return None, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "reasons": sorted(reasons), "samples": len(sm)}


def flops_per_point(cfg, costs):
    """add/sub/mul/div/sqrt = 1, FMA = 2, each transcendental at its measured cost in DFMA issue slots
    (x 2 flop).  `div` in transc_per_point marks divides whose real cost is the measured one instead of 1."""
    f = float(cfg["flop_per_point"])
    for k, n in cfg["transc_per_point"].items():
        c = costs.get(k)
        if c is None:
            continue
        f += n * 2.0 * c - (n if k == "div" else 0)
    return f


def cpu_reference_run(cfg, steps, warmup, threads, budget_s=20.0):
    """The oracle's restatement on a bounded sample: W_s chains x `steps` adaptive steps on the full dataset, chains split
    over `threads` threads.  W_s is sized from a probe so that the timed part stays near budget_s."""
    from oracle import oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_walker
    ow = oracle_walker(orc, cfg, use_builtin_formula=all(not isinstance(f, str) for f in cfg["function"]))
    N = sum(len(d[0]) for d in cfg["data"])
    th0 = cfg["theta0"]
    probe_w = min(len(th0), max(threads, 1))
    secs, _ = ow.bench_many(th0[:probe_w], 30000, threads, max_steps=max(1, warmup))   # doubles as warm-up
    per_chain_step = secs * min(threads, probe_w) / (probe_w * max(1, warmup))
    w_s = int(budget_s / max(per_chain_step * steps / threads, 1e-9))
    w_s = max(min(threads, len(th0)), min(len(th0), (w_s // threads) * threads))
    secs, _ = ow.bench_many(th0[:w_s], 30000, threads, max_steps=steps)
    value = w_s * steps * N / secs
    return dict(value=value, seconds=secs, chains=w_s, steps=steps, threads=threads,
                sample="%d chains x %d adaptive steps x %d points (full dataset), chains split over %d threads"
                       % (w_s, steps, N, threads))


# FP64-pipe instructions per point-eval of each loop variant of the likelihood kernel (SASS count, tools/sass_mix.py;
# profiles/r02_sass_*.txt): index = MFIT_MODE_* (guard-free, clamp-only, recording guard, fused argument)
def instr_per_point(cfg, mode_tiles):
    by_mode = cfg.get("fp64_instr_by_mode") or [cfg["fp64_instr_per_point"]] * 4
    tot = float(sum(mode_tiles))
    if tot <= 0:
        return float(cfg["fp64_instr_per_point"]), None
    return sum(i * t for i, t in zip(by_mode, mode_tiles)) / tot, [t / tot for t in mode_tiles]


class Device:
    """torch is the plumbing: device selection, a non-default stream shared with the library, events, NCCL rendezvous"""

    def __init__(self, local_rank, world):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.world = torch, dist, world
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU baseline)")
        torch.cuda.set_device(local_rank)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        self.tstream = torch.cuda.Stream()
        torch.cuda.set_stream(self.tstream)
        self.stream = C.c_void_p(self.tstream.cuda_stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def events(self):
        return self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device="cuda")
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]


def device_run(dev, L, _lib, w, K, Wm, repeat_s, kernel_steps=60, l_matrix=None):
    """warm-up, then K-step blocks of walker-adaptive-steps-full timed on the device, then the likelihood kernel on its own.
    Returns dict(block_ms [max over ranks per block], launches_per_block, kernel_s, kernel_launches, mode_tiles, stats)."""
    h = w._h
    cfgc = _lib.AdaptCfg(30000, 10.0, _lib.AUTO_NONE, _lib.OPT_COVARIANCE, 0, 0, 0)
    lm = None if l_matrix is None else np.ascontiguousarray(l_matrix, dtype=np.float64)
    w._chk(L.mfit_run_begin(h, C.byref(cfgc), None if lm is None else lm.ctypes.data_as(C.POINTER(C.c_double)), None, None, 0))
    done, fin = C.c_int64(), C.c_int()
    w._chk(L.mfit_run_advance(h, Wm, C.byref(done), C.byref(fin)))           # warm-up steps
    w._chk(L.mfit_flush_l2(h))
    blocks, total, launches = [], 0.0, 0
    while True:
        dev.barrier()
        l0 = L.mfit_launch_count(h)
        e0, e1 = dev.events()
        e0.record()
        w._chk(L.mfit_run_advance(h, K, C.byref(done), C.byref(fin)))        # one timed block of exactly K steps
        e1.record()
        dev.barrier()
        assert done.value == K, "ran %d steps instead of %d" % (done.value, K)
        ms = dev.max_over_ranks([e0.elapsed_time(e1)])[0]
        launches = L.mfit_launch_count(h) - l0
        blocks.append(ms)
        total += ms * 1e-3
        if total >= repeat_s or len(blocks) >= MAX_BLOCKS or (len(blocks) + 2) * K + Wm + kernel_steps >= 29000:
            break
    # the likelihood kernel on its own: a run advances the chains as groups on separate streams (one group's per-chain
    # kernels overlap the others' likelihood kernels), where an event interval around a kernel would include its wait for
    # SM slots; with timing on the library runs one group, records events around every eval launch and counts loop variants
    w._chk(L.mfit_set_timing(h, 1))
    tiles = (C.c_int64 * 4)()
    L.mfit_eval_mode_tiles(h, tiles)   # reset
    w._chk(L.mfit_run_advance(h, kernel_steps, C.byref(done), C.byref(fin)))
    ksec, kl = C.c_double(), C.c_int64()
    L.mfit_eval_kernel_time(h, C.byref(ksec), C.byref(kl))
    L.mfit_eval_mode_tiles(h, tiles)
    w._chk(L.mfit_set_timing(h, 0))
    st = w.stats()
    return dict(block_ms=blocks, launches=int(launches), kernel_s=ksec.value, kernel_launches=int(kl.value),
                kernel_steps=kernel_steps, mode_tiles=[int(t) for t in tiles], accept=float(np.mean(st["accept_200"])),
                steps_total=int(st["steps_this_run"].max()))


def roofline_of(cfg, run, W, N_owned, dfma, costs):
    per_launch_s = run["kernel_s"] / max(run["kernel_launches"], 1)
    evals = W * N_owned * run["kernel_steps"] / max(run["kernel_launches"], 1)   # point-evals per likelihood-kernel launch
    instr, mix = instr_per_point(cfg, run["mode_tiles"])
    achieved = evals * instr * 2.0 / per_launch_s / 1e12
    fpp = flops_per_point(cfg, costs)
    return dict(bound="fp64", achieved=achieved, peak=dfma / 1e12, unit="TFLOP/s", frac=achieved / (dfma / 1e12),
                kernel="mfit_eval_kernel", kernel_ms=per_launch_s * 1e3, fp64_instr_per_point_eval=instr,
                loop_variant_mix=dict(zip(("guard_free", "clamp_only", "guarded", "fused"), mix)) if mix else None,
                survey_8d={"flop_per_point_eval": fpp, "achieved": evals * fpp / per_launch_s / 1e12,
                           "frac": evals * fpp / per_launch_s / 1e12 / (dfma / 1e12)})


def extra_record(dev, L, _lib, workloads, name, dfma, costs, local_rank, steps, warmup=3, **kw):
    """one short device-timed run of another named configuration on this GPU"""
    walker_set = name == "NVSET"
    cfg = workloads.nvset(**kw) if walker_set else workloads.CONFIGS[name](**kw)
    N = cfg["data"][1].shape[1] if walker_set else sum(len(d[0]) for d in cfg["data"])
    W = cfg["theta0"].shape[0]
    make = workloads.make_walker_set if walker_set else workloads.make_walker
    t0 = time.perf_counter()
    w = make(cfg, seed=7, device=local_rank, trace_capacity=warmup + 2 * steps + 1200, stream=dev.stream)
    create_s = time.perf_counter() - t0
    run = device_run(dev, L, _lib, w, steps, warmup, 0.0, kernel_steps=min(steps, 20))
    run_ad = None
    if not walker_set:
        # The run above starts the way the reference does -- first l-matrix = diag(theta), :899-900 -- and is too short to
        # leave that start-up phase (wild proposals: the guarded / clamp-only loop variants).  A second run on the same
        # chains with :l-matrix diag(1e-3 |theta*|) shows the loop adapted chains run (what a long run spends its time in).
        try:
            run_ad = device_run(dev, L, _lib, w, steps, 2, 0.0, kernel_steps=min(steps, 20),
                                l_matrix=np.diag(1e-3 * np.abs(cfg["theta_star"])))
        except Exception:
            run_ad = None
    w.close()
    ms = float(np.median(run["block_ms"]))
    rec = {"workload": "%s: %s" % (name, (workloads.nvset if walker_set else workloads.CONFIGS[name]).__doc__.strip().split("\n")[0]),
           "chains": W, "points": N, "steps": steps, "ms_per_step": ms / steps, "value": W * steps * N / (ms * 1e-3),
           "unit": UNIT, "mh_steps_per_sec": W * steps / (ms * 1e-3), "walker_create_s": create_s,
           "gpu_launches": run["launches"], "mean_accept_200": run["accept"]}
    r = roofline_of(cfg, run, W, N, dfma, costs)
    if walker_set:   # bound by HBM, not by the FP64 pipe
        hbm_peak, hbm_src = load_peaks()
        per_launch_s = run["kernel_s"] / max(run["kernel_launches"], 1)
        by = float(cfg["bytes_per_point"]) * W * N
        rec["roofline"] = {"bound": "hbm", "achieved": by / per_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                           "frac": by / per_launch_s / 1e9 / hbm_peak, "kernel": "mfit_eval_perchain_reg_kernel",
                           "kernel_ms": per_launch_s * 1e3, "peak_source": "MEASURED_PEAKS.json hbm_gbs (%s)" % hbm_src,
                           "fp64_frac": r["frac"]}
    else:
        rec["roofline"] = {k: r[k] for k in ("bound", "achieved", "peak", "unit", "frac", "kernel_ms", "fp64_instr_per_point_eval")}
        rec["roofline"]["loop_variant_mix"] = r["loop_variant_mix"]
        if run_ad is not None:
            ra = roofline_of(cfg, run_ad, W, N, dfma, costs)
            ms_a = float(np.median(run_ad["block_ms"]))
            rec["adapted"] = {"what": "the same chains, run started with :l-matrix diag(1e-3 |theta*|): proposals of an adapted run",
                              "ms_per_step": ms_a / steps, "value": W * steps * N / (ms_a * 1e-3), "unit": UNIT,
                              "roofline_frac": ra["frac"], "kernel_ms": ra["kernel_ms"], "loop_variant_mix": ra["loop_variant_mix"],
                              "mean_accept_200": run_ad["accept"]}
    return rec


def c1_record(dev, L, _lib, workloads, local_rank, no_cpu):
    """C1 = the reference's own case (test.lisp:14-24): ONE walker, Lorentzian, 1000 points, walker-adaptive-steps n=100000.
    A single chain is launch-latency bound on a GPU; reported next to one CPU thread of the restatement and the only SBCL
    datum there is (6.34 s for 1e5 steps on 334 points = 15.8 k steps/s)."""
    cfg = workloads.c1(W=1)
    n = 100000
    rec = {"workload": "C1: one walker, Lorentzian, 1000 points, walker-adaptive-steps-full n=100000 auto=nil",
           "sbcl_published_steps_per_sec": 15.8e3, "sbcl_source": "test.lisp:23 (334 points, unknown machine)"}
    w = workloads.make_walker(cfg, seed=11, device=local_rank, trace_capacity=n + 10, stream=dev.stream)
    w.adaptive_steps_full(n=3000, temperature=10, auto=None)       # warm-up: first-use costs, graph capture
    dev.barrier()
    e0, e1 = dev.events()
    t0 = time.perf_counter()
    e0.record()
    w.adaptive_steps_full(n=n, temperature=10, auto=None)
    e1.record()
    dev.barrier()
    wall = time.perf_counter() - t0
    steps = int(w.stats()["steps_this_run"][0])
    w.close()
    rec.update({"gpu_steps": steps, "gpu_seconds": max(wall, e0.elapsed_time(e1) * 1e-3), "gpu_mh_steps_per_sec": steps / max(wall, 1e-9),
                "gpu_point_evals_per_sec": steps * 1000 / max(wall, 1e-9)})
    if not no_cpu:
        from oracle import oracle as orc
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import oracle_walker
        ow = oracle_walker(orc, cfg, use_builtin_formula=True)
        secs, _ = ow.bench_many(cfg["theta0"][:1], n, 1, max_steps=n)
        rec.update({"cpu_1t_mh_steps_per_sec": n / secs, "cpu_1t_seconds": secs,
                    "gpu_over_cpu_1t": (steps / max(wall, 1e-9)) / (n / secs)})
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="C2")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the records of the other named configurations")
    ap.add_argument("--repeat-s", type=float, default=MIN_TIMED_S, help="device time to collect in K-step blocks (0: one block)")
    ap.add_argument("--shard", default="chains", choices=["chains", "datasets"],
                    help="N > 1: 'chains' = W more chains per GPU, no collective (weak scaling, the default); 'datasets' = a "
                         "global fit's data sets spread over the ranks, every rank holds all chains, per-step exchange of "
                         "the per-data-set terms (strong scaling; e.g. --config C4 --gpus 8)")
    args = ap.parse_args()
    K, Wm = args.steps, max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE %d" % (args.gpus, world))

    from lisp_mcmc_b200 import workloads
    walker_set = args.config == "NVSET"     # walker set with per-walker data: an HBM-bound kernel
    cfg_fn = workloads.nvset if walker_set else workloads.CONFIGS[args.config]
    base = cfg_fn()
    W_per_gpu = base["theta0"].shape[0]
    N = base["data"][1].shape[1] if walker_set else sum(len(d[0]) for d in base["data"])
    if walker_set and (args.impl == "reference" or args.gpus > 1):
        raise SystemExit("--config NVSET is a one-GPU kernel bench (no CPU arm)")
    config = {"workload": "%s: %s" % (args.config, cfg_fn.__doc__.strip().split("\n")[0]), "points": N, "shard": args.shard,
              "chains_per_gpu": W_per_gpu, "params": len(base["keys"]), "driver": "walker-adaptive-steps-full n=30000 "
              "temperature=10 auto=nil", "l2": "likelihood inputs (2.4 MB) are L2/SMEM-resident by design; chain state "
              "+ trace are streamed; an L2 flush (256 MB write) runs between warm-up and the timed steps"}

    # ---------------------------------------------------------------- reference arm: CPU restatement
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        r = cpu_reference_run(base, K, Wm, threads, budget_s=args.cpu_budget)
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K,
                "warmup": Wm, "ms_per_step": r["seconds"] / K * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "mh_steps_per_sec": r["chains"] * K / r["seconds"],
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "SBCL is not installable in this image: this is the C++ restatement of the reference algorithm "
                        "(oracle/), not the Lisp itself; the only SBCL datum is 5.3e6 point-evals/s (test.lisp:23)"}
        print(json.dumps(line))
        return

    # ---------------------------------------------------------------- B200 arm
    dev = Device(local_rank, world)
    from lisp_mcmc_b200 import _lib
    L = _lib.lib()
    by_dataset = args.shard == "datasets" and world > 1
    W_total = W_per_gpu if by_dataset else W_per_gpu * world
    full = cfg_fn(W=W_total) if (world > 1 and not by_dataset) else base
    lo = 0 if by_dataset else rank * W_per_gpu
    theta0 = np.ascontiguousarray(full["theta0"][lo:lo + W_per_gpu])
    make = workloads.make_walker_set if walker_set else workloads.make_walker
    n_owned = N
    if by_dataset:
        from lisp_mcmc_b200 import distributed as Dist
        owners = Dist.shard_datasets(len(base["function"]), world)
        make = lambda cfg_, **kw_: workloads.make_walker(cfg_, nccl=Dist.process_comm(rank, world), peers=Dist.peer_exchange(rank, world),
                                                         dataset_owners=owners, **kw_)
        n_owned = sum(len(d[0]) for d, o in zip(base["data"], owners) if o == rank)
    kw0 = {} if walker_set else {"theta0": theta0}
    # rows the ring must hold: warm-up + the repeated blocks (bounded by MAX_BLOCKS and by the run's 30000 steps)
    cap_rows = min(29500, Wm + (MAX_BLOCKS + 2) * K + 200)
    est_ms = 0.5 if args.config == "C2" else 20.0
    cap_rows = min(cap_rows, Wm + int((args.repeat_s * 1e3 / est_ms) * 3) + 3 * K + 1300)
    w = make(base, seed=2024, device=local_rank, chain_id_base=lo, trace_capacity=cap_rows, stream=dev.stream, **kw0)
    dfma, c_exp, c_log, c_sin, c_div = (C.c_double() for _ in range(5))
    w._chk(L.mfit_measure_fp64(w._h, C.byref(dfma), C.byref(c_exp), C.byref(c_log), C.byref(c_sin), C.byref(c_div)))
    costs = {"exp": c_exp.value, "log": c_log.value, "sin": c_sin.value, "div": c_div.value}
    sampler = ClockSampler(local_rank)
    sampler.start()
    run = device_run(dev, L, _lib, w, K, Wm, args.repeat_s)
    clocks = sampler.stop()
    w.close()

    # ---------------------------------------------------------------- e2e through the public API, host buffers
    cfgc = _lib.AdaptCfg(30000, 10.0, _lib.AUTO_NONE, _lib.OPT_COVARIANCE, 0, 0, 0)
    done, fin = C.c_int64(), C.c_int()

    def e2e_pass(steps, seed):
        t = [time.perf_counter()]
        w2 = make(base, seed=seed, device=local_rank, chain_id_base=lo, trace_capacity=steps + 1100, stream=dev.stream, **kw0)
        t.append(time.perf_counter())
        w2._chk(L.mfit_run_begin(w2._h, C.byref(cfgc), None, None, None, 0))
        t.append(time.perf_counter())
        w2._chk(L.mfit_run_advance(w2._h, steps, C.byref(done), C.byref(fin)))
        t.append(time.perf_counter())
        out = w2.trace(steps + 1)
        t.append(time.perf_counter())
        return w2, out, [1e3 * (b - a) for a, b in zip(t, t[1:])]

    for seed in (2025, 2027):          # untimed warm-up of the same path (first-use costs of the process, pool growth)
        w2, _, _ = e2e_pass(K, seed)
        w2.close()
    e2e_runs = []
    for rep in range(5):
        dev.barrier()
        e0, e1 = dev.events()
        t0 = time.perf_counter()
        e0.record()
        w2, (th, pr, nv), phases = e2e_pass(K, 2026)
        e1.record()
        dev.torch.cuda.synchronize()
        e2e_s = max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)
        tc = time.perf_counter()
        w2.close()
        phases.append(1e3 * (time.perf_counter() - tc))
        e2e_runs.append((dev.max_over_ranks([e2e_s * 1e3])[0], phases))
    e2e_runs.sort(key=lambda r: r[0])
    e2e_ms_med, phases = e2e_runs[len(e2e_runs) // 2]
    h2d = (base["data"][1].nbytes + 2 * 8 * N if walker_set else sum(3 * 8 * len(d[0]) for d in base["data"])) + theta0.nbytes
    d2h = th.nbytes + pr.nbytes + nv.nbytes

    line = None
    if rank == 0:
        blocks = run["block_ms"]
        ms_med = float(np.median(blocks))
        value = W_total * K * N / (ms_med * 1e-3)
        e2e_value = W_total * K * N / (e2e_ms_med * 1e-3)
        roof = roofline_of(base, run, W_per_gpu, n_owned, dfma.value, costs)
        hbm_peak, hbm_src = load_peaks()
        per_launch_s = roof["kernel_ms"] * 1e-3
        alg_bytes = 24.0 * N + 8.0 * (len(base["keys"]) + 1) * W_per_gpu   # one data pass + one theta/partial row per chain
        traffic, traffic_src = ncu_traffic() if args.config == "C2" else (None, None)
        roof.update({"traffic": traffic, "traffic_source": traffic_src,
                     "what": "FP64-pipe instructions executed (%.2f per point-eval: SASS count per loop variant weighted by the "
                             "tiles each ran, x2 flop) / kernel time, against the DFMA issue peak" % roof["fp64_instr_per_point_eval"],
                     "peak_source": "DFMA issue peak measured live by mfit_measure_fp64 (MEASURED_PEAKS.json has no FP64 figure)",
                     "kernel_share_of_step": (run["kernel_s"] / run["kernel_steps"]) / (ms_med * 1e-3 / K) if world == 1 else None,
                     "hbm": {"achieved": alg_bytes / per_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / per_launch_s / 1e9 / hbm_peak, "peak_source": hbm_src}})
        roof["survey_8d"].update({"dfma_equiv_cost": costs, "note": "algorithmic flops with exp/log/sin/div at libdevice's "
                                  "measured cost; above 1 where the kernel's own routines are cheaper"})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_med / K, "higher_is_better": True, "scaling": "strong" if by_dataset else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "mh_steps_per_sec": W_total * K / (ms_med * 1e-3),
                "blocks": {"n": len(blocks), "steps_each": K, "timed_s": sum(blocks) * 1e-3, "median_ms": ms_med,
                           "min_ms": min(blocks), "max_ms": max(blocks), "first_ms": blocks[0],
                           "note": "value / ms_per_step are the median K-step block of one continuing run"},
                "clocks": clocks, "gpu_launches": run["launches"],
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                        "ms": e2e_ms_med, "runs": len(e2e_runs),
                        "breakdown_ms": dict(zip(("walker_create", "run_begin", "run_%d_steps" % K, "trace_read", "close_untimed"), phases)),
                        "what": "walker_create (host x,y,sigma,theta0 -> HBM) + %d adaptive steps + trace read-back; median of %d" % (K, len(e2e_runs))},
                "roofline": roof, "mean_accept_200": run["accept"]}
        if walker_set:   # bound by HBM, not by the FP64 pipe: report that roofline as the primary one
            by = float(base["bytes_per_point"]) * W_per_gpu * N
            roof["fp64"] = {k: roof[k] for k in ("achieved", "peak", "frac", "unit")}
            roof.update({"bound": "hbm", "achieved": by / per_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": by / per_launch_s / 1e9 / hbm_peak, "traffic": None,
                         "what": "algorithmic bytes (%g per point-eval: y/sigma; x and 1/sigma are shared vectors) / kernel time, "
                                 "against the measured copy bandwidth (%s)" % (base["bytes_per_point"], hbm_src)})
            roof["kernel"] = "mfit_eval_perchain_reg_kernel"
            roof["peak_source"] = "MEASURED_PEAKS.json hbm_gbs (%s)" % hbm_src
            roof.pop("hbm", None)
            roof.pop("survey_8d", None)
            line["config"]["l2"] = "524 MB of per-walker data per pass (> 126 MB L2): streamed from HBM every step"

    # ---------------------------------------------------------------- the other named configurations
    extra = {}
    from lisp_mcmc_b200 import distributed as Dist
    if not args.no_extras and args.config == "C2" and not by_dataset:
        if world == 1:
            for name, steps, kw in (("C3", 12, {}), ("C4", 20, {}), ("C5", 6, {"W": 8192}), ("NVSET", 40, {})):
                try:
                    extra[name.lower()] = extra_record(dev, L, _lib, workloads, name, dfma.value, costs, local_rank, steps, **kw)
                except Exception as e:   # an extra must never take the headline down
                    extra[name.lower()] = {"error": "%s: %s" % (type(e).__name__, e)}
            try:
                extra["c1"] = c1_record(dev, L, _lib, workloads, local_rank, args.no_cpu)
            except Exception as e:
                extra["c1"] = {"error": "%s: %s" % (type(e).__name__, e)}
        else:
            # N > 1: the named multi-GPU shapes.  C3 = 16384 chains split over the ranks (strong), C5 = 8192 chains per GPU
            # (weak), C4 = the global fit with its data sets spread over the ranks (the one path with a per-step exchange)
            def sharded(name, steps, W_total_, W_local, lo_, scaling, **mk):
                cfgx = workloads.CONFIGS[name](W=W_total_)
                Nx = sum(len(d[0]) for d in cfgx["data"])
                th = np.ascontiguousarray(cfgx["theta0"][lo_:lo_ + W_local])
                wx = workloads.make_walker(cfgx, theta0=th, seed=5, device=local_rank, chain_id_base=lo_,
                                           trace_capacity=3 * steps + 1300, stream=dev.stream, **mk)
                rx = device_run(dev, L, _lib, wx, steps, 3, 0.0, kernel_steps=min(steps, 10))
                # the same chains started with :l-matrix diag(1e-3 |theta*|): the loop variant adapted chains run (see extra_record)
                ra = device_run(dev, L, _lib, wx, steps, 2, 0.0, kernel_steps=min(steps, 10),
                                l_matrix=np.diag(1e-3 * np.abs(cfgx["theta_star"])))
                wx.close()
                msx, msa = float(np.median(rx["block_ms"])), float(np.median(ra["block_ms"]))
                return {"workload": "%s: %s" % (name, workloads.CONFIGS[name].__doc__.strip().split("\n")[0]), "scaling": scaling,
                        "chains_total": W_total_, "chains_per_gpu": W_local, "points": Nx, "steps": steps, "ms_per_step": msx / steps,
                        "value": W_total_ * steps * Nx / (msx * 1e-3), "unit": UNIT,
                        "kernel_ms": rx["kernel_s"] / max(rx["kernel_launches"], 1) * 1e3, "gpu_launches": rx["launches"],
                        "adapted": {"ms_per_step": msa / steps, "value": W_total_ * steps * Nx / (msa * 1e-3), "unit": UNIT,
                                    "kernel_ms": ra["kernel_s"] / max(ra["kernel_launches"], 1) * 1e3}}
            try:
                Wc3 = 16384
                lo3, cnt3 = Dist.shard_chains(Wc3, rank, world)
                extra["c3_chains"] = sharded("C3", 12, Wc3, cnt3, lo3, "strong")
            except Exception as e:
                extra["c3_chains"] = {"error": "%s: %s" % (type(e).__name__, e)}
            try:
                extra["c5_chains"] = sharded("C5", 5, 8192 * world, 8192, 8192 * rank, "weak")
            except Exception as e:
                extra["c5_chains"] = {"error": "%s: %s" % (type(e).__name__, e)}
            try:
                c4 = workloads.c4()
                owners = Dist.shard_datasets(len(c4["function"]), world)
                w4 = workloads.make_walker(c4, nccl=Dist.process_comm(rank, world), peers=Dist.peer_exchange(rank, world),
                                           dataset_owners=owners, seed=5,
                                           device=local_rank, trace_capacity=1300, stream=dev.stream)
                r4 = device_run(dev, L, _lib, w4, 20, 3, 0.0, kernel_steps=10)
                w4.close()
                ms4 = float(np.median(r4["block_ms"]))
                N4 = sum(len(d[0]) for d in c4["data"])
                extra["c4_datasets"] = {"workload": "C4 global fit, %d data sets over %d GPUs, every rank holds all %d chains; "
                                        "per-step exchange of the per-data-set terms (%s)" % (
                                            len(owners), world, c4["theta0"].shape[0],
                                            "peer stores over NVLink + flags" if Dist.peer_exchange(rank, world) else "NCCL all-gather"),
                                        "scaling": "strong", "steps": 20, "ms_per_step": ms4 / 20,
                                        "value": c4["theta0"].shape[0] * 20 * N4 / (ms4 * 1e-3), "unit": UNIT,
                                        "kernel_ms": r4["kernel_s"] / max(r4["kernel_launches"], 1) * 1e3, "gpu_launches": r4["launches"]}
            except Exception as e:
                extra["c4_datasets"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        if extra:
            line["extra"] = extra
        if not args.no_cpu and world == 1 and not walker_set:
            threads = os.cpu_count() or 1
            r = cpu_reference_run(base, min(K, 50), 2, threads, budget_s=args.cpu_budget)
            r1 = cpu_reference_run(base, min(K, 50), 2, 1, budget_s=min(args.cpu_budget, 8.0))
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"],
                                    "cpu_1t": {"value": r1["value"], "unit": UNIT, "cores": 1, "sample": r1["sample"]},
                                    "sbcl_published": {"value": 5.3e6, "unit": UNIT, "source": "test.lisp:23 (6.34 s, 1e5 steps, 334 points)"}}
        print(json.dumps(line))
    if world > 1:
        dev.dist.destroy_process_group()


if __name__ == "__main__":
    main()
