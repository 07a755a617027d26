#!/usr/bin/env python
"""bench.py -- likelihood point-evals/s and MH steps/s of the walker-adaptive-steps hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2] [--impl reference]

A "step" is one Metropolis-Hastings step of every chain: one pass of the likelihood kernel over the whole
dataset for all W proposals, then accept/reject/adapt on the device (walker-take-step, mcmc-fitting.lisp:1072-1095,
inside the loop of walker-adaptive-steps-full :902-942).  The N=1 workload is BASELINE.json configs[1]
(C2: Gaussian + linear background, 100k points, 4096 walkers, log-normal); with N>1 every rank holds the dataset
and W more chains (weak scaling, Philox streams keyed by global chain id, no collective on the data path).

value   = W_total * K * N_points / (max-over-ranks device time of K steps), inputs resident in HBM.
e2e     = the same quantity through the public API with host buffers: walker_create (H2D of x, y, sigma, theta0),
          K adaptive steps, read-back of the K-step trace (D2H), all inside the timed region.
roofline= the likelihood kernel against the FP64 pipe: achieved = FP64 instructions the kernel executes per launch
          (counted in its SASS, 2 flop each) / its CUDA-event duration; peak = the DFMA issue rate measured live on
          this GPU (MEASURED_PEAKS.json has no FP64 figure).  SURVEY 8(d)'s algorithmic-flop accounting (each
          transcendental at libdevice's measured cost) is reported beside it -- see DESIGN.md.
cpu_baseline / --impl reference: the CPU restatement of the reference algorithm (oracle/, C++ -O2, compiled
          model formula -- a conservative stand-in for SBCL, which is not installable here) on the host cores.
"""
import argparse
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


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
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
        time.sleep(0.15)
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
    """The oracle's multi-threaded restatement on a bounded sample: W_s chains x `steps` adaptive steps on the full
    dataset.  W_s is sized from a probe so that the timed part stays near budget_s."""
    from oracle import oracle as orc
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle_walker
    ow = oracle_walker(orc, cfg, use_builtin_formula=True)
    N = sum(len(d[0]) for d in cfg["data"])
    th0 = cfg["theta0"]
    probe_w = max(threads, 1)
    secs, _ = ow.bench_many(th0[:probe_w], 30000, threads, max_steps=max(1, warmup))   # doubles as warm-up
    per_chain_step = secs * threads / (probe_w * max(1, warmup))
    w_s = int(budget_s / max(per_chain_step * steps / threads, 1e-9))
    w_s = max(threads, min(len(th0), (w_s // threads) * threads))
    secs, _ = ow.bench_many(th0[:w_s], 30000, threads, max_steps=steps)
    value = w_s * steps * N / secs
    return dict(value=value, seconds=secs, chains=w_s, steps=steps, threads=threads,
                sample="%d chains x %d adaptive steps x %d points (full dataset), chains split over %d threads"
                       % (w_s, steps, N, threads))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="C2")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--shard", default="chains", choices=["chains", "datasets"],
                    help="N > 1: 'chains' = W more chains per GPU, no collective (weak scaling, the default); 'datasets' = a "
                         "global fit's data sets spread over the ranks, every rank holds all chains, per-step NCCL "
                         "all-gather of the per-data-set terms (strong scaling; e.g. --config C4 --gpus 8)")
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
    make = workloads.make_walker_set if walker_set else workloads.make_walker
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
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    import ctypes as C
    from lisp_mcmc_b200 import _lib
    L = _lib.lib()
    by_dataset = args.shard == "datasets" and world > 1
    W_total = W_per_gpu if by_dataset else W_per_gpu * world
    full = cfg_fn(W=W_total) if (world > 1 and not by_dataset) else base
    lo = 0 if by_dataset else rank * W_per_gpu
    theta0 = np.ascontiguousarray(full["theta0"][lo:lo + W_per_gpu])
    if by_dataset:
        from lisp_mcmc_b200 import distributed as Dist
        owners = Dist.shard_datasets(len(base["function"]), world)
        # every walker set brings up its own communicator: an ncclUniqueId serves one ncclCommInitRank round only
        make = lambda cfg_, **kw_: workloads.make_walker(cfg_, nccl=(Dist.broadcast_nccl_unique_id(), rank, world),
                                                         dataset_owners=owners, **kw_)
        n_owned = sum(len(d[0]) for d, o in zip(base["data"], owners) if o == rank)
    tstream = torch.cuda.Stream()          # a real (non-default) stream shared by torch's events and the library
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    kw0 = {} if walker_set else {"theta0": theta0}
    w = make(base, seed=2024, device=local_rank, chain_id_base=lo, trace_capacity=Wm + K + 1200,
             stream=C.c_void_p(stream), **kw0)
    h = w._h
    dfma, c_exp, c_log, c_sin, c_div = (C.c_double() for _ in range(5))
    w._chk(L.mfit_measure_fp64(h, C.byref(dfma), C.byref(c_exp), C.byref(c_log), C.byref(c_sin), C.byref(c_div)))
    costs = {"exp": c_exp.value, "log": c_log.value, "sin": c_sin.value, "div": c_div.value}
    cfgc = _lib.AdaptCfg(30000, 10.0, _lib.AUTO_NONE, _lib.OPT_COVARIANCE, 0, 0, 0)
    w._chk(L.mfit_run_begin(h, C.byref(cfgc), None, None, None, 0))
    done = C.c_int64()
    fin = C.c_int()
    w._chk(L.mfit_run_advance(h, Wm, C.byref(done), C.byref(fin)))           # warm-up steps
    w._chk(L.mfit_flush_l2(h))
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    launches0 = L.mfit_launch_count(h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    w._chk(L.mfit_run_advance(h, K, C.byref(done), C.byref(fin)))            # the timed K steps
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    launches = L.mfit_launch_count(h) - launches0
    assert done.value == K, "ran %d steps instead of %d" % (done.value, K)
    # the likelihood kernel on its own: the run above advances the chains as two groups on two streams (one group's
    # per-chain kernels overlap the other's likelihood kernel), where an event interval around a kernel would include
    # its wait for SM slots; with timing on the library runs one group and records events around every eval launch
    Kk = min(K, 100)
    w._chk(L.mfit_set_timing(h, 1))
    w._chk(L.mfit_run_advance(h, Kk, C.byref(done), C.byref(fin)))
    ksec, klaunch = C.c_double(), C.c_int64()
    L.mfit_eval_kernel_time(h, C.byref(ksec), C.byref(klaunch))
    st = w.stats()
    assert np.all(st["steps_this_run"] == Wm + K + Kk)
    acc = float(np.mean(st["accept_200"]))
    w.close()

    # ---------------------------------------------------------------- e2e through the public API, host buffers
    def e2e_pass(steps, seed):
        w2 = make(base, seed=seed, device=local_rank, chain_id_base=lo, trace_capacity=steps + 1100,
                  stream=C.c_void_p(stream), **kw0)
        w2._chk(L.mfit_run_begin(w2._h, C.byref(cfgc), None, None, None, 0))
        w2._chk(L.mfit_run_advance(w2._h, steps, C.byref(done), C.byref(fin)))
        out = w2.trace(steps + 1)
        return w2, out

    w2, _ = e2e_pass(3, 2025)       # untimed warm-up of the same path (first-use costs of the process)
    w2.close()
    barrier()
    t0 = time.perf_counter()
    e0.record()
    w2, (th, pr, nv) = e2e_pass(K, 2026)
    e1.record()
    torch.cuda.synchronize()
    e2e_s = max(e0.elapsed_time(e1) * 1e-3, time.perf_counter() - t0)
    w2.close()
    h2d = (base["data"][1].nbytes + 2 * 8 * N if walker_set else sum(3 * 8 * len(d[0]) for d in base["data"])) + theta0.nbytes
    d2h = th.nbytes + pr.nbytes + nv.nbytes

    t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max = float(t[0]), float(t[1])
    if rank == 0:
        value = W_total * K * N / (ms_max * 1e-3)
        e2e_value = W_total * K * N / (e2e_ms_max * 1e-3)
        fpp = flops_per_point(base, costs)
        per_launch_s = ksec.value / max(klaunch.value, 1)
        evals = W_per_gpu * (n_owned if by_dataset else N) * Kk / max(klaunch.value, 1)   # point-evals per likelihood-kernel launch
        instr = float(base["fp64_instr_per_point"])
        achieved_tf = evals * instr * 2.0 / per_launch_s / 1e12
        survey_tf = evals * fpp / per_launch_s / 1e12
        hbm_peak, hbm_src = load_peaks()
        alg_bytes = 24.0 * N + 8.0 * (len(base["keys"]) + 1) * W_per_gpu   # one data pass + one theta/partial row per chain
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "strong" if by_dataset else "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "mh_steps_per_sec": W_total * K / (ms_max * 1e-3),
                "clocks": clocks, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / K, "d2h_bytes_per_step": d2h / K,
                        "what": "walker_create (host x,y,sigma,theta0 -> HBM) + %d adaptive steps + trace read-back" % K},
                "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": dfma.value / 1e12, "unit": "TFLOP/s",
                             "frac": achieved_tf / (dfma.value / 1e12),
                             # dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (profiles/r01_eval_kernel_ncu.md)
                             "traffic": 2638336 if args.config == "C2" else None,
                             "what": "FP64-pipe instructions executed (%.1f per point-eval, SASS count, x2 flop) / "
                                     "kernel time, against the DFMA issue peak" % instr,
                             "peak_source": "DFMA issue peak measured live by mfit_measure_fp64 (MEASURED_PEAKS.json has no FP64 figure)",
                             "kernel": "mfit_eval_kernel", "kernel_ms": per_launch_s * 1e3,
                             "kernel_share_of_step": (ksec.value / Kk) / (ms_max * 1e-3 / K) if world == 1 else None,
                             "survey_8d": {"flop_per_point_eval": fpp, "achieved": survey_tf,
                                           "frac": survey_tf / (dfma.value / 1e12), "dfma_equiv_cost": costs,
                                           "note": "algorithmic flops with exp/log/sin/div at libdevice's measured cost; "
                                                   "above 1 where the kernel's own routines are cheaper"},
                             "hbm": {"achieved": alg_bytes / per_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                     "frac": alg_bytes / per_launch_s / 1e9 / hbm_peak, "peak_source": hbm_src}},
                "mean_accept_200": acc}
        if walker_set:   # bound by HBM, not by the FP64 pipe: report that roofline as the primary one
            r = line["roofline"]
            by = float(base["bytes_per_point"]) * evals
            r["fp64"] = {k: r[k] for k in ("achieved", "peak", "frac", "unit")}
            r.update({"bound": "hbm", "achieved": by / per_launch_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                      "frac": by / per_launch_s / 1e9 / hbm_peak, "traffic": 549064960,
                      "what": "algorithmic bytes (%g per point-eval: y/sigma; x and 1/sigma are shared vectors) / kernel time, "
                              "against the measured copy bandwidth (%s)" % (base["bytes_per_point"], hbm_src)})
            r["kernel"] = "mfit_eval_perchain_kernel"
            r["peak_source"] = "MEASURED_PEAKS.json hbm_gbs (%s)" % hbm_src
            r.pop("hbm", None)
            r.pop("survey_8d", None)
            line["config"]["l2"] = "524 MB of per-walker data per pass (> 126 MB L2): streamed from HBM every step"
        if not args.no_cpu and world == 1 and not walker_set:
            threads = os.cpu_count() or 1
            r = cpu_reference_run(base, min(K, 50), 2, threads, budget_s=args.cpu_budget)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": threads, "kind": "port", "sample": r["sample"]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
