#!/usr/bin/env python
"""bench.py — atom-updates/s of the Gibbs-sampler hot path on the BASELINE.json workload.

Workload (config.workload): synthetic dense 20000x5000, nPatterns=20 (BASELINE.json configs[2], the
configuration the metric is quoted on; recipe SURVEY.md 8(d)).  A "step" is one MCMC iteration of the
hot path exactly as GapsRunner.cpp:294-296 drives it: nA ~ Poisson(atomsA), nP ~ Poisson(atomsP),
A.update(nA) -> P.sync(A) -> P.update(nP) -> A.sync(P).

  Two ways through update() are measured in every run, on the SAME chain state (--mode picks which one is `value`):
    sweep (default) : the row-parallel sweep (cgb_params.updateMode = CGB_UPDATE_SWEEP; sweep.cuh) — the north star's
           design: whole update() in one launch, one CTA per factor row, row staged once in shared memory, counter-based
           device-side draws, transport between adjacent rows.  A different chain from the reference's for the same seed
           (validated statistically against the reference, bit for bit against oracle/).
    exact : the reference's chain proposal for proposal (host generator + resident evaluator grid).  Its numbers are
           reported under "exact_mode" (or as `value` with --mode exact), measured at the same steady-state atom count:
           the chain is grown by the sweep (fast), then the samplers are switched over (cgb_sampler_set_update_mode).
  value  : atom-updates/s over K timed steps with D/S/AP/factors resident in HBM, after a ramp that grows
           the chain from zero atoms to its steady state (untimed) and W warm-up steps.
  e2e    : the same metric through the reference-facing entry point cgb_run (= gaps::run) on HOST buffers:
           upload of both data orientations, the whole two-phase run from zero atoms, statistics, download
           of Amean/Asd/Pmean/Psd — wall clock of the call.  `--impl reference` runs the reference's own
           gaps::run on the identical call (same matrix, seed, iterations) on the host cores and reports the
           updates per second of its SAMPLER LOOP (its own clock, GapsRunner.cpp:450,473) — the time it spends
           loading the matrix before the loop is quoted in `sample`, not counted.
  roofline: eval kernel, algorithmic bytes (16L/20L/32L read + 4L per changed AP row) over
           CUDA-event time of every launch, from a separate pass with per-launch events enabled.
  cpu_baseline: oracle/_ref (the unmodified reference, OpenMP, all host threads) on a bounded sample.

Multi-GPU (--gpus N under torchrun): the path does not shard inside one chain ("replicas only");
each rank runs an independent chain on its own shard-sized matrix (what distributed CoGAPS does per set),
no data-path collective, scaling = weak; time is the max over ranks.
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

G, S, K = 20000, 5000, 20
DATA_SEED = 20260117
CHAIN_SEED = 42
RAMP_ITERS = 400         # untimed sweep iterations growing the chain from zero atoms to its steady state before warm-up
E2E_ITERS = 100          # iterations per phase of the end-to-end / reference gaps::run call
EXACT_STEPS = 10         # timed exact-mode steps at the steady state (each ~35 ms there)


def make_data(g=G, s=S, k=K, seed=DATA_SEED, zero_fraction=0.0):
    """SURVEY 8(d) C3 recipe: noisy non-negative rank-k matrix, max < 50, fp32; C4: the same with
    `zero_fraction` of the entries zeroed uniformly at random."""
    rng = np.random.default_rng(seed)
    a0 = (rng.gamma(2.0, 0.5, (g, k)) * (rng.random((g, k)) < 0.3)).astype(np.float32)
    p0 = (rng.gamma(2.0, 0.5, (s, k)) * (rng.random((s, k)) < 0.3)).astype(np.float32)
    m = a0 @ p0.T
    noise = rng.standard_normal(m.shape, dtype=np.float32)
    d = np.maximum(m * (1.0 + 0.1 * noise), 0.0)
    d *= np.float32(40.0 / max(float(d.max()), 1e-9))
    del m, noise
    if zero_fraction > 0.0:
        for r0 in range(0, g, 4096):          # row blocks: keeps the mask's footprint small at 50000x30000
            blk = d[r0:r0 + 4096]
            blk[rng.random(blk.shape, dtype=np.float32) < zero_fraction] = 0.0
    return np.ascontiguousarray(d, dtype=np.float32)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def sweep_traffic(args):
    """DRAM bytes per sweep_kernel launch (dram__bytes_read.sum + dram__bytes_write.sum, one ncu --set full capture,
    profiles/sweep_kernel_traffic.json) — quoted only for the workload it was captured on"""
    path = os.path.join(ROOT, "profiles", "sweep_kernel_traffic.json")
    if (args.rows, args.cols, args.patterns) != (20000, 5000, 20) or not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f).get("dram_bytes_per_launch")


class ClockSampler(object):
    """SM clock and throttle reasons of the GPU, sampled from before the ramp (NVML every 10 ms; an nvidia-smi loop every
    50 ms where NVML cannot be had); the JSON line reports the samples that fall inside the timed region (plus the nearest
    one on either side when the region is short)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        """index: a GPU index, a comma-separated list of them (one nvidia-smi process watches them all), or None: no sampling
        (ranks other than 0 under torchrun — eight polling loops perturb the launches of all ranks)"""
        self.rows = []
        self.proc = None
        self.nvml = None
        self.t_begin = self.t_end = None
        if index is None:
            return
        # NVML in a thread of this process (what nvidia-smi itself reads), every 10 ms: the timed region is about a tenth
        # of a second, an nvidia-smi loop delivers one sample in that time.  nvidia-smi stays as the fallback.
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = (pynvml, pynvml.nvmlDeviceGetHandleByIndex(int(index)))
            self.nvml_stop = threading.Event()
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def _poll_nvml(self):
        nv, h = self.nvml
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = 0
        names = (("hw_slowdown", "HwSlowdown"), ("hw_thermal_slowdown", "HwThermalSlowdown"),
                 ("sw_thermal_slowdown", "SwThermalSlowdown"), ("sw_power_cap", "SwPowerCap"))
        while not self.nvml_stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                try:
                    watts = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                except Exception:
                    watts = 0.0
                flags = []
                for _, suffix in names:
                    bit = getattr(nv, "nvmlClocksEventReason" + suffix, None)
                    if bit is None:
                        bit = getattr(nv, "nvmlClocksThrottleReason" + suffix, 0)
                    flags.append("Active" if (mask & bit) else "Not Active")
                # the same row nvidia-smi would print for Q
                self.rows.append((time.perf_counter(), "%d, %d, %.2f, %s" % (sm, mx, watts, ", ".join(flags))))
            except Exception:
                pass
            self.nvml_stop.wait(0.01)

    def begin(self):
        self.t_begin = time.perf_counter()

    def end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.proc is None and self.nvml is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)   # let the sample that covers the end of the region arrive
        if self.nvml is not None:
            self.nvml_stop.set()
            self.thread.join(timeout=2)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        t1 = self.t_end if self.t_end is not None else float("inf")
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        before = [r for r in self.rows if r[0] < t0][-1:]
        after = [r for r in self.rows if r[0] > t1][:1]
        chosen = inside if len(inside) >= 2 else before + inside + after
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, row in chosen:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_region": len(inside),
                "source": "NVML, every 10 ms" if self.nvml is not None else "nvidia-smi -lms 50"}


class Chain(object):
    """Both samplers of one factorisation, driven like runOnePhase (GapsRunner.cpp:272-327)."""

    def __init__(self, data, k, seed, sparse=False, updateMode=0):
        import cogaps_b200 as cg
        from cogaps_b200._runhelp import make_params
        self.cg = cg
        self.params = make_params(nPatterns=k, seed=seed, useSparseOptimization=1 if sparse else 0)
        self.rs = cg.GapsRandomState(seed)
        # GapsRunner.cpp:402-406: A sampler sees the data transposed, P sampler as given
        self.A = cg.GibbsSampler(data, True, True, 0.01, 100.0, self.params, self.rs)
        self.P = cg.GibbsSampler(data, False, False, 0.01, 100.0, self.params, self.rs)
        self.rng = cg.GapsRng(self.rs)
        self.A.sync(self.P)
        self.P.sync(self.A)
        self.A.extraInitialization()
        self.P.extraInitialization()
        if updateMode:
            self.A.setUpdateMode(updateMode)
            self.P.setUpdateMode(updateMode)

    def step(self):
        nA = self.rng.poisson(float(max(self.A.nAtoms(), 10)))
        nP = self.rng.poisson(float(max(self.P.nAtoms(), 10)))
        self.A.update(nA)
        self.P.sync(self.A)
        self.P.update(nP)
        self.A.sync(self.P)
        return nA + nP

    def ramp(self, iters):
        for i in range(iters):
            temp = min(1.0, 2.0 * i / iters)      # annealing as in the equilibration phase
            self.A.setAnnealingTemp(temp)
            self.P.setAnnealingTemp(temp)
            self.step()
        self.A.setAnnealingTemp(1.0)
        self.P.setAnnealingTemp(1.0)


def run_reference(args, data):
    """The reference's own gaps::run (oracle/_ref, unmodified sources, OpenMP over all host threads)."""
    from oracle.harness import RefLib
    variant = "fast" if RefLib.available("fast") else "scalar"
    ref = RefLib(variant)
    # every core this process may run on (torchrun exports OMP_NUM_THREADS=1, which would otherwise clamp the
    # reference to one thread; it passes its thread count explicitly to every `omp parallel`)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    threads = max(threads, ref.max_threads())
    # warm-up: the first OpenMP region of a process creates the thread team (about a second on these boxes), which
    # would otherwise be charged to the reference's sampler loop
    ref.run(make_data(64, 48, 3, DATA_SEED), seed=1, nPatterns=3, nIterations=5, outputFrequency=0, maxThreads=threads)
    t0 = time.time()
    res = ref.run(data, seed=CHAIN_SEED, nPatterns=args.patterns, nIterations=args.e2e_iters, outputFrequency=0,
                  maxThreads=threads, useSparseOptimization=1 if args.sparse else 0)
    # The metric is atom updates per second of SAMPLER-LOOP time (what GapsResult::totalRunningTime covers,
    # GapsRunner.cpp:450,473): the reference's own clock readings give that interval with sub-second resolution.  The
    # time gaps::run spends before the loop (Matrix copies, sampler constructors) is not the path and is reported
    # beside it, not inside it.
    call = res.totalRunningTime
    wall = res.secondsSamplerLoop if res.secondsSamplerLoop > 0 else call
    value = res.totalUpdates / wall
    sample = ("one gaps::run call, %d iterations/phase from zero atoms, %d atom updates; %.1f s in the sampler loop "
              "(the value), %.1f s loading before it, %.1f s for the whole gaps::run call = %.0f updates/s, %.1f s incl. "
              "host matrix conversion; build: oracle/_ref %s, %d OpenMP threads"
              % (args.e2e_iters, res.totalUpdates, wall, res.secondsLoading, call, res.totalUpdates / call,
                 time.time() - t0, variant, threads))
    return value, wall, int(res.totalUpdates), threads, sample


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ramp", type=int, default=RAMP_ITERS)
    ap.add_argument("--e2e-iters", type=int, default=E2E_ITERS)
    ap.add_argument("--rows", type=int, default=G)
    ap.add_argument("--cols", type=int, default=S)
    ap.add_argument("--patterns", type=int, default=K)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sparse", action="store_true",
                    help="BASELINE.json configs[3]: sparseOptimization (SparseGibbsSampler) on a matrix with --zeros of its "
                         "entries zeroed; give the shape with --rows/--cols/--patterns (50000 30000 50 for C4)")
    ap.add_argument("--zeros", type=float, default=0.95)
    ap.add_argument("--mode", default="sweep", choices=["sweep", "exact"],
                    help="which path through update() the headline `value` / `e2e` / `roofline` describe (both are measured)")
    ap.add_argument("--exact-steps", type=int, default=EXACT_STEPS)
    ap.add_argument("--no-c5", action="store_true", help="N > 1: skip the configs[4] record (sharded 200000x30000 + P all-gather)")
    ap.add_argument("--c5-genes", type=int, default=30000)
    ap.add_argument("--c5-cells", type=int, default=200000)
    ap.add_argument("--c5-patterns", type=int, default=50)
    ap.add_argument("--c5-ramp", type=int, default=150)
    ap.add_argument("--c5-mode", default="sweep", choices=["sweep", "exact"])
    ap.add_argument("--c5-steps", type=int, default=5)
    ap.add_argument("--chains", type=int, default=0,
                    help="extra leg: this many independent chains on ONE GPU, one host thread each, every resident grid "
                         "taking 1/chains of the device (what distributed CoGAPS does with several sets per worker)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "synthetic dense %dx%d nPatterns=%d" % (args.rows, args.cols, args.patterns)
    if args.sparse:
        workload = "synthetic sparse %dx%d (%.0f%% zeros) nPatterns=%d" % (args.rows, args.cols, 100 * args.zeros, args.patterns)
    config = {"workload": workload, "data_seed": DATA_SEED, "chain_seed": CHAIN_SEED,
              "sampler": "asynchronous, sparse normal model (sparseOptimization)" if args.sparse
              else "asynchronous, dense normal model, default uncertainty",
              "step": "one MCMC iteration: A.update(Poisson(atomsA)) + P.sync + P.update(Poisson(atomsP)) + A.sync",
              "ramp_iterations": args.ramp,
              "device_mode": {"sweep": "one sweep_kernel launch per update(): one CTA per factor row, rows handed out longest chain first; "
                                       "a row of up to 10240 floats keeps its D and AP lines in shared memory (4 rows per SM), a longer one "
                                       "its AP line only (D through L2, 2 rows per SM); then one sweep_transport_kernel launch",
                              "exact": "resident grid per update(); each proposal streamed to its cluster through pinned host memory as it "
                                       "is generated; per-row commit versions instead of grid barriers"
                                       if os.environ.get("COGAPS_PERSISTENT", "1") != "0" else "one eval-kernel launch per batch"},
              "l2": ("sparse model: per-proposal traffic is gathers of factor rows (k floats) at the row's non-zeros; "
                     "the CSR rows + both factors exceed L2 at 50000x30000, no flush needed") if args.sparse
              else "inputs larger than L2 (126 MB): 1.6 GB of resident D / AP in two orientations; every update() walks all rows of one "
                   "orientation (0.8 GB in, the rewritten AP lines out) and every sync rewrites the other's 0.4 GB — nothing a step reads "
                   "is left in L2 from the step before, no flush needed",
              "parallelism": "replicas x%d (one chain per GPU on its own copy of the matrix, same seeds: per-GPU work identical; no data-path collective)" % world
              if world > 1 else "single chain"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        data = make_data(args.rows, args.cols, args.patterns, DATA_SEED, args.zeros if args.sparse else 0.0)
        value, wall, updates, threads, sample = run_reference(args, data)
        line = {"impl": "reference", "metric": "atom_updates_per_s", "value": value, "unit": "atom-updates/s",
                "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": wall * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "atom-updates/s", "cores": threads, "kind": "reference",
                                 "sample": sample},
                "e2e": {"value": value, "unit": "atom-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import cogaps_b200 as cg
    from cogaps_b200._lib import check
    check(cg.lib().cgb_set_device(local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max_sum(elapsed, sums):
        """max over ranks of the device time, sum over ranks of the counts"""
        if world == 1:
            return elapsed, [float(x) for x in sums]
        t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        u = torch.tensor([float(x) for x in sums], dtype=torch.float64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        return float(t.item()), [float(x) for x in u.tolist()]

    def timed_steps(chain, steps, clock=None):
        """exactly `steps` iterations between barriers; device time from a CUDA event pair around them"""
        chain.A.resetCounters()
        chain.P.resetCounters()
        launches0 = cg.lib().cgb_kernel_launch_count()
        barrier()
        if clock is not None:
            clock.begin()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        ev0.record()
        asked = 0
        for _ in range(steps):
            asked += chain.step()
        torch.cuda.synchronize()      # every update() ends with a stream synchronise; the last A.sync(P) may still run
        ev1.record()
        barrier()
        host = time.perf_counter() - t0
        if clock is not None:
            clock.end()
        cA, cP = chain.A.counters(), chain.P.counters()
        return {"elapsed": ev0.elapsed_time(ev1) * 1e-3, "host": host, "asked": asked,
                "made": int(cA.nProposalsTotal + cP.nProposalsTotal), "cA": cA, "cP": cP,
                "launches": int(cg.lib().cgb_kernel_launch_count() - launches0), "steps": steps}

    dense = not args.sparse
    mode = args.mode
    config["update_mode"] = ("sweep: row-parallel (one CTA per factor row, whole update() in one launch, Philox draws, transport "
                             "between adjacent rows); a different chain from the reference's for the same seed — validated "
                             "statistically against it and bit for bit against oracle/; the reference's own chain is measured "
                             "in the same run under exact_mode") if mode == "sweep" else \
                            "exact: the reference's chain proposal for proposal (host generator + resident evaluator grid)"
    # every rank factorises the same synthetic matrix with the same chain seed: weak scaling with the per-GPU work exactly
    # fixed (with per-rank seeds the replicas' atom counts, and with them the work per step, differed by several per cent
    # and the max over ranks measured the busiest replica, not the hardware)
    data = make_data(args.rows, args.cols, args.patterns, DATA_SEED, args.zeros if args.sparse else 0.0)
    # rank 0 watches its own GPU (one polling loop per job: a query over all eight GPUs takes longer than the timed region)
    clocks = ClockSampler(local_rank) if rank == 0 else ClockSampler(None)
    t_setup = time.time()
    # the chain grows to its steady state in sweep mode (seconds instead of minutes), untimed
    chain = Chain(data, args.patterns, CHAIN_SEED, sparse=args.sparse, updateMode=1)
    chain.ramp(args.ramp)
    for _ in range(args.warmup):
        chain.step()
    setup_s = time.time() - t_setup
    peak, peak_src = load_peaks()
    L_A, L_P = args.cols, args.rows

    # ---- timed region 1: exactly K steps of the sweep ----
    sweep = None
    if True:
        r = timed_steps(chain, args.steps, clocks if mode == "sweep" else None)
        el, (made, launches) = reduce_max_sum(r["elapsed"], [r["made"], r["launches"]])
        cA, cP = r["cA"], r["cP"]
        per_rank = None
        if world > 1:
            # every rank's own step time and kernel time: one straggler, or all of them slower than a lone GPU?
            mine = torch.tensor([r["elapsed"] / args.steps * 1e3, (cA.secondsKernel + cP.secondsKernel) / args.steps * 1e3, r["host"] / args.steps * 1e3],
                                dtype=torch.float64, device="cuda")
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            per_rank = {"ms_per_step": [float(t[0]) for t in allr], "kernel_ms_per_step": [float(t[1]) for t in allr],
                        "host_ms_per_step": [float(t[2]) for t in allr]}
        ktime = cA.secondsKernel + cP.secondsKernel
        kbytes = cA.algorithmicBytes + cP.algorithmicBytes
        sweep = {"value": made / el, "unit": "atom-updates/s", "ms_per_step": el / args.steps * 1e3, "steps": args.steps,
                 "atom_updates": int(made), "gpu_launches": int(launches),
                 "atoms": {"A": int(chain.A.nAtoms()), "P": int(chain.P.nAtoms())},
                 "host_ms_per_step": r["host"] / args.steps * 1e3,
                 "kernel_ms_per_step": {"A": cA.secondsKernel / args.steps * 1e3, "P": cP.secondsKernel / args.steps * 1e3},
                 "per_rank": per_rank,
                 "roofline": {"bound": "hbm", "kernel": ("sweep_kernel" if dense else "sweep_sparse_kernel") + " + sweep_transport_kernel (one update() = one launch of each)",
                              "achieved": kbytes / max(ktime, 1e-12) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": kbytes / max(ktime, 1e-12) / 1e9 / peak, "peak_source": peak_src,
                              "algorithmic_bytes_per_launch": kbytes / (2.0 * args.steps),
                              "avg_launch_us": ktime / (2.0 * args.steps) * 1e6, "launches": 2 * args.steps,
                              "traffic": sweep_traffic(args) if dense else None,
                              "A_side": {"GBps": cA.algorithmicBytes / max(cA.secondsKernel, 1e-12) / 1e9},
                              "P_side": {"GBps": cP.algorithmicBytes / max(cP.secondsKernel, 1e-12) / 1e9},
                              "how": "cudaEvent pair on the launching stream around the row sweep + transport launches of every "
                                     "update() in the timed region; bytes = SURVEY 8(d) algorithmic bytes of the proposals evaluated "
                                     "(16 L per birth/death scan, 20 L per same-row move/exchange, 32 L per two-row one, +4 L per "
                                     "rewritten AP line).  The kernel stages a row once and serves all its proposals from shared "
                                     "memory, so the bytes it really moves are far fewer (traffic_estimate; ncu capture in profiles/) "
                                     "and a fraction above 1 is possible: the sweep is bound by instruction issue and the serial "
                                     "decision of each proposal, not by HBM"}}
        # what the sweep really moves per update(): every active row's D and AP lines once in, dirty AP lines once out
        lines_in = 2.0
        if dense:
          sweep["roofline"]["traffic_estimate"] = {
            "bytes_per_launch": (args.rows * L_A + args.cols * L_P) * 4.0 * (lines_in + 1.0) / 2.0,
            "how": "upper bound: (D + AP read, AP written) x every row of both samplers, per update() launch; measured per "
                   "launch by ncu in profiles/r2_sweep_kernel_ncu_final.csv (roofline.traffic)"}
          sweep["roofline"]["dram_frac_estimate"] = sweep["roofline"]["traffic_estimate"]["bytes_per_launch"] / \
            max(ktime / (2.0 * args.steps), 1e-12) / 1e9 / peak

    # ---- timed region 2: the reference's own chain (exact mode) from the same state ----
    chain.A.setUpdateMode(0)
    chain.P.setUpdateMode(0)
    for _ in range(args.warmup):
        chain.step()
    exact_steps = args.steps if (mode == "exact") else min(args.steps, args.exact_steps)
    r = timed_steps(chain, exact_steps, clocks if mode == "exact" else None)
    el, (made, launches) = reduce_max_sum(r["elapsed"], [r["asked"], r["launches"]])
    cA, cP = r["cA"], r["cP"]
    queued = cA.nProposalsQueued + cP.nProposalsQueued
    batches = cA.nBatches + cP.nBatches
    # every proposal: one 64-byte task record per CTA of its cluster written to pinned host memory and pulled by
    # the device (two-row moves / exchanges send two), one 16-byte outcome record written back
    segA, segP = chain.A.reductionOrder()[2], chain.P.reductionOrder()[2]
    h2d_step = (cA.nProposalsQueued * 64.0 * segA + cP.nProposalsQueued * 64.0 * segP) / exact_steps
    d2h_step = queued * 16.0 / exact_steps
    exact = {"value": made / el, "unit": "atom-updates/s", "ms_per_step": el / exact_steps * 1e3, "steps": exact_steps,
             "atom_updates": int(made), "gpu_launches": int(launches),
             "atoms": {"A": int(chain.A.nAtoms()), "P": int(chain.P.nAtoms())},
             "host_ms_per_step": r["host"] / exact_steps * 1e3,
             "batches_per_step": batches / exact_steps, "proposals_per_batch": queued / max(batches, 1),
             "host_generate_s_per_step": (cA.secondsHostGenerate + cP.secondsHostGenerate) / exact_steps,
             "device_wait_s_per_step": (cA.secondsDeviceWait + cP.secondsDeviceWait) / exact_steps,
             "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step}
    clock_info = clocks.stop()

    # ---- exact-mode roofline: the eval kernel of the timed region itself.  The resident kernel is launched once per
    # update() (2 per step); its duration comes from CUDA events on its stream, its bytes are the algorithmic
    # bytes of the proposals it evaluated.  The duration includes the time the grid waits for the
    # host generator: that is the launch as it runs in the product.
    chisq_ms = None
    timings = {}
    if rank == 0:
        resident_bytes = cA.algorithmicBytes + cP.algorithmicBytes
        resident_time = cA.secondsKernel + cP.secondsKernel
        n_launch = 2 * exact_steps
        # the same device code launched once per conflict-free batch, each launch bracketed by CUDA events
        for smp in (chain.A, chain.P):
            smp.setPersistent(False)
            smp.setKernelTiming(True)
            smp.resetCounters()
        for _ in range(2):
            chain.step()
        rA, rP = chain.A.counters(), chain.P.counters()
        for smp in (chain.A, chain.P):
            smp.setKernelTiming(False)
            smp.setPersistent(True)
        bytes_total = rA.algorithmicBytes + rP.algorithmicBytes
        ktime = rA.secondsKernel + rP.secondsKernel
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "eval_kernel_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
        resident = os.environ.get("COGAPS_PERSISTENT", "1") != "0"
        per_batch = {"kernel": "eval_kernel (one launch per conflict-free batch; same device code)",
                     "achieved": bytes_total / ktime / 1e9, "frac": bytes_total / ktime / 1e9 / peak,
                     "algorithmic_bytes_per_launch": bytes_total / max(rA.nBatches + rP.nBatches, 1),
                     "avg_launch_us": ktime / max(rA.nBatches + rP.nBatches, 1) * 1e6,
                     "traffic": traffic,
                     "A_side": {"GBps": rA.algorithmicBytes / max(rA.secondsKernel, 1e-12) / 1e9,
                                "avg_launch_us": rA.secondsKernel / max(rA.nBatches, 1) * 1e6,
                                "proposals_per_launch": rA.nProposalsQueued / max(rA.nBatches, 1)},
                     "P_side": {"GBps": rP.algorithmicBytes / max(rP.secondsKernel, 1e-12) / 1e9,
                                "avg_launch_us": rP.secondsKernel / max(rP.nBatches, 1) * 1e6,
                                "proposals_per_launch": rP.nProposalsQueued / max(rP.nBatches, 1)},
                     "how": "cudaEvent pair around every launch on the launching stream, 2 steps after the timed region"}
        if resident and resident_time > 0:
            achieved = resident_bytes / resident_time / 1e9
            roofline = {"bound": "hbm", "kernel": "eval_stream_kernel (resident grid: alphaParameters scan + epilogue + AP commit)",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "peak_source": peak_src, "traffic": None,
                        "algorithmic_bytes_per_launch": resident_bytes / n_launch,
                        "avg_launch_us": resident_time / n_launch * 1e6, "launches": n_launch,
                        "how": "timed region: one launch per update(), cudaEvent pair on its stream; bytes = SURVEY 8(d) "
                               "algorithmic bytes of the proposals evaluated; includes the time the grid waits for the host generator",
                        "per_batch_launch": per_batch}
        else:
            roofline = {"bound": "hbm", "kernel": per_batch["kernel"], "achieved": per_batch["achieved"], "peak": peak,
                        "unit": "GB/s", "frac": per_batch["frac"], "peak_source": peak_src, "traffic": traffic,
                        "algorithmic_bytes_per_launch": per_batch["algorithmic_bytes_per_launch"],
                        "avg_launch_us": per_batch["avg_launch_us"], "how": per_batch["how"],
                        "A_side": per_batch["A_side"], "P_side": per_batch["P_side"]}
        # ---- the scan kernel with enough work: alphaParameters probes (DenseNormalModel.cpp:162-240 through the same
        # staging + scan + reduce code, no proposal epilogue), every row once, ONE launch (probe_kernel), CUDA events
        # around it.  This is what the kernel sustains when the generator is not the limit.
        if dense:
            rngq = np.random.default_rng(5)
            sat = {}
            for nm, smp, nrows in (("A", chain.A, args.rows), ("P", chain.P, args.cols)):
                rows_q = rngq.permutation(nrows)
                nq = rows_q.size
                variant = (rngq.random(nq) < 0.5).astype(np.float64)       # half single-column, half same-row pairs
                c1 = rngq.integers(0, args.patterns, nq)
                c2 = (c1 + 1 + rngq.integers(0, args.patterns - 1, nq)) % args.patterns
                q = np.stack([variant, rows_q, c1, rows_q, c2, np.zeros(nq)], axis=1)
                smp.alphaParameters(q[:1024])                                # warm-up launch
                smp.setKernelTiming(True)
                smp.resetCounters()
                smp.alphaParameters(q)
                c = smp.counters()
                smp.setKernelTiming(False)
                L = args.cols if nm == "A" else args.rows
                nbytes = float(((variant == 0) * 16.0 + (variant == 1) * 20.0).sum()) * L
                sat[nm] = {"GBps": nbytes / max(c.secondsKernel, 1e-12) / 1e9, "launches": int(c.nBatches),
                           "avg_launch_us": c.secondsKernel / max(c.nBatches, 1) * 1e6, "tasks": int(nq), "row_length": int(L)}
            tot_b = sum(v["GBps"] * v["avg_launch_us"] * v["launches"] for v in sat.values())
            tot_t = sum(v["avg_launch_us"] * v["launches"] for v in sat.values())
            roofline["scan_saturated"] = {"achieved": tot_b / tot_t, "frac": tot_b / tot_t / peak, "A_side": sat["A"], "P_side": sat["P"],
                                          "how": "cgb_sampler_alpha_parameters: the eval kernel's staging + scan + reduce on one "
                                                 "(row, column) probe per row, all rows in one launch, cudaEvent pair around it; "
                                                 "bytes = 16 L (one column) or 20 L (two columns of one row) per probe"}
        exact["roofline"] = roofline
        # ---- chi-sq wall time (second half of BASELINE.json's metric) and the other whole-matrix passes of the path ----
        def wall_ms(fn, n=5):
            fn()
            torch.cuda.synchronize()
            t = time.perf_counter()
            for _ in range(n):
                fn()
            torch.cuda.synchronize()
            return (time.perf_counter() - t) / n * 1e3
        chisq_ms = wall_ms(chain.P.chiSq)
        nbytes = 8.0 * args.rows * args.cols            # D and AP once (S is derived)
        timings = {"chisq_ms": chisq_ms, "chisq_frac_of_hbm_peak": nbytes / (chisq_ms * 1e-3) / 1e9 / peak if dense else None,
                   "sync_ms": wall_ms(lambda: chain.A.sync(chain.P)) if dense else None,
                   "extra_init_ms": wall_ms(chain.P.extraInitialization, 3) if dense else None}
        if dense:
            timings["sync_frac_of_hbm_peak"] = nbytes / (timings["sync_ms"] * 1e-3) / 1e9 / peak
            timings["extra_init_frac_of_hbm_peak"] = 4.0 * args.rows * args.cols / (timings["extra_init_ms"] * 1e-3) / 1e9 / peak
        # GapsStatistics::meanChiSq (GapsStatistics.cpp:63-87): chi-square of the posterior-mean factors against D, every element
        # of D read once, the k-term products formed on the fly
        stats = cg.GapsStatistics(args.rows, args.cols, args.patterns)
        stats.update(chain.A, chain.P)
        timings["mean_chisq_ms"] = wall_ms(lambda: stats.meanChiSq(chain.P), 3)
        timings["mean_chisq_frac_of_hbm_peak"] = 4.0 * args.rows * args.cols / (timings["mean_chisq_ms"] * 1e-3) / 1e9 / peak
        del stats
    del chain

    # ---- several chains sharing the device (the generator of ONE exact-mode chain cannot keep a B200 busy) ----
    multi = None
    if rank == 0 and args.chains > 1:
        check(cg.lib().cgb_set_resident_share(args.chains))
        chains = [Chain(data, args.patterns, CHAIN_SEED + 100 + i, sparse=args.sparse) for i in range(args.chains)]
        gate = threading.Barrier(args.chains + 1)
        done_updates = [0] * args.chains

        def drive(i):
            chains[i].ramp(min(args.ramp, 50))
            for _ in range(args.warmup):
                chains[i].step()
            gate.wait()
            n = 0
            for _ in range(exact_steps):
                n += chains[i].step()
            done_updates[i] = n
            gate.wait()

        workers = [threading.Thread(target=drive, args=(i,)) for i in range(args.chains)]
        for w in workers:
            w.start()
        gate.wait()
        tm0 = time.perf_counter()
        gate.wait()
        torch.cuda.synchronize()
        tm = time.perf_counter() - tm0
        for w in workers:
            w.join()
        multi = {"chains": args.chains, "value": sum(done_updates) / tm, "unit": "atom-updates/s (sum over chains)",
                 "ms_per_step": tm / exact_steps * 1e3, "per_chain": [u / tm for u in done_updates],
                 "how": "independent exact-mode chains on one GPU, one host thread each, cgb_set_resident_share(chains)"}
        del chains
        check(cg.lib().cgb_set_resident_share(1))

    # ---- end to end: cgb_run (= gaps::run) on host buffers, both modes ----
    def e2e_of(update_mode, n_calls):
        # identical calls (same seed, same chain), the median by wall clock: the call is about a second and its
        # allocation / teardown share varies from box to box and run to run
        calls = []
        for _ in range(n_calls):
            t0 = time.perf_counter()
            res = cg.gaps_run(data, seed=CHAIN_SEED, nPatterns=args.patterns, nIterations=args.e2e_iters,
                              outputFrequency=0, maxThreads=1, useSparseOptimization=1 if args.sparse else 0,
                              updateMode=update_mode)
            calls.append((time.perf_counter() - t0, float(res.totalRunningTime), int(res.totalUpdates)))
        wall, loop_s, upd = sorted(calls)[len(calls) // 2]
        n_it = 2 * args.e2e_iters
        upload = 2.0 * data.nbytes
        results = 4.0 * 2 * (args.rows + args.cols) * args.patterns
        per_update_h2d = (h2d_step * exact_steps / max(exact["atom_updates"], 1)) if update_mode == 0 else 0.0
        per_update_d2h = 16.0 if update_mode == 0 else 0.0
        return {"value": upd / wall, "unit": "atom-updates/s",
                "h2d_bytes_per_step": (upload + per_update_h2d * upd) / n_it + (0.0 if update_mode == 0 else 2 * 512.0),
                "d2h_bytes_per_step": (results + per_update_d2h * upd) / n_it + (0.0 if update_mode == 0 else 2 * 72.0),
                "call": "cgb_run (gaps::run): host fp32 matrix in, Amean/Asd/Pmean/Psd out; updateMode=%d" % update_mode,
                "iterations_per_phase": args.e2e_iters, "atom_updates": upd, "wall_s": wall,
                "wall_s_of_each_call": [c[0] for c in calls], "how": "median of %d identical calls" % n_calls,
                # the part of the call the reference arm's value covers (its sampler loop, GapsRunner.cpp:450,473)
                "sampler_loop_s": loop_s, "sampler_loop_value": upd / max(loop_s, 1e-9)}

    if rank == 0 and not args.no_e2e:
        sweep["e2e"] = e2e_of(1, 5)
        exact["e2e"] = e2e_of(0, 3)

    # ---- BASELINE.json configs[4] on N > 1 GPUs: every rank runs its 200000/N x 30000 k=50 sparse shard, then the NCCL
    # all-gather of the per-shard P rows through the C ABI (cgb_comm_init / cgb_allgather_rows) ----
    c5 = None
    if world > 1 and not args.no_c5:
        c5 = run_c5(args, rank, world, local_rank, barrier, reduce_max_sum)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, wall, upd, threads, sample = run_reference(args, data)
        cpu_baseline = {"value": v, "unit": "atom-updates/s", "cores": threads, "kind": "reference", "sample": sample}

    if rank == 0:
        head = sweep if mode == "sweep" else exact
        line = {"metric": "atom_updates_per_s", "value": head["value"], "unit": "atom-updates/s", "n_gpus": args.gpus,
                "steps": head["steps"], "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "clocks": clock_info, "e2e": head.get("e2e"),
                "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "cpu_baseline": cpu_baseline,
                "mode": mode, "atoms": head["atoms"], "setup_s": setup_s,
                "chisq_ms": chisq_ms, "whole_matrix_passes": timings,
                "timer": "CUDA events around the K timed steps (device clock), max over ranks; host perf_counter over the "
                         "same region: %.3f ms per step" % head["host_ms_per_step"]}
        line["sweep_mode"] = sweep
        line["exact_mode"] = exact
        # kept at top level for continuity with round 1 (exact mode)
        line["host_generate_s_per_step"] = exact["host_generate_s_per_step"]
        line["device_wait_s_per_step"] = exact["device_wait_s_per_step"]
        if multi is not None:
            line["multi_chain"] = multi
        if c5 is not None:
            line["c5"] = c5
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_c5(args, rank, world, local_rank, barrier, reduce_max_sum):
    """BASELINE.json configs[4]: distributed scCoGAPS row-shard of 200000 x 30000, nPatterns=50 (sparse model) over the
    ranks, then the all-gather of the per-shard P rows (stitchTogether, R/DistributedCogaps.R:226-278) through the C ABI's
    own NCCL communicator, straight from the samplers' device matrices."""
    import torch
    import torch.distributed as dist
    import cogaps_b200 as cg
    genes, cells_total, k = args.c5_genes, args.c5_cells, args.c5_patterns
    base = cells_total // world
    sizes = [base] * (world - 1) + [cells_total - base * (world - 1)]      # R/SubsetData.R:63-75,90
    cells = sizes[rank]
    t0 = time.time()
    data = make_data(genes, cells, k, DATA_SEED + 1000 + rank, 0.95)        # genes x this rank's cells (C4 recipe)
    chain = Chain(data, k, CHAIN_SEED + rank, sparse=True, updateMode=1 if args.c5_mode == "sweep" else 0)
    del data
    chain.ramp(args.c5_ramp)
    setup_s = time.time() - t0
    chain.A.resetCounters()
    chain.P.resetCounters()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    asked = 0
    for _ in range(args.c5_steps):
        asked += chain.step()
    torch.cuda.synchronize()
    ev1.record()
    barrier()
    elapsed = ev0.elapsed_time(ev1) * 1e-3
    if args.c5_mode == "sweep":
        # the sweep rounds each row's share of the proposals stochastically: count the proposals actually made
        asked = int(chain.A.counters().nProposalsTotal + chain.P.counters().nProposalsTotal)
    # the C ABI's communicator: rank 0 makes the id, the launcher's process group carries it to the others
    ids = [cg.Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    comm = cg.Comm(ids[0], rank, world)
    comm.allgatherRows(chain.P, sizes)                                      # warm-up: communicator channels, staging buffers
    barrier()
    full = comm.allgatherRows(chain.P, sizes)                               # (cells_total, k) on the host of every rank
    gather_ms = comm.last_ms
    mine = chain.P.getMatrix()
    row0 = sum(sizes[:rank])
    ok = bool(np.array_equal(full[row0:row0 + cells], mine))
    checksum = float(full.astype(np.float64).sum())
    own = float(mine.astype(np.float64).sum())
    el, (updates, own_sum, oks) = reduce_max_sum(elapsed, [asked, own, 1.0 if ok else 0.0])
    gm, _ = reduce_max_sum(gather_ms, [0.0])
    atoms = {"A": int(chain.A.nAtoms()), "P": int(chain.P.nAtoms())}
    del chain, comm
    if rank != 0:
        return None
    return {"workload": "synthetic sparse %dx%d (95%% zeros) nPatterns=%d, cells sharded over %d GPUs (%s per GPU)"
                        % (genes, cells_total, k, world, sizes),
            "value": updates / el, "unit": "atom-updates/s (sum over shards)", "ms_per_step": el / args.c5_steps * 1e3,
            "steps": args.c5_steps, "ramp_iterations": args.c5_ramp, "per_rank_updates_per_s": updates / el / world,
            "sampler": "asynchronous, sparse normal model, %s mode" % args.c5_mode, "atoms_rank0": atoms, "setup_s_rank0": setup_s,
            "allgather": {"what": "per-shard P rows (cells x nPatterns): the sparse model's row copy, straight from the samplers' device memory",
                          "api": "cgb_comm_init + cgb_allgather_rows (ncclAllGather on the library's own stream)",
                          "allgather_us": gm * 1e3, "bytes_per_rank": int(max(sizes) * ((k + 3) // 4 * 4) * 4),
                          "bytes_total": int(max(sizes) * ((k + 3) // 4 * 4) * 4 * world),
                          "every_rank_found_its_own_rows_in_place": bool(oks == world),
                          "checksum_matches_sum_of_shards": bool(abs(checksum - own_sum) <= 1e-6 * max(1.0, abs(own_sum)))},
            "scaling": "strong (the matrix is fixed; each GPU owns cells_total / N cells)"}


if __name__ == "__main__":
    sys.exit(main())
