#!/usr/bin/env python
"""bench.py — atom-updates/s of the Gibbs-sampler hot path on the BASELINE.json workload.

Workload (config.workload): synthetic dense 20000x5000, nPatterns=20 (BASELINE.json configs[2], the
configuration the metric is quoted on; recipe SURVEY.md 8(d)).  A "step" is one MCMC iteration of the
hot path exactly as GapsRunner.cpp:294-296 drives it: nA ~ Poisson(atomsA), nP ~ Poisson(atomsP),
A.update(nA) -> P.sync(A) -> P.update(nP) -> A.sync(P).

  value  : atom-updates/s over K timed steps with D/S/AP/factors resident in HBM, after a ramp that grows
           the chain from zero atoms (untimed) and W warm-up steps.  Host generator + device evaluator:
           every batch's proposals go H2D as kernel parameters and its outcomes come back D2H, inside the
           timed region, because that is what the path is.
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
RAMP_ITERS = 50          # untimed iterations growing the chain from zero atoms before warm-up
E2E_ITERS = 30           # iterations per phase of the end-to-end / reference gaps::run call


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


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons, sampled every 100 ms from before the ramp; the JSON line reports the
    samples that fall inside the timed region (plus the nearest one on either side when the region is short)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.t_begin = self.t_end = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def begin(self):
        self.t_begin = time.perf_counter()

    def end(self):
        self.t_end = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)   # let the sample that covers the end of the region arrive
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
                "reasons": sorted(reasons), "samples": len(sm), "samples_inside_timed_region": len(inside)}


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
              "device_mode": "resident grid per update(); each proposal streamed to its cluster through pinned host memory as it is generated; per-row commit versions instead of grid barriers"
              if os.environ.get("COGAPS_PERSISTENT", "1") != "0" else "one eval-kernel launch per batch",
              "l2": ("sparse model: per-proposal traffic is gathers of factor rows (k floats) at the row's non-zeros; "
                     "the CSR rows + both factors exceed L2 at 50000x30000, no flush needed") if args.sparse
              else "inputs larger than L2: 1.6 GB of resident D/AP streamed ~40 GB per step, no flush needed",
              "parallelism": "replicas x%d (one independent chain per GPU, no data-path collective)" % world
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

    data = make_data(args.rows, args.cols, args.patterns, DATA_SEED + rank, args.zeros if args.sparse else 0.0)
    clocks = ClockSampler(local_rank)
    t_setup = time.time()
    chain = Chain(data, args.patterns, CHAIN_SEED + rank, sparse=args.sparse)
    chain.ramp(args.ramp)
    for _ in range(args.warmup):
        chain.step()
    setup_s = time.time() - t_setup

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region: exactly K steps ----
    chain.A.resetCounters()
    chain.P.resetCounters()
    launches0 = cg.lib().cgb_kernel_launch_count()
    barrier()
    clocks.begin()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    updates = 0
    for _ in range(args.steps):
        updates += chain.step()
    torch.cuda.synchronize()      # the last A.sync(P) is still in flight on the samplers' streams
    ev1.record()
    barrier()
    elapsed_host = time.perf_counter() - t0
    # the region on the device clock: two CUDA events around the K steps (a step is a host-driven pipeline of
    # resident kernels on the samplers' own streams, every update() ends with a stream synchronise, so the
    # second event is reached when the last step is complete)
    elapsed = ev0.elapsed_time(ev1) * 1e-3
    clocks.end()
    clock_info = clocks.stop()
    launches = cg.lib().cgb_kernel_launch_count() - launches0
    cA, cP = chain.A.counters(), chain.P.counters()
    queued = cA.nProposalsQueued + cP.nProposalsQueued
    batches = cA.nBatches + cP.nBatches
    # every proposal: one 64-byte task record per CTA of its cluster written to pinned host memory and pulled by
    # the device (two-row moves / exchanges send two), one 16-byte outcome record written back
    segA, segP = chain.A.reductionOrder()[2], chain.P.reductionOrder()[2]
    h2d_step = (cA.nProposalsQueued * 64.0 * segA + cP.nProposalsQueued * 64.0 * segP) / args.steps
    d2h_step = queued * 16.0 / args.steps

    if world > 1:
        t = torch.tensor([elapsed], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_max = float(t.item())
        u = torch.tensor([float(updates), float(launches)], dtype=torch.float64, device="cuda")
        dist.all_reduce(u, op=dist.ReduceOp.SUM)
        total_updates, total_launches = float(u[0].item()), int(u[1].item())
    else:
        elapsed_max, total_updates, total_launches = elapsed, float(updates), int(launches)
    value = total_updates / elapsed_max

    # ---- roofline: the eval kernel of the timed region itself.  The resident kernel is launched once per
    # update() (2 per step); its duration comes from CUDA events on its stream, its bytes are the algorithmic
    # bytes of the proposals it evaluated.  The duration includes the time the grid waits for the
    # host generator: that is the launch as it runs in the product.
    roofline = None
    if rank == 0:
        peak, peak_src = load_peaks()
        resident_bytes = cA.algorithmicBytes + cP.algorithmicBytes
        resident_time = cA.secondsKernel + cP.secondsKernel
        n_launch = 2 * args.steps
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
        if not args.sparse:
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
        # chi-sq wall time (second half of BASELINE.json's metric)
        tcs = time.perf_counter()
        for _ in range(5):
            chain.P.chiSq()
        chisq_ms = (time.perf_counter() - tcs) / 5 * 1e3
    atomsA, atomsP = chain.A.nAtoms(), chain.P.nAtoms()
    del chain

    # ---- several chains sharing the device (the generator of ONE chain cannot keep a B200 busy) ----
    multi = None
    if rank == 0 and args.chains > 1:
        check(cg.lib().cgb_set_resident_share(args.chains))
        chains = [Chain(data, args.patterns, CHAIN_SEED + 100 + i, sparse=args.sparse) for i in range(args.chains)]
        gate = threading.Barrier(args.chains + 1)
        done_updates = [0] * args.chains

        def drive(i):
            chains[i].ramp(args.ramp)
            for _ in range(args.warmup):
                chains[i].step()
            gate.wait()
            n = 0
            for _ in range(args.steps):
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
                 "ms_per_step": tm / args.steps * 1e3, "per_chain": [u / tm for u in done_updates],
                 "how": "independent chains on one GPU, one host thread each, cgb_set_resident_share(chains); same workload, "
                        "ramp and step count as the single-chain line"}
        del chains
        check(cg.lib().cgb_set_resident_share(1))

    # ---- end to end: cgb_run (= gaps::run) on host buffers ----
    e2e = None
    if rank == 0 and not args.no_e2e:
        # three identical calls (same seed, same chain), the median by wall clock: the call is about a second and its
        # allocation / teardown share varies from box to box and run to run
        calls = []
        for _ in range(3):
            t0 = time.perf_counter()
            res = cg.gaps_run(data, seed=CHAIN_SEED, nPatterns=args.patterns, nIterations=args.e2e_iters,
                              outputFrequency=0, maxThreads=1, useSparseOptimization=1 if args.sparse else 0)
            calls.append((time.perf_counter() - t0, float(res.totalRunningTime)))
        wall, loop_s = sorted(calls)[1]
        n_it = 2 * args.e2e_iters
        upload = 2.0 * data.nbytes
        results = 4.0 * 2 * (args.rows + args.cols) * args.patterns
        e2e = {"value": res.totalUpdates / wall, "unit": "atom-updates/s",
               "h2d_bytes_per_step": (upload + h2d_step * args.steps / max(updates, 1) * res.totalUpdates) / n_it,
               "d2h_bytes_per_step": (results + 16.0 * res.totalUpdates) / n_it,
               "call": "cgb_run (gaps::run): host fp32 matrix in, Amean/Asd/Pmean/Psd out",
               "iterations_per_phase": args.e2e_iters, "atom_updates": int(res.totalUpdates), "wall_s": wall,
               "wall_s_of_each_call": [c[0] for c in calls], "how": "median of three identical calls",
               # the part of the call the reference arm's value covers (its sampler loop, GapsRunner.cpp:450,473)
               "sampler_loop_s": loop_s,
               "sampler_loop_value": res.totalUpdates / max(loop_s, 1e-9)}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, wall, upd, threads, sample = run_reference(args, data)
        cpu_baseline = {"value": v, "unit": "atom-updates/s", "cores": threads, "kind": "reference", "sample": sample}

    if rank == 0:
        line = {"metric": "atom_updates_per_s", "value": value, "unit": "atom-updates/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": elapsed_max / args.steps * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic", "config": config, "clocks": clock_info, "e2e": e2e,
                "gpu_launches": total_launches, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "chisq_ms": chisq_ms, "atoms": {"A": int(atomsA), "P": int(atomsP)},
                "batches_per_step": batches / args.steps, "proposals_per_batch": queued / max(batches, 1),
                "host_generate_s_per_step": (cA.secondsHostGenerate + cP.secondsHostGenerate) / args.steps,
                "device_wait_s_per_step": (cA.secondsDeviceWait + cP.secondsDeviceWait) / args.steps,
                "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step, "setup_s": setup_s,
                "timer": "CUDA events around the K timed steps (device clock), max over ranks; host perf_counter over the "
                         "same region: %.3f ms per step" % (elapsed_host / args.steps * 1e3)}
        if multi is not None:
            line["multi_chain"] = multi
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
