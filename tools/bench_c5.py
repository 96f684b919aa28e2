"""BASELINE.json configs[4]: distributed scCoGAPS row-shard of a (genes x cells) single-cell matrix across N GPUs with
the NCCL all-gather of per-shard P rows.  One rank per GPU under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_c5.py --genes 30000 --cells-total 200000 --patterns 50 --steps 5

Each rank owns cells_total / N cells: it builds its own block with the C4 recipe (bench.make_data, 95 % zeros, seed +
rank — "8 independent column blocks"), runs an independent sparse-model chain on it (what distributed CoGAPS
does per set: no per-iteration communication), and at the end all ranks all-gather their P rows (cells x patterns)
STRAIGHT FROM THE SAMPLERS' DEVICE MATRICES (cgb_sampler_device_matrix, zero copy into torch through
__cuda_array_interface__), timed with CUDA events.  Rank 0 prints one JSON line.

--dry-run (any machine, `--backend gloo`): no chain, random factor rows on the host — exercises the sharding arithmetic,
the gather and the JSON without a GPU (tests/test_host_logic.py).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class _DeviceArray(object):
    """A raw device pointer dressed as a CUDA array for torch.as_tensor (no copy)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def shard_sizes(total, world):
    """floor(total / nSets) per set, the remainder to the last one (R/SubsetData.R:63-75,90)"""
    base = total // world
    return [base] * (world - 1) + [total - base * (world - 1)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=30000)
    ap.add_argument("--cells-total", type=int, default=200000)
    ap.add_argument("--patterns", type=int, default=50)
    ap.add_argument("--zeros", type=float, default=0.95)
    ap.add_argument("--ramp", type=int, default=20)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--backend", default="nccl", choices=["nccl", "gloo"])
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    use_cuda = not args.dry_run
    if use_cuda:
        if not torch.cuda.is_available():
            raise SystemExit("bench_c5.py needs CUDA devices (or --dry-run --backend gloo)")
        torch.cuda.set_device(local)
    if world > 1:
        if use_cuda and args.backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(args.backend)
    device = torch.device("cuda", local) if use_cuda else torch.device("cpu")

    sizes = shard_sizes(args.cells_total, world)
    cells = sizes[rank]
    k = args.patterns
    ld = (cells + 31) // 32 * 32            # the samplers' pattern-major stride
    ldmax = (max(sizes) + 31) // 32 * 32

    def barrier():
        if use_cuda:
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if use_cuda:
            torch.cuda.synchronize()

    updates, elapsed, setup_s, atoms = 0, 0.0, 0.0, (0, 0)
    chain = None
    if use_cuda:
        import bench
        import cogaps_b200 as cg
        from cogaps_b200._lib import check
        check(cg.lib().cgb_set_device(local))
        t0 = time.time()
        data = bench.make_data(args.genes, cells, k, bench.DATA_SEED + rank, args.zeros)     # genes x this rank's cells
        chain = bench.Chain(data, k, bench.CHAIN_SEED + rank, sparse=True)
        chain.ramp(args.ramp)
        for _ in range(args.warmup):
            chain.step()
        setup_s = time.time() - t0
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            updates += chain.step()
        torch.cuda.synchronize()
        ev1.record()
        barrier()
        elapsed = ev0.elapsed_time(ev1) * 1e-3
        atoms = (int(chain.A.nAtoms()), int(chain.P.nAtoms()))
        ptr, stride = chain.P.deviceMatrix()                    # [k][stride] on the device, stride == ld
        assert stride == ld, (stride, ld)
        mine = torch.as_tensor(_DeviceArray(ptr, (k, ld)), device=device)
    else:
        rng = np.random.default_rng(rank)
        host = np.zeros((k, ld), np.float32)
        host[:, :cells] = rng.random((k, cells), dtype=np.float32)
        mine = torch.from_numpy(host)

    # ---- the exchange: all-gather of the per-shard P rows (pattern-major blocks padded to the widest shard) ----
    send = torch.zeros((k, ldmax), dtype=torch.float32, device=device)
    send[:, :ld] = mine
    gathered = [torch.empty_like(send) for _ in range(world)]
    gather_ms = 0.0
    if world > 1:
        for _ in range(2):                                      # warm-up: communicator setup, first-use allocations
            dist.all_gather(gathered, send)
        barrier()
        if use_cuda:
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            dist.all_gather(gathered, send)
            g1.record()
            torch.cuda.synchronize()
            gather_ms = g0.elapsed_time(g1)
        else:
            t0 = time.perf_counter()
            dist.all_gather(gathered, send)
            gather_ms = (time.perf_counter() - t0) * 1e3
    else:
        gathered = [send]
    # P (cells_total x k): shard r's cells are rows sum(sizes[:r]) ... (row-shard order; a random partition would be
    # undone here with the subset indices as stitchTogether does)
    full = torch.cat([gathered[r][:, :sizes[r]] for r in range(world)], dim=1).t().contiguous()
    checksum = float(full.double().sum().item())
    own = float(mine[:, :cells].double().sum().item())

    stats = torch.tensor([elapsed, gather_ms, float(updates), own], dtype=torch.float64, device=device)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    else:
        mx, sm = stats, stats
    if rank == 0:
        elapsed_max, gather_max = float(mx[0].item()), float(mx[1].item())
        total_updates, sum_of_shards = float(sm[2].item()), float(sm[3].item())
        line = {"metric": "atom_updates_per_s", "unit": "atom-updates/s", "n_gpus": world, "steps": args.steps,
                "value": (total_updates / elapsed_max) if elapsed_max > 0 else None,
                "ms_per_step": (elapsed_max / args.steps * 1e3) if elapsed_max > 0 else None,
                "scaling": "strong (the matrix is fixed, each GPU owns cells_total / N cells)", "dtype": "f32", "data": "synthetic",
                "config": {"workload": "synthetic sparse %dx%d (%.0f%% zeros) nPatterns=%d, row-sharded over %d GPUs"
                                       % (args.genes, args.cells_total, 100 * args.zeros, k, world),
                           "cells_per_gpu": sizes, "sampler": "asynchronous, sparse normal model", "ramp_iterations": args.ramp,
                           "dry_run": bool(args.dry_run)},
                "allgather": {"what": "per-shard P rows, pattern-major device blocks straight from the samplers (zero copy)",
                              "ms": gather_max, "bytes_per_rank": int(k * ldmax * 4), "bytes_total": int(k * ldmax * 4 * world),
                              "backend": args.backend if world > 1 else "none",
                              "gathered_shape": list(full.shape),
                              "checksum_matches_sum_of_shards": bool(abs(checksum - sum_of_shards) <= 1e-6 * max(1.0, abs(sum_of_shards)))},
                "atoms_rank0": {"A": atoms[0], "P": atoms[1]}, "setup_s_rank0": setup_s}
        print(json.dumps(line))
    del chain
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
