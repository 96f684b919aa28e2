"""bench.py's reference arm runs on the host cores only, so its whole code path can be exercised without a GPU:
`bench.py --impl reference` on a small shape must print ONE JSON line with the contract's keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    from oracle.harness import RefLib
    if not RefLib.available("scalar"):
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rows", "200", "--cols", "80",
                          "--patterns", "3", "--e2e-iters", "6", "--steps", "5", "--warmup", "3"],
                         check=True, capture_output=True, text=True, env=env, timeout=300).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference"
    assert line["metric"] == "atom_updates_per_s" and line["unit"] == "atom-updates/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["config"]["workload"] == "synthetic dense 200x80 nPatterns=3"
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert "sampler loop" in cb["sample"] and "whole gaps::run call" in cb["sample"]
    # the value is the reference's own sampler-loop clock, never more than the whole call
    assert line["ms_per_step"] > 0


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    done = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                          check=True, capture_output=True, text=True, env=env, timeout=120)
    assert done.stdout.strip() == ""


def test_c5_row_shard_tool_dry_run_world_2_gloo():
    """tools/bench_c5.py (BASELINE configs[4]: row-shard over N GPUs + all-gather of the per-shard P rows): the
    sharding arithmetic, the gather of ragged pattern-major blocks and the JSON line, without a GPU."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "tools", "bench_c5.py"), "--dry-run", "--backend", "gloo",
           "--cells-total", "777", "--genes", "40", "--patterns", "6"]
    done = subprocess.run(cmd, check=True, capture_output=True, text=True, timeout=300)
    lines = [ln for ln in done.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["config"]["cells_per_gpu"] == [388, 389]         # remainder to the last set
    g = line["allgather"]
    assert g["gathered_shape"] == [777, 6] and g["checksum_matches_sum_of_shards"] is True
    assert g["bytes_per_rank"] == 6 * 416 * 4 and g["backend"] == "gloo"                 # padded to the widest shard's stride
