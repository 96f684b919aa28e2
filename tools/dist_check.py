"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL): distributed single-cell CoGAPS on a synthetic
matrix; the per-shard factor rows are all-gathered on the device and the result must equal the single-process
run on rank 0's GPU bit for bit (same subsets, same seeds)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cogaps_b200 as cg
from cogaps_b200._lib import check
from cogaps_b200.distributed import distributedCogaps
from tests.cases import synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
check(cg.lib().cgb_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
data = synthetic("syn:300:%d:4:3" % (64 * world))
params = cg.CogapsParams(nPatterns=4, nIterations=80, seed=11, distributed="single-cell")
params.setParam("nSets", world)
res = distributedCogaps(data, params, outputFrequency=40)
dist.barrier()
ok = True
if rank == 0:
    dist.destroy_process_group()
    ref = distributedCogaps(data, params, outputFrequency=40)       # no process group: all subsets here
    ok = np.array_equal(res.sampleFactors, ref.sampleFactors) and np.array_equal(res.featureLoadings, ref.featureLoadings)
    print("dist_check world=%d: P %s A %s identical_to_single_process=%s" % (world, res.sampleFactors.shape,
                                                                          res.featureLoadings.shape, ok))
else:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
