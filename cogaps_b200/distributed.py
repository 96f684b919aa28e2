"""Distributed CoGAPS (single-cell / genome-wide) with one process per GPU.

Reference: R/DistributedCogaps.R:48-119 (driver), :129-217 (consensus by pattern matching), :226-278
(stitchTogether), R/SubsetData.R:63-116 (createSets).  The reference fans the subsets out over BiocParallel worker
processes and passes R objects back; there is no per-iteration communication.  Here rank r of a
torch.distributed job owns the subsets r, r+W, ... and runs them on its own B200; the only exchanges are

  1. all-gather of each set's unmatched pattern matrix (L x k) before the consensus,
  2. (the consensus is computed redundantly and deterministically on every rank — no broadcast needed),
  3. all-gather of the per-shard factor rows (mean and sd) after the fixed-matrix pass — the `rbind` of
     stitchTogether — plus a sum of the per-set meanChiSq.

With NCCL the gathered tensors live on the device; with gloo (CPU tests) on the host.  Without an initialised
process group everything runs in this process, one subset after another.

Differences from the reference, on purpose: R's `sample()` stream is not reproducible outside R, so subsets are drawn
with numpy's PCG64 seeded by `params.seed` (same partition rule: floor(total/nSets) per set, remainder to the last,
each sorted).  The reference forces its *sequential* sampler inside distributed runs
(`allParams$asynchronousUpdates <- FALSE`, R/DistributedCogaps.R:28-29); pass `sequentialSampler=True` for exactly
that chain on the device (one proposal per host<->device round trip), the default keeps the asynchronous sampler,
which samples the same posterior and is what a GPU is for.  `agnes(..., "complete")` + `cutree` is scipy's complete
linkage on the same 1 - correlation dissimilarity.
"""
import copy
import os

import numpy as np


def createSets(total, nSets, seed, explicitSets=None, samplingAnnotation=None, samplingWeight=None, names=None):
    """createSets (R/SubsetData.R:85-116): list of 1-based index arrays, each sorted.

    explicitSets       sampleWithExplictSets (:8-30): index lists as given, or lists of names looked up in `names`
                       (geneNames for genome-wide, sampleNames for single-cell) with `which(allNames %in% set)`
    samplingAnnotation sampleWithAnnotationWeights (:39-58): every set draws floor(total/nSets) group labels with
                       probability proportional to samplingWeight (a {group: weight} mapping), then that many members
                       of each group WITH replacement — a set may hold an index more than once, as in the reference
    otherwise          sampleUniformly (:67-79): a random partition, floor(total/nSets) per set, the rest to the last
    """
    if explicitSets is not None:
        if len(explicitSets) != nSets:
            raise ValueError("nSets does not match number of explicit sets given")
        if all(len(s) and all(isinstance(x, str) for x in s) for s in explicitSets):
            if names is None:
                raise ValueError("named explicitSets need geneNames / sampleNames")
            allNames = np.asarray(list(names), dtype=object)
            sets = []
            for s in explicitSets:
                wanted = set(s)
                if not wanted.issubset(set(allNames.tolist())):
                    raise ValueError("some named genes in explicitSets not found")
                sets.append(np.nonzero(np.array([n in wanted for n in allNames]))[0].astype(np.int64) + 1)
            return sets
        return [np.asarray(s, dtype=np.int64) for s in explicitSets]   # returned as given (:13)
    rng = np.random.default_rng(int(seed))
    setSize = total // nSets
    if samplingAnnotation is not None:
        annotation = np.asarray(list(samplingAnnotation), dtype=object)
        if annotation.size != total:
            raise ValueError("samplingAnnotation must label every row (column) being partitioned")
        groups = sorted(set(annotation.tolist()))
        weight = np.array([float(dict(samplingWeight)[g]) for g in groups], dtype=np.float64)   # sorted by name (:42-45)
        if (weight < 0).any() or weight.sum() <= 0:
            raise ValueError("samplingWeight must be non-negative and not all zero")
        members = {g: np.nonzero(annotation == g)[0].astype(np.int64) + 1 for g in groups}
        sets = []
        for _ in range(nSets):
            counts = rng.multinomial(setSize, weight / weight.sum())       # sample(groups, setSize, TRUE, prob=weight)
            picks = [rng.choice(members[g], size=int(c), replace=True) for g, c in zip(groups, counts)]
            sets.append(np.sort(np.concatenate(picks)) if picks else np.zeros(0, np.int64))
        return sets
    remaining = np.arange(1, total + 1, dtype=np.int64)
    sets = []
    for _ in range(nSets - 1):
        pick = rng.choice(remaining.size, size=setSize, replace=False)
        sets.append(np.sort(remaining[pick]))
        remaining = np.delete(remaining, pick)
    sets.append(np.sort(remaining))
    return sets


def _cor_columns(m):
    m = np.asarray(m, dtype=np.float64)
    c = m - m.mean(axis=0, keepdims=True)
    n = np.sqrt((c * c).sum(axis=0))
    n[n == 0] = np.nan
    return (c.T @ c) / np.outer(n, n)


def corcut(allPatterns, cut, minNS):
    """R/DistributedCogaps.R:195-217"""
    from scipy.cluster.hierarchy import fcluster, linkage
    from scipy.spatial.distance import squareform
    dist = 1.0 - _cor_columns(allPatterns)
    if np.isnan(dist).any():
        raise ValueError("NA values in correlation of patterns")
    if allPatterns.shape[1] == 1:
        ids = np.array([1])
    else:
        np.fill_diagonal(dist, 0.0)
        z = linkage(squareform(np.maximum((dist + dist.T) / 2.0, 0.0), checks=False), method="complete")
        ids = fcluster(z, t=cut, criterion="maxclust")
    clusters = []
    seen = []
    for i in ids:                      # unique() keeps first-appearance order
        if i not in seen:
            seen.append(i)
    for i in seen:
        cols = np.where(ids == i)[0]
        if cols.size >= minNS:
            clusters.append(allPatterns[:, cols])
    return clusters


def corrToMeanPattern(cluster):
    """R/DistributedCogaps.R:183-187"""
    meanPat = cluster.mean(axis=1)
    out = []
    for j in range(cluster.shape[1]):
        out.append(round(float(np.corrcoef(cluster[:, j], meanPat)[0, 1]), 3))
    return np.array(out)


def patternMatch(allPatterns, cut, minNS, maxNS):
    """R/DistributedCogaps.R:143-177"""
    clusters = corcut(allPatterns, cut, minNS)
    guard = 0
    while True:
        big = [i for i, c in enumerate(clusters) if c.shape[1] > maxNS]
        if not big or guard > 1000:
            break
        guard += 1
        split = corcut(clusters[big[0]], 2, minNS)
        if not split:
            clusters.pop(big[0])
            continue
        clusters[big[0]] = split[0]
        if len(split) > 1:
            clusters.append(split[1])
    if not clusters:
        raise ValueError("pattern matching left no cluster (raise nSets or lower minNS)")
    mean = np.zeros((allPatterns.shape[0], len(clusters)))
    for j, c in enumerate(clusters):
        w = corrToMeanPattern(c) ** 3
        mean[:, j] = (c * w).sum(axis=1) / w.sum()
    consensus = mean / mean.max(axis=0, keepdims=True)
    return consensus.astype(np.float32), clusters


def findConsensusMatrix(unmatchedPatterns, params):
    """R/DistributedCogaps.R:129-135"""
    allPatterns = np.concatenate([np.asarray(u, dtype=np.float64) for u in unmatchedPatterns], axis=1)
    return patternMatch(allPatterns, params.cut, params.minNS, params.maxNS)


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except ImportError:
        pass
    return None


def _all_gather_rows(local, counts, device):
    """All-gather row blocks of unequal height: `local` is a list of (setIndex, 2-D float32 array) owned by
    this rank; `counts[i]` the row count of set i.  Returns the list of blocks for every set, in set order."""
    dist = _dist()
    if dist is None:
        out = [None] * len(counts)
        for i, a in local:
            out[i] = np.asarray(a, dtype=np.float32)
        return out
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    ncol = int(local[0][1].shape[1]) if local else 0
    t = torch.tensor([ncol], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ncol = int(t.item())
    perRank = (len(counts) + world - 1) // world
    maxRows = max(counts)
    buf = torch.zeros((perRank, maxRows, ncol), dtype=torch.float32, device=device)
    for slot, (i, a) in enumerate(local):
        buf[slot, :a.shape[0], :] = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf)          # NCCL over NVLink when the tensors are on the device
    out = [None] * len(counts)
    for r in range(world):
        g = gathered[r].cpu().numpy()
        for slot, i in enumerate(range(r, len(counts), world)):
            out[i] = g[slot, :counts[i], :].copy()
    return out


def stitchTogether(blocksMean, blocksSd, sets, total):
    """R/DistributedCogaps.R:226-278: rbind the per-set rows and undo the random partition when every row
    was used exactly once."""
    mean = np.concatenate(blocksMean, axis=0)
    sd = np.concatenate(blocksSd, axis=0)
    setIndices = np.concatenate(sets)
    if mean.shape[0] == total and np.array_equal(np.sort(setIndices), np.arange(1, total + 1)):
        reorder = np.argsort(setIndices, kind="stable")
        mean, sd = mean[reorder], sd[reorder]
    return mean, sd


def _default_runner(data, params, uncertainty, subset, subsetDim, runKw):
    """callInternalCoGAPS (R/DistributedCogaps.R:12-35): one ordinary CoGAPS run on a subset."""
    from .api import CoGAPS
    p = copy.copy(params)
    p.distributed = None
    p.subsetIndices = np.asarray(subset, dtype=np.uint32)
    p.subsetDim = subsetDim
    return CoGAPS(data, p, **runKw)


def distributedCogaps(data, params, uncertainty=None, nThreads=1, messages=False, outputFrequency=1000,
                      transposeData=False, runner=None, device=None, sequentialSampler=False, concurrentSets=1):
    """distributedCogaps (R/DistributedCogaps.R:48-119).  Returns a CogapsResult-like object.

    concurrentSets > 1: this rank runs that many of its subsets at once, one host thread each, the resident grids
    sharing the GPU (cgb_set_resident_share) — the reference's BiocParallel workers (R/DistributedCogaps.R:60-68)
    with threads for processes.  One chain's sequential proposal generator cannot keep a B200 busy; several can."""
    from .api import CogapsResult
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    if device is None:
        device = "cpu"
        if dist is not None and dist.get_backend() == "nccl":
            import torch
            device = torch.device("cuda", torch.cuda.current_device())
    runner = runner or _default_runner
    genomeWide = params.distributed == "genome-wide"
    if isinstance(data, (str, os.PathLike)):             # a data file: every subset run reads its own rows (columns)
        from .api import getFileInfo
        nrow, ncol = getFileInfo(data)["dimensions"]
    else:
        nrow, ncol = data.shape
    subsetRows = bool(transposeData) != genomeWide       # createSets, R/SubsetData.R:87-88
    total = nrow if subsetRows else ncol
    sets = createSets(total, params.nSets, params.seed, params.explicitSets, params.samplingAnnotation, params.samplingWeight,
                      names=params.geneNames if genomeWide else params.sampleNames)
    if min(len(s) for s in sets) < params.nPatterns:
        raise ValueError("data subset dimension less than nPatterns")
    subsetDim = 1 if genomeWide else 2
    mine = list(range(rank, len(sets), world))
    runKw = dict(nThreads=nThreads, messages=bool(messages) and rank == 0, outputFrequency=outputFrequency,
                 uncertainty=uncertainty, transposeData=transposeData, asynchronousUpdates=not sequentialSampler)
    counts = [len(s) for s in sets]
    nOther = (ncol if subsetRows else nrow)              # length of the un-partitioned dimension

    def run_all(p):
        """one run per subset this rank owns, concurrentSets at a time"""
        if concurrentSets <= 1 or len(mine) <= 1:
            return [(i, runner(data, p, uncertainty, sets[i], subsetDim, dict(runKw, workerID=i + 1))) for i in mine]
        from concurrent.futures import ThreadPoolExecutor
        from ._lib import lib, check
        check(lib().cgb_set_resident_share(min(concurrentSets, len(mine))))
        try:
            with ThreadPoolExecutor(max_workers=concurrentSets) as pool:
                futs = [(i, pool.submit(runner, data, p, uncertainty, sets[i], subsetDim, dict(runKw, workerID=i + 1))) for i in mine]
                return [(i, f.result()) for i, f in futs]
        finally:
            check(lib().cgb_set_resident_share(1))

    # ---- pass 1: ordinary runs on each subset, then match patterns across subsets ----
    firstPass = None
    if params.fixedPatterns is None:
        firstPass = run_all(params)
        unmatchedLocal = [(i, (r.sampleFactors if genomeWide else r.featureLoadings)) for i, r in firstPass]
        unmatched = _all_gather_rows(unmatchedLocal, [nOther] * len(sets), device)
        consensus, clusters = findConsensusMatrix(unmatched, params)
    else:
        consensus, clusters, unmatched = np.asarray(params.fixedPatterns, dtype=np.float32), None, None

    # ---- pass 2: the matched matrix is fixed, only the partitioned factor is sampled ----
    p2 = copy.copy(params)
    p2.nPatterns = int(consensus.shape[1])
    p2.cut = min(p2.cut, p2.nPatterns)
    p2.fixedPatterns = consensus
    p2.whichMatrixFixed = "P" if genomeWide else "A"
    final = run_all(p2)

    # ---- stitch: all-gather of the per-shard rows (mean and sd), sum of meanChiSq ----
    meanLocal = [(i, (r.featureLoadings if genomeWide else r.sampleFactors)) for i, r in final]
    sdLocal = [(i, (r.loadingStdDev if genomeWide else r.factorStdDev)) for i, r in final]
    blocksMean = _all_gather_rows(meanLocal, counts, device)
    blocksSd = _all_gather_rows(sdLocal, counts, device)
    mean, sd = stitchTogether(blocksMean, blocksSd, sets, total)
    chisq = float(sum(r.metadata["meanChiSq"] for _, r in final))
    updates = float(sum(r.metadata["totalUpdates"] for _, r in final))
    if dist is not None:
        import torch
        t = torch.tensor([chisq, updates], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        chisq, updates = float(t[0].item()), float(t[1].item())

    class _Res(object):
        pass
    res = _Res()
    other = consensus
    zeros = np.zeros_like(other)
    if genomeWide:
        res.Amean, res.Asd, res.Pmean, res.Psd = mean, sd, other, zeros
    else:
        res.Pmean, res.Psd, res.Amean, res.Asd = mean, sd, other, zeros
    res.meanChiSq = chisq
    res.chisqHistory = np.zeros(0, np.float32)
    res.atomHistoryA = np.zeros(0, np.uint32)
    res.atomHistoryP = np.zeros(0, np.uint32)
    res.totalUpdates = int(updates)
    res.totalRunningTime = 0.0
    res.averageQueueLengthA = res.averageQueueLengthP = 0.0
    res.seed = params.seed
    res.pumpMatrix = res.meanPatternAssignment = None
    res.snapshotsA = res.snapshotsP = np.zeros((0, 0, 0), np.float32)
    res.nSnapshotsEquilibration = 0
    out = CogapsResult(res, p2, {})
    out.metadata["firstPassResults"] = [r for _, r in firstPass] if firstPass else None
    out.metadata["unmatchedPatterns"] = unmatched
    out.metadata["clusteredPatterns"] = clusters
    out.metadata["CorrToMeanPattern"] = [corrToMeanPattern(c) for c in clusters] if clusters else None
    out.metadata["subsets"] = sets
    return out
