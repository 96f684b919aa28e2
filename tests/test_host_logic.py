"""CPU-side tests (no GPU): the C ABI loads and exports what include/cogaps_b200.h declares, the host halves of
the path (RNG streams, lookup tables, parameter validation) agree with the oracle, compute entry points fail
loudly without a device, and the multi-process distributed driver works over gloo with world_size 2."""
import ctypes as C
import os
import re
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    import cogaps_b200 as cg
    from cogaps_b200._lib import EXPORTS
    header = open(os.path.join(ROOT, "include", "cogaps_b200.h")).read()
    declared = set(re.findall(r"\b(cgb_[a-z0-9_]+)\s*\(", header))
    assert len(declared) > 50
    lib = cg.lib()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert declared == set(EXPORTS), declared.symmetric_difference(set(EXPORTS))
    assert b"sm_100a" in lib.cgb_build_report()


def test_struct_layouts_match_the_header():
    """cgb_params_default fills the ctypes mirror exactly like CgbParams.defaults() (layout + defaults)."""
    import cogaps_b200 as cg
    from cogaps_b200._abi import CgbParams
    p = CgbParams()
    cg.lib().cgb_params_default(C.byref(p))
    d = CgbParams.defaults()
    assert p.struct_size == C.sizeof(CgbParams)
    for name, _ in CgbParams._fields_:
        if name in ("subsetIndices", "fixedPatterns"):
            continue
        assert getattr(p, name) == getattr(d, name), name
    # reference defaults, GapsParameters.h:79-114
    assert (p.nPatterns, p.nIterations, p.outputFrequency, p.asynchronousUpdates) == (3, 1000, 500, 1)
    assert p.alphaA == pytest.approx(0.01) and p.maxGibbsMassP == 100.0 and p.whichMatrixFixed == ord("N")


@pytest.mark.skipif(has_gpu(), reason="checks the no-device error path")
def test_compute_calls_fail_loudly_without_a_device():
    import cogaps_b200 as cg
    with pytest.raises(cg.CogapsError) as e:
        cg.gaps_run(np.ones((6, 5), np.float32), nPatterns=2, nIterations=2)
    assert e.value.code == -2          # CGB_ENODEVICE: no silent CPU fallback exists
    with pytest.raises(cg.CogapsError):
        cg.GapsStatistics(4, 3, 2)


def test_lookup_tables_match_oracle(oracle):
    """Random.cpp:269-295 — the library's built-in tables are the oracle's, bit for bit."""
    import cogaps_b200 as cg
    rs = cg.GapsRandomState(7)
    for a, b in zip(rs.tables(), oracle.tables()):
        assert np.array_equal(bits(a), bits(b))


@pytest.mark.parametrize("seed", [1, 42, 969, 4294967295])
def test_host_rng_streams_match_oracle(oracle, seed):
    """math/Random.cpp:32-200 through cgb_rng_*: seeder, PCG32, inclusive ranges, Poisson, truncated draws."""
    import cogaps_b200 as cg
    n = 300
    rs = cg.GapsRandomState(seed)
    assert [rs.nextSeed() for _ in range(5)] == oracle.rng_stream(seed, 0, 5).tolist()

    def stream(fn):
        r = cg.GapsRng(cg.GapsRandomState(seed))
        return [fn(r) for _ in range(n)]
    assert stream(lambda r: r.uniform32()) == oracle.rng_stream(seed, 1, n).tolist()
    assert stream(lambda r: r.uniform32(0, 9)) == oracle.rng_stream(seed, 2, n, a=0, b=9).tolist()
    assert stream(lambda r: r.uniform32(0, 4000000000)) == oracle.rng_stream(seed, 2, n, a=0, b=4000000000).tolist()
    big = 18446744073709551600
    assert stream(lambda r: r.uniform64(1, big)) == oracle.rng_stream(seed, 3, n, a=1, b=big).tolist()
    u = np.array(stream(lambda r: r.uniform()), np.float32)
    assert np.array_equal(bits(u), oracle.rng_stream(seed, 4, n).astype(np.uint32))
    for lam in (0.5, 4.9, 10.0, 3400.0, 215000.0):
        got = stream(lambda r: r.poisson(lam))
        assert got == oracle.rng_stream(seed, 5, n, lam=lam).astype(np.int64).tolist()
    for f in ((0.0, 5.0, 1.2, 0.7), (0.0, 50.0, -3.0, 0.5), (-1.5, 2.5, 0.2, 3.0), (0.0, 50.0, 80.0, 1.0)):
        got = stream(lambda r: r.truncNormal(*f))
        want = oracle.rng_stream(seed, 7, n, f=f)
        for g, w in zip(got, want):
            if w == 0xFFFFFFFFFFFFFFFF:
                assert g is None
            else:
                assert np.float32(g).view(np.uint32) == np.uint32(w)
    for f in ((3.0, 1.7), (0.02, 0.4)):
        got = np.array(stream(lambda r: r.truncGammaUpper(*f)), np.float32)
        assert np.array_equal(bits(got), oracle.rng_stream(seed, 8, n, f=f + (0, 0)).astype(np.uint32))
    # exponential uses the portable log: within 1 ulp of the libm-based reference stream
    got = np.array(stream(lambda r: r.exponential(0.37)), np.float32)
    want = oracle.rng_stream(seed, 6, n, f=(0.37, 0, 0, 0)).astype(np.uint32).view(np.float32)
    assert np.all(np.abs(got - want) <= np.abs(want) * 2.0 ** -22)


def test_host_portable_log_is_the_oracle_log(oracle):
    import cogaps_b200 as cg
    rng = np.random.default_rng(3)
    for x in np.concatenate([rng.random(3000), [0.0, 1.0, 2.0 ** -149, 0.5]]).astype(np.float32):
        a, b = cg.lib().cgb_debug_host_logf(float(x)), oracle.portable_logf(x)
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32)


def test_generator_division_by_fixed_divisors_is_exact():
    """The host generator divides positions by the bin length and bins by the pattern count with a multiply-high
    (atomic_domain.h FastDivU64); bin/row/col of a position decide which matrix element an atom belongs to
    (ProposalQueue.cpp:174-175), so the quotient must equal the hardware divide for every input."""
    import cogaps_b200 as cg
    f = cg.lib().cgb_debug_fastdiv
    rng = np.random.default_rng(11)
    M = (1 << 64) - 1
    divisors = [1, 2, 3, 7, 20, 50, 1363 * 7, 20000 * 20, M // (20000 * 20), M // (25 * 3), M // (200000 * 50), M, M - 1, (1 << 63), (1 << 32) + 1]
    for d in divisors:
        xs = [0, 1, d - 1, d, d + 1, 2 * d - 1 if 2 * d - 1 <= M else M, M, M - 1, M // 2] + [int(v) for v in rng.integers(0, M, 200, dtype=np.uint64)]
        xs += [min(M, d * int(q) + int(r)) for q, r in zip(rng.integers(0, max(M // d, 1), 50, dtype=np.uint64), rng.integers(0, 3, 50))]
        for x in xs:
            x = max(0, min(M, x))
            assert f(d, x) == x // d, (d, x)


def test_reduction_order_is_a_function_of_row_length():
    from cogaps_b200.sampler import reduction_order_for_length
    t, v, nseg, seg = reduction_order_for_length(5000)
    assert t % 32 == 0 and v == 4 and seg % 4 == 0 and nseg * seg >= 5000
    for length in (1, 9, 1363, 5000, 20000, 30000):
        t, v, nseg, seg = reduction_order_for_length(length)
        assert nseg * seg >= length and (nseg - 1) * seg < length


def test_params_validation_mirrors_reference():
    """setValidity("CogapsParams"), R/class-CogapsParams.R:126-193"""
    import cogaps_b200 as cg
    p = cg.CogapsParams(nPatterns=3)
    assert (p.nIterations, p.alphaA, p.maxGibbsMassA, p.nSets, p.cut, p.minNS, p.maxNS) == (50000, 0.01, 100.0, 4, 3, 2, 6)
    for bad in (dict(nPatterns=0), dict(nPatterns=2.5), dict(nPatterns=3, nIterations=0), dict(nPatterns=3, alphaA=0),
                dict(nPatterns=3, seed=0), dict(nPatterns=3, whichMatrixFixed="Q"),
                dict(nPatterns=3, whichMatrixFixed="A"), dict(nPatterns=3, subsetDim=1)):
        with pytest.raises(ValueError):
            cg.CogapsParams(**bad)
    with pytest.raises(ValueError):
        cg.CogapsParams(nPatterns=3, nSets=5)         # "nSets must be set after CogapsParams are intialized"
    p.setParam("nSets", 6)
    assert (p.minNS, p.maxNS) == (3, 9)
    with pytest.raises(ValueError):
        cg.CogapsParams(nPatterns=3, distributed="single-cell", fixedPatterns=np.ones((4, 3)), whichMatrixFixed="P")


def test_subset_indices_are_bounds_checked_before_anything_else():
    """The reference indexes its matrix with whatever it is given (Matrix.cpp:30-69: undefined behaviour past the end);
    the library refuses — and does so before it looks for a device, so this runs anywhere."""
    import ctypes as C
    import cogaps_b200 as cg
    from cogaps_b200._lib import lib
    from cogaps_b200._runhelp import make_params, fptr
    data = np.ones((6, 4), np.float32)
    rs = cg.GapsRandomState(1)
    for transpose, subsetRows, bad, good in ((1, 0, [7], [6]), (0, 1, [7], [6]), (1, 1, [5], [4]), (0, 0, [5], [4]), (0, 0, [0], [1])):
        for idx, want in ((bad, -1), (good, None)):
            p = make_params(nPatterns=2, subsetIndices=idx, subsetGenes=subsetRows)
            h = C.c_void_p()
            rc = lib().cgb_sampler_create(fptr(data), 6, 4, 0, transpose, subsetRows, 0.01, 100.0, C.byref(p), rs._h, C.byref(h))
            if want is not None:
                assert rc == want and b"subset ind" in lib().cgb_last_error()
            elif rc == 0:
                lib().cgb_sampler_destroy(h)          # a GPU is present: the sampler was really built
            else:
                assert rc == -2                       # CGB_ENODEVICE: the indices passed, the device is what is missing


# ---------------------------------------------------------------------------------------------
# distributed driver
# ---------------------------------------------------------------------------------------------
def test_create_sets_partitions_like_reference():
    """R/SubsetData.R:63-75: floor(total/nSets) per set, remainder to the last, sorted, disjoint, complete."""
    from cogaps_b200.distributed import createSets
    sets = createSets(103, 4, seed=42)
    assert [len(s) for s in sets] == [25, 25, 25, 28]
    allidx = np.concatenate(sets)
    assert np.array_equal(np.sort(allidx), np.arange(1, 104))
    assert all(np.all(np.diff(s) > 0) for s in sets)
    again = createSets(103, 4, seed=42)
    assert all(np.array_equal(a, b) for a, b in zip(sets, again))
    with pytest.raises(ValueError):
        createSets(10, 3, 1, explicitSets=[[1, 2], [3]])


def test_create_sets_with_annotation_weights_and_named_sets():
    """sampleWithAnnotationWeights (R/SubsetData.R:39-58) and named explicit sets (:15-29)."""
    import cogaps_b200 as cg
    from cogaps_b200.distributed import createSets
    annotation = ["b"] * 40 + ["a"] * 50 + ["c"] * 10
    sets = createSets(100, 4, seed=3, samplingAnnotation=annotation, samplingWeight={"a": 3.0, "c": 1.0, "b": 0.0})
    assert [len(s) for s in sets] == [25, 25, 25, 25]                      # floor(total / nSets) draws each
    allidx = np.concatenate(sets)
    assert allidx.min() >= 41 and allidx.max() <= 100                       # weight 0: group "b" (1..40) never drawn
    assert all(np.all(np.diff(s) >= 0) for s in sets)                       # sorted, repeats allowed (replace=TRUE)
    share_a = np.mean((allidx >= 41) & (allidx <= 90))
    assert 0.55 < share_a < 0.95                                            # about 3:1 in favour of "a"
    assert all(np.array_equal(x, y) for x, y in zip(sets, createSets(100, 4, seed=3, samplingAnnotation=annotation,
                                                                      samplingWeight={"a": 3.0, "c": 1.0, "b": 0.0})))
    names = ["g%d" % i for i in range(1, 11)]
    sets = createSets(10, 2, seed=1, explicitSets=[["g3", "g1", "g7"], ["g10", "g2"]], names=names)
    assert [s.tolist() for s in sets] == [[1, 3, 7], [2, 10]]               # which(allNames %in% set)
    with pytest.raises(ValueError, match="not found"):
        createSets(10, 2, seed=1, explicitSets=[["g3", "nope"], ["g2"]], names=names)
    assert [s.tolist() for s in createSets(10, 2, seed=1, explicitSets=[[5, 2, 9], [1]])] == [[5, 2, 9], [1]]  # as given

    # validity rules of the S4 class (R/class-CogapsParams.R:161-189) and the dedicated setter
    p = cg.CogapsParams(nPatterns=3, distributed="genome-wide")
    with pytest.raises(ValueError, match="setAnnotationWeights"):
        p.setParam("samplingWeight", {"a": 1.0})
    p.setAnnotationWeights(["a", "b", "a"], {"a": 1.0, "b": 2.0})
    with pytest.raises(ValueError, match="mismatched size"):
        p.setAnnotationWeights(["a", "b", "a"], {"a": 1.0})
    p = cg.CogapsParams(nPatterns=3, distributed="single-cell")
    with pytest.raises(ValueError, match="length of explicitSets"):
        p.setParam("explicitSets", [[1, 2], [3, 4]])                        # nSets is 4
    assert p.explicitSets is None                                           # a rejected change leaves no trace
    p.setParam("nSets", 2)
    p.setParam("explicitSets", [[1, 2], [3, 4]])
    with pytest.raises(ValueError, match="numeric or character"):
        p.setParam("explicitSets", [[1, 2], ["x"]])
    with pytest.raises(ValueError, match="manual pattern matching"):
        cg.CogapsParams(nPatterns=3, distributed="single-cell", fixedPatterns=np.ones((4, 3)), whichMatrixFixed="A")


def _synthetic_patterns(nSets=4, k=3, length=60, seed=0):
    rng = np.random.default_rng(seed)
    base = rng.gamma(2.0, 1.0, (length, k))
    out = []
    for _ in range(nSets):
        perm = rng.permutation(k)
        out.append((base[:, perm] * rng.uniform(0.5, 2.0, k) + 0.05 * rng.random((length, k))).astype(np.float32))
    return base, out


def test_consensus_recovers_shared_patterns():
    """patternMatch, R/DistributedCogaps.R:143-177: permuted, rescaled, noisy copies cluster back together."""
    import cogaps_b200 as cg
    from cogaps_b200.distributed import findConsensusMatrix
    base, unmatched = _synthetic_patterns()
    params = cg.CogapsParams(nPatterns=3, distributed="single-cell")
    consensus, clusters = findConsensusMatrix(unmatched, params)
    assert consensus.shape == (60, 3) and len(clusters) == 3 and all(c.shape[1] == 4 for c in clusters)
    assert np.allclose(consensus.max(axis=0), 1.0)
    cor = np.abs(np.corrcoef(consensus.T, base.T)[:3, 3:])
    assert np.all(cor.max(axis=1) > 0.99) and sorted(cor.argmax(axis=1).tolist()) == [0, 1, 2]


def _naive_complete_linkage(dist, k):
    """agnes(diss, method="complete") + cutree(k), R/DistributedCogaps.R:206-207, restated the slow way: start from
    singletons, merge the two clusters whose FARTHEST members are closest until k clusters are left; ids numbered by the
    first member in column order (what cutree reports)."""
    clusters = [[i] for i in range(dist.shape[0])]
    while len(clusters) > k:
        best = None
        for a in range(len(clusters)):
            for b in range(a + 1, len(clusters)):
                d = max(dist[i, j] for i in clusters[a] for j in clusters[b])
                if best is None or d < best[0]:
                    best = (d, a, b)
        _, a, b = best
        clusters[a] = sorted(clusters[a] + clusters[b])
        del clusters[b]
    clusters.sort(key=lambda c: c[0])
    ids = np.zeros(dist.shape[0], dtype=int)
    for n, c in enumerate(clusters):
        ids[c] = n + 1
    return ids


def test_corcut_against_a_naive_complete_linkage():
    """corcut (scipy's linkage + fcluster standing in for R's agnes + cutree, R/DistributedCogaps.R:195-217) against the
    definition of complete linkage written out: same partition of the patterns, clusters in order of first appearance,
    clusters below minNS dropped — on 40 random pattern sets of 3 to 7 planted groups."""
    from cogaps_b200.distributed import corcut
    rng = np.random.default_rng(7)
    for trial in range(40):
        groups, per, length = int(rng.integers(3, 8)), int(rng.integers(2, 6)), 50
        base = rng.random((length, groups))
        cols = [base[:, g] * rng.uniform(0.5, 2.0) + 0.25 * rng.random(length) for g in range(groups) for _ in range(per)]
        allPatterns = np.stack([cols[i] for i in rng.permutation(len(cols))], axis=1)
        cut, minNS = groups, int(rng.integers(1, 4))
        got = corcut(allPatterns, cut, minNS)
        c = allPatterns - allPatterns.mean(axis=0, keepdims=True)
        n = np.sqrt((c * c).sum(axis=0))
        dist = 1.0 - (c.T @ c) / np.outer(n, n)
        ids = _naive_complete_linkage(dist, cut)
        want = [allPatterns[:, ids == i] for i in range(1, cut + 1) if (ids == i).sum() >= minNS]
        assert len(got) == len(want), trial
        for g, w in zip(got, want):
            assert g.shape == w.shape and np.array_equal(g, w), trial


def test_pattern_match_known_answer():
    """patternMatch on a case small enough to follow by hand (R/DistributedCogaps.R:143-177): two clusters of two patterns
    each; the consensus of a cluster is weighted.mean(row, round(cor(pattern, rowMeans), 3)^3) per row, scaled to a
    maximum of 1 — written out here with plain Python sums."""
    from cogaps_b200.distributed import patternMatch
    a1 = np.array([1.0, 2.0, 3.0, 4.0, 0.5, 0.2])
    a2 = np.array([1.2, 1.8, 3.3, 3.9, 0.4, 0.1])
    b1 = np.array([4.0, 0.5, 3.0, 0.1, 2.0, 5.0])
    b2 = np.array([3.6, 0.9, 2.8, 0.3, 2.4, 4.6])
    allPatterns = np.stack([a1, b1, a2, b2], axis=1)
    consensus, clusters = patternMatch(allPatterns, cut=2, minNS=2, maxNS=4)
    assert len(clusters) == 2
    assert np.array_equal(clusters[0], np.stack([a1, a2], axis=1)) and np.array_equal(clusters[1], np.stack([b1, b2], axis=1))

    def pearson(x, y):
        n = len(x)
        mx, my = sum(x) / n, sum(y) / n
        sxy = sum((p - mx) * (q - my) for p, q in zip(x, y))
        return sxy / (sum((p - mx) ** 2 for p in x) * sum((q - my) ** 2 for q in y)) ** 0.5

    for col, (u, v) in enumerate(((a1, a2), (b1, b2))):
        mean = [(p + q) / 2 for p, q in zip(u, v)]
        wu, wv = round(pearson(list(u), mean), 3) ** 3, round(pearson(list(v), mean), 3) ** 3
        want = [(wu * p + wv * q) / (wu + wv) for p, q in zip(u, v)]
        want = [x / max(want) for x in want]
        assert np.allclose(consensus[:, col], want, rtol=1e-6), col
    # a cluster beyond maxNS is cut in two (splitCluster); one that falls below minNS by the cut is dropped
    six = np.stack([a1, a2, a1 * 1.1 + 0.01, b1, b2, b1 * 0.9 + 0.02], axis=1)
    consensus2, clusters2 = patternMatch(six, cut=1, minNS=2, maxNS=3)
    assert sorted(c.shape[1] for c in clusters2) == [3, 3]
    assert consensus2.shape == (6, 2) and np.allclose(consensus2.max(axis=0), 1.0)


class _FakeResult(object):
    def __init__(self, A, P):
        self.featureLoadings, self.sampleFactors = A, P
        self.loadingStdDev, self.factorStdDev = 0.1 * A, 0.1 * P
        self.metadata = dict(meanChiSq=1.5, totalUpdates=10)


def _fake_runner(data, params, uncertainty, subset, subsetDim, runKw):
    """Stands in for a GPU run: the 'factorisation' of a subset is a deterministic function of its indices."""
    base, _ = _synthetic_patterns(k=params.nPatterns, length=data.shape[0])
    idx = np.asarray(subset, dtype=np.int64)
    if params.fixedPatterns is None:
        A = base[:, ::-1] if int(idx[0]) % 2 else base
        P = np.outer(idx, np.arange(1, params.nPatterns + 1))
    else:
        A = np.asarray(params.fixedPatterns)
        P = np.outer(idx, np.arange(1, A.shape[1] + 1))
    return _FakeResult(A.astype(np.float32), P.astype(np.float32))


def _run_distributed(world, rank, port, out):
    import torch.distributed as dist
    if world > 1:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import cogaps_b200 as cg
    from cogaps_b200.distributed import distributedCogaps
    data = np.ones((60, 103), np.float32)
    params = cg.CogapsParams(nPatterns=3, distributed="single-cell", seed=42)
    res = distributedCogaps(data, params, runner=_fake_runner)
    if rank == 0:
        np.savez(out, P=res.sampleFactors, Psd=res.factorStdDev, A=res.featureLoadings, chisq=res.metadata["meanChiSq"],
                 updates=res.metadata["totalUpdates"])
    if world > 1:
        dist.destroy_process_group()


def _spawn_entry(rank, world, port, out):
    _run_distributed(world, rank, port, out)


def test_distributed_single_cell_world_size_2_gloo(tmp_path):
    """subsets round-robin over ranks, all-gather of unequal row blocks, rows restored to data
    order (stitchTogether).  Two gloo processes must reproduce the single-process result exactly."""
    import torch.multiprocessing as mp
    single = str(tmp_path / "single.npz")
    _run_distributed(1, 0, 0, single)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    multi = str(tmp_path / "multi.npz")
    mp.spawn(_spawn_entry, args=(2, port, multi), nprocs=2, join=True)
    a, b = np.load(single), np.load(multi)
    for k in ("P", "Psd", "A", "chisq", "updates"):
        assert np.array_equal(a[k], b[k]), k
    # every sample's row is the one computed from its own (1-based) index: the partition was undone
    assert np.array_equal(a["P"][:, 0], np.arange(1, 104, dtype=np.float32))
    assert a["P"].shape == (103, 3) and a["A"].shape == (60, 3)
    assert float(a["chisq"]) == pytest.approx(4 * 1.5) and int(a["updates"]) == 40


def test_concurrent_sets_give_the_sequential_result():
    """distributedCogaps(concurrentSets=n): a rank's subsets run from n host threads at once (the reference's
    BiocParallel workers, R/DistributedCogaps.R:60-68); the stitched result does not depend on it."""
    import cogaps_b200 as cg
    from cogaps_b200.distributed import distributedCogaps
    data = np.ones((60, 103), np.float32)
    params = cg.CogapsParams(nPatterns=3, distributed="single-cell", seed=42)
    one = distributedCogaps(data, params, runner=_fake_runner)
    many = distributedCogaps(data, params, runner=_fake_runner, concurrentSets=3)
    assert np.array_equal(one.sampleFactors, many.sampleFactors)
    assert np.array_equal(one.featureLoadings, many.featureLoadings)
    assert one.metadata["meanChiSq"] == many.metadata["meanChiSq"]


def test_header_is_plain_c(tmp_path):
    """The boundary is a C ABI: include/cogaps_b200.h must compile as C99 with nothing but <stdint.h>, and a C
    program must link against the library by the declared names."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text(
        '#include "cogaps_b200.h"\n'
        "int main(void)\n"
        "{\n"
        "    cgb_params p;\n"
        "    cgb_params_default(&p);\n"
        "    return (p.struct_size == sizeof(cgb_params) && p.nPatterns == 3 && cgb_build_report() != 0) ? 0 : 1;\n"
        "}\n")
    exe = tmp_path / "abi"
    libdir = os.path.join(ROOT, "cogaps_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                           str(src), "-o", str(exe), "-L", libdir, "-lcogaps_b200", "-Wl,-rpath," + libdir])
    assert subprocess.call([str(exe)]) == 0


# ---------------------------------------------------------------------------------------------
# the product's proposal generator against the oracle's trace, no GPU needed
# ---------------------------------------------------------------------------------------------
REPLAY_CASES = {
    "modsim": ("modsim", dict(seed=42, nPatterns=3, nIterations=300, outputFrequency=0)),
    "gist": ("gist", dict(seed=42, nPatterns=7, nIterations=40, outputFrequency=0)),
    "gist_sparse": ("gist", dict(seed=9, nPatterns=4, nIterations=40, outputFrequency=0, useSparseOptimization=1)),
    "syn_wide": ("syn:40:700:4:9", dict(seed=9, nPatterns=4, nIterations=40, outputFrequency=0)),
    "gist_fixedP": ("gist", dict(seed=8, nPatterns=3, nIterations=40, outputFrequency=0, whichMatrixFixed="P")),
}


def _replay(data, trace, **kw):
    import ctypes as C
    from cogaps_b200._lib import lib
    from cogaps_b200._runhelp import make_params, fptr
    data = np.ascontiguousarray(data, dtype=np.float32)
    p = make_params(**kw)
    checked = C.c_uint64()
    rc = lib().cgb_debug_replay_generator(fptr(data), data.shape[0], data.shape[1], C.byref(p),
                                          trace.ctypes.data_as(C.c_void_p), trace.size, C.byref(checked))
    return rc, checked.value, lib().cgb_debug_replay_message().decode()


@pytest.mark.parametrize("name", sorted(REPLAY_CASES))
def test_generator_replays_the_oracle_trace(oracle, name):
    """ProposalQueue + AtomicDomain of the LIBRARY (what cgb_sampler_update drives) against the oracle's record of a
    whole run: every queued proposal — type, bins, positions, atom masses, PCG state, batch boundaries — must be the
    one the oracle (pinned to the reference's ProposalQueue.cpp / ConcurrentAtomicDomain.cpp) evaluated at that point,
    given the same outcomes.  This holds the sequential half of the hot path to the reference without a GPU."""
    from tests.cases import load_data
    dataset, kw = REPLAY_CASES[name]
    data = load_data(dataset)
    if "whichMatrixFixed" in kw:
        kw = dict(kw, fixedPatterns=np.random.default_rng(7).gamma(2.0, 0.5, (data.shape[1], kw["nPatterns"])).astype(np.float32))
    res = oracle.run(data, trace_capacity=400000, **kw)
    assert 1000 < res.trace_total <= 400000
    rc, checked, msg = _replay(data, res.trace, **kw)
    assert rc == 0, msg
    assert checked == res.trace_total


def test_generator_replay_notices_a_wrong_trace(oracle):
    from tests.cases import load_data
    data = load_data("modsim")
    kw = dict(seed=42, nPatterns=3, nIterations=60, outputFrequency=0)
    trace = oracle.run(data, trace_capacity=100000, **kw).trace
    for field, at in (("rngState", 50), ("r1", 333), ("accepted", 40), ("newMass1", 700)):
        bad = trace.copy()
        if field == "accepted":
            at = int(np.nonzero((bad["type"] == ord("B")) & (bad["accepted"] == 1))[0][5])
            bad[field][at] = 0                      # a birth that now fails: the atom count diverges from here on
        elif field == "newMass1":
            at = int(np.nonzero(bad["accepted"] == 1)[0][200])
            bad[field][at] += np.float32(0.25)      # seen when that atom is next picked
        else:
            bad[field][at] ^= 1
        rc, checked, msg = _replay(data, bad, **kw)
        assert rc != 0 and checked >= at and msg, (field, checked, msg)
        if field in ("rngState", "r1"):
            assert checked == at and ("rng state" in msg or "r1" in msg)
    rc, checked, msg = _replay(data, trace[:-7], **kw)
    assert rc != 0 and "more proposals than the trace" in msg
    # lambda (alpha * sqrt(k / nonZeroMean(D)), through the samplers' own running sum) is held too: a same-bin exchange
    # draws its new mass with scale 1 / lambda, so other data under the same trace shows up in an atom's mass
    rc, checked, msg = _replay(data * np.float32(1.5), trace, **kw)
    assert rc != 0 and "mass bits" in msg
    rc, checked, msg = _replay(data, trace, **dict(kw, seed=43))
    assert rc == 0          # not a typo: the seeder starts from seed|1 (Random.cpp:222-223), 42 and 43 are one stream
    rc, checked, msg = _replay(data, trace, **dict(kw, seed=44))
    assert rc != 0 and checked == 0


@pytest.mark.parametrize("shape", [(1, 1), (3, 5), (37, 16), (40, 17), (129, 31), (500, 203)])
def test_running_sum_behind_lambda_is_order_exact(shape):
    """gaps::nonZeroMean (MatrixMath.cpp:39-55) is one fp32 running sum: the ORDER of the additions is the result.  The
    samplers take it row by row, or — for the sampler whose rows run down the caller's columns — through 16-row strips
    gathered into contiguous memory first; both must be the plain sequential sum in that order, bit for bit."""
    import ctypes as C
    from cogaps_b200._lib import lib, check
    from cogaps_b200._runhelp import fptr
    rng = np.random.default_rng(shape[0] * 1000 + shape[1])
    data = (rng.gamma(2.0, 3.0, shape) * (rng.random(shape) < 0.7)).astype(np.float32)
    for by_columns, order in ((0, data.ravel()), (1, data.T.ravel())):
        want = np.add.accumulate(order, dtype=np.float32)[-1]        # accumulate is strictly sequential
        s, n = C.c_float(), C.c_uint32()
        check(lib().cgb_debug_running_sum(fptr(data), shape[0], shape[1], by_columns, C.byref(s), C.byref(n)))
        assert np.float32(s.value).view(np.uint32) == np.float32(want).view(np.uint32)
        assert n.value == int((data > 0).sum())


@pytest.mark.parametrize("family", ["gamma", "counts", "sparse_small", "wide_range", "rare_negative", "rare_huge", "tiny"])
def test_running_sum_fast_routes_are_order_exact(family):
    """The same, on matrices large enough for the fast routes (csrc/sampler.h): accumulateRun adds chunks as integer
    multiples of the sum's ulp while the sum stays in one binade (any tie, binade change, negative, huge or NaN element
    sends the chunk the plain way), and the column-walking sum has helper threads gather its strips ahead.  Every family
    below must give the bits of the plain sequential fp32 sum: float data, counts (ties at every turn), mostly zeros, twelve
    orders of magnitude, rare negatives, rare huge values, values so small the sum stays subnormal for a while."""
    import ctypes as C
    from cogaps_b200._lib import lib, check
    from cogaps_b200._runhelp import fptr
    rng = np.random.default_rng(len(family))
    shape = (4200, 4099)                                             # 17.2 M elements: above the pipelining threshold
    if family == "gamma":
        data = rng.gamma(2.0, 3.0, shape)
    elif family == "counts":
        data = rng.poisson(1.3, shape).astype(np.float64)
    elif family == "sparse_small":
        data = rng.random(shape) * 0.004 * (rng.random(shape) < 0.05)
    elif family == "wide_range":
        data = np.exp(rng.uniform(-14.0, 14.0, shape))
    elif family == "rare_negative":
        data = rng.gamma(2.0, 3.0, shape)
        data[rng.random(shape) < 1e-5] *= -1.0
    elif family == "rare_huge":
        data = rng.gamma(2.0, 3.0, shape)
        data[rng.random(shape) < 1e-6] = 3.0e12
    else:
        data = rng.random(shape) * 1e-44
    data = np.ascontiguousarray(data, dtype=np.float32)
    for by_columns, order in ((0, data.ravel()), (1, data.T.ravel())):
        want = np.add.accumulate(order, dtype=np.float32)[-1]
        s, n = C.c_float(), C.c_uint32()
        check(lib().cgb_debug_running_sum(fptr(data), shape[0], shape[1], by_columns, C.byref(s), C.byref(n)))
        assert np.float32(s.value).view(np.uint32) == np.float32(want).view(np.uint32), (family, by_columns, s.value, want)
        assert n.value == int((data > 0).sum())


def test_distributed_driver_with_named_explicit_sets():
    """explicitSets given by sample name travel through distributedCogaps (R/SubsetData.R:15-29): each subset run sees
    exactly the named columns, and the stitched rows come back in the order of the sets."""
    import cogaps_b200 as cg
    from cogaps_b200.distributed import distributedCogaps
    rng = np.random.default_rng(3)
    data = rng.gamma(2.0, 1.0, (30, 12)).astype(np.float32)
    names = ["cell%02d" % j for j in range(12)]
    params = cg.CogapsParams(nPatterns=3, distributed="single-cell", seed=42)
    params.setParam("sampleNames", names)
    params.setParam("nSets", 2)
    params.setParam("explicitSets", [names[0:12:2], names[1:12:2]])
    seen = []

    def runner(d, p, unc, subset, subsetDim, runKw):
        seen.append((np.asarray(subset).tolist(), subsetDim))
        return _fake_runner(d, p, unc, subset, subsetDim, runKw)

    res = distributedCogaps(data, params, runner=runner)
    assert seen[0] == ([1, 3, 5, 7, 9, 11], 2) and seen[1] == ([2, 4, 6, 8, 10, 12], 2)
    assert [s.tolist() for s in res.metadata["subsets"]] == [[1, 3, 5, 7, 9, 11], [2, 4, 6, 8, 10, 12]]
    assert res.sampleFactors.shape == (12, 3)


def test_cogaps_accepts_a_data_file_name(tmp_path):
    """CoGAPS(data = "file.mtx" / .csv / .tsv / .gct) — R/CoGAPS.R:90-156 with cogaps_from_file_cpp underneath: the checks
    that need no device, and the distributed driver's plumbing with a stand-in runner."""
    import cogaps_b200 as cg
    from cogaps_b200.distributed import distributedCogaps
    rng = np.random.default_rng(5)
    m = rng.gamma(2.0, 1.0, (30, 12)).astype(np.float32)
    path = tmp_path / "data.csv"
    cg.write_matrix_csv(path, m)
    assert cg.getFileInfo(path)["dimensions"] == (30, 12)
    with pytest.raises(ValueError, match="same data type"):
        cg.CoGAPS(path, nPatterns=3, uncertainty=np.ones_like(m), messages=False)
    with pytest.raises(ValueError, match="file path"):
        cg.CoGAPS(m, nPatterns=3, uncertainty=str(path), messages=False)
    with pytest.raises(cg.CogapsError):
        cg.CoGAPS(tmp_path / "data.txt", nPatterns=3, messages=False)                 # unsupported extension
    seen = []

    def runner(d, p, unc, subset, subsetDim, runKw):
        seen.append((str(d), len(subset), subsetDim))
        return _fake_runner(np.zeros((30, 12), np.float32), p, unc, subset, subsetDim, runKw)

    params = cg.CogapsParams(nPatterns=3, distributed="single-cell", seed=42)
    params.setParam("nSets", 3)
    res = distributedCogaps(str(path), params, runner=runner)
    assert seen[0] == (str(path), 4, 2) and len(seen) == 6                            # three subsets of four cells, two passes
    assert res.sampleFactors.shape == (12, 3)


@pytest.mark.parametrize("seed,nBins,nOps", [(1, 60, 20000), (2, 3, 5000), (3, 4000, 120000), (4, 1, 2000), (5, 400000, 200000)])
def test_atomic_domain_against_a_naive_model(seed, nBins, nOps):
    """atomic_domain.h (first-atom-of-each-bin table + hierarchical bitmap + linked atoms) against the reference's own
    structures restated naively — a sorted map and a swap-erase vector (ConcurrentAtomicDomain.cpp:14-132) — under
    random inserts, batched erases and in-gap moves: crowded bins (60 bins, thousands of atoms), a single bin, the
    bench-like sparse regime, positions in the last bin and at domainLength itself."""
    import ctypes as C
    from cogaps_b200._lib import lib
    done = C.c_uint64()
    rc = lib().cgb_debug_domain_fuzz(seed, nBins, nOps, C.byref(done))
    assert rc == 0 and done.value == nOps, (done.value, lib().cgb_debug_replay_message().decode())
