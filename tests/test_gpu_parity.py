"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Contract (DESIGN.md "Parity"):
  * integer / atom bookkeeping — atom-count histories, total updates, queue lengths: EXACT;
  * (s, s_mu) of the alphaParameters scans: bit-exact against the oracle run in the device's reduction
    order, and within 1e-5 * sum|terms| of the reference's own (scalar) order;
  * A/P posterior means and sds: rtol 1e-4 (north_star); chi-square: rtol 1e-4.
The oracle is run in reduce mode "device" (the kernel's association order, queried through
cgb_reduction_order_for_length) with the portable log; everything else in it is the arithmetic pinned
bit-for-bit to the reference by tests/test_oracle_vs_reference.py.
"""
import os

import numpy as np
import pytest

from tests.cases import RUN_CASES, load_data

pytestmark = pytest.mark.gpu

RTOL_MEANS = 1e-4   # north_star: "A/P posterior means within 1e-4 relative"
RTOL_CHISQ = 1e-4


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def device_options(oracle, nGenes, nSamples):
    from cogaps_b200.sampler import reduction_order_for_length
    # A sampler rows have length nSamples, P sampler rows have length nGenes
    return oracle.options(reduce="device", math="portable", orderA=reduction_order_for_length(nSamples),
                          orderP=reduction_order_for_length(nGenes))


def case_inputs(name, **extra):
    case = RUN_CASES[name]
    data = load_data(case["data"])
    kw = dict(case["params"])
    kw.update(extra)
    unc = np.maximum(0.15 * data, 0.2).astype(np.float32) if case.get("uncertainty") else None
    if case.get("fixed"):
        rows = data.shape[1] if kw["whichMatrixFixed"] == "P" else data.shape[0]
        rng = np.random.default_rng(7)
        kw["fixedPatterns"] = rng.gamma(2.0, 0.5, (rows, kw["nPatterns"])).astype(np.float32)
    return data, unc, kw


def dims(data, kw):
    g, s = (data.shape[1], data.shape[0]) if kw.get("transposeData") else data.shape
    if kw.get("subsetIndices") is not None:
        if kw.get("subsetGenes"):
            g = len(kw["subsetIndices"])
        else:
            s = len(kw["subsetIndices"])
    return g, s


def assert_close(a, b, rtol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, "%s: max error %.3g of scale (tolerance %.1g)" % (what, err, rtol)


def test_device_log_is_the_oracle_log(oracle):
    import ctypes as C
    from cogaps_b200._lib import lib, check
    from cogaps_b200._runhelp import fptr
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.random(50000), 2.0 ** rng.uniform(-126, 0, 5000), [0.0, 1.0, 0.5, 2.0 ** -149]]).astype(np.float32)
    out = np.zeros_like(xs)
    check(lib().cgb_debug_logf(fptr(xs), fptr(out), xs.size))
    ref = np.array([oracle.portable_logf(x) for x in xs], np.float32)
    host = np.array([lib().cgb_debug_host_logf(float(x)) for x in xs], np.float32)
    assert np.array_equal(bits(out), bits(ref))
    assert np.array_equal(bits(host), bits(ref))


@pytest.mark.parametrize("shape", [(37, 23, 4), (64, 40, 3), (50, 129, 6), (12, 3000, 5), (9, 11000, 3)])
@pytest.mark.parametrize("with_unc", [False, True])
def test_alpha_parameters_lockstep(oracle, shape, with_unc):
    """DenseNormalModel.cpp:162-240 through cgb_sampler_alpha_parameters (the update's own kernel)."""
    import cogaps_b200 as cg
    from cogaps_b200._runhelp import make_params
    g, s, k = shape
    rng = np.random.default_rng(g * 1000 + s)
    data = rng.gamma(2.0, 1.0, (g, s)).astype(np.float32)
    data[rng.random((g, s)) < 0.2] = 0
    A = (rng.gamma(2.0, 0.5, (g, k)) * (rng.random((g, k)) < 0.6)).astype(np.float32)
    Pm = (rng.gamma(2.0, 0.5, (s, k)) * (rng.random((s, k)) < 0.6)).astype(np.float32)
    unc = np.maximum(0.2 * data, 0.3).astype(np.float32) if with_unc else None
    q = []
    for _ in range(80):
        r1, r2 = rng.integers(0, g, 2)
        c1, c2 = rng.integers(0, k, 2)
        v = int(rng.integers(0, 3))
        if v == 1 and rng.random() < 0.5:
            r2 = r1
        q.append((v, r1, c1, r2, c2, -float(rng.random())))

    params = make_params(nPatterns=k)
    rs = cg.GapsRandomState(1)
    a = cg.GibbsSampler(data, True, True, 0.01, 100.0, params, rs)
    p = cg.GibbsSampler(data, False, False, 0.01, 100.0, params, rs)
    if unc is not None:
        a.setUncertainty(unc, True, True)
        p.setUncertainty(unc, False, False)
    a.setMatrix(A)
    p.setMatrix(Pm)
    a.sync(p)
    p.sync(a)
    a.extraInitialization()
    p.extraInitialization()
    s_gpu, smu_gpu = a.alphaParameters(q)

    opts = oracle.options(reduce="device", orderA=a.reductionOrder(), orderP=p.reductionOrder())
    s_dev, smu_dev, ap = oracle.alpha_parameters(data, A, Pm, q, uncertainty=unc, want_ap=True, options=opts)
    # AP rebuilt on the device (extraInitialization) is bit-exact
    for row in (0, g // 2, g - 1):
        assert np.array_equal(bits(a.apRow(row)), bits(ap[row]))
    assert np.array_equal(bits(s_gpu), bits(s_dev))
    assert np.array_equal(bits(smu_gpu), bits(smu_dev))
    # and against the reference's own summation order: fp32 reassociation error only
    s_ref, smu_ref = oracle.alpha_parameters(data, A, Pm, q, uncertainty=unc)
    assert np.all(np.abs(s_gpu - s_ref) <= 1e-5 * np.maximum(np.abs(s_ref), 1.0) * np.sqrt(s))
    assert np.all(np.abs(smu_gpu - smu_ref) <= 1e-5 * np.maximum(np.abs(s_ref) + np.abs(smu_ref), 1.0) * np.sqrt(s))
    # chiSq: DenseNormalModel.cpp:56-68
    cs = oracle.chisq(data, A, Pm, uncertainty=unc)
    assert a.chiSq() == pytest.approx(float(cs[0]), rel=RTOL_CHISQ)
    assert p.chiSq() == pytest.approx(float(cs[1]), rel=RTOL_CHISQ)
    assert p.dataSparsity() == pytest.approx(float(cs[2]), abs=1e-7)


def test_bulk_probe_launch_matches_oracle(oracle):
    """More queries than one parameter block holds go out as ONE launch with the proposals in device memory
    (probe_kernel); two-row queries of the same call take the pairing path.  Same bits either way."""
    import cogaps_b200 as cg
    from cogaps_b200._runhelp import make_params
    g, s, k = 700, 129, 6
    rng = np.random.default_rng(3)
    data = rng.gamma(2.0, 1.0, (g, s)).astype(np.float32)
    A = (rng.gamma(2.0, 0.5, (g, k)) * (rng.random((g, k)) < 0.6)).astype(np.float32)
    Pm = (rng.gamma(2.0, 0.5, (s, k)) * (rng.random((s, k)) < 0.6)).astype(np.float32)
    q = []
    for _ in range(900):
        r1, r2 = rng.integers(0, g, 2)
        c1, c2 = rng.integers(0, k, 2)
        v = int(rng.integers(0, 3))
        if v == 1 and rng.random() < 0.8:
            r2 = r1
        q.append((v, r1, c1, r2, c2, -float(rng.random())))
    params = make_params(nPatterns=k)
    rs = cg.GapsRandomState(1)
    a = cg.GibbsSampler(data, True, True, 0.01, 100.0, params, rs)
    p = cg.GibbsSampler(data, False, False, 0.01, 100.0, params, rs)
    a.setMatrix(A)
    p.setMatrix(Pm)
    a.sync(p)
    p.sync(a)
    a.extraInitialization()
    p.extraInitialization()
    s_gpu, smu_gpu = a.alphaParameters(q)
    opts = oracle.options(reduce="device", orderA=a.reductionOrder(), orderP=p.reductionOrder())
    s_dev, smu_dev = oracle.alpha_parameters(data, A, Pm, q, options=opts)
    assert np.array_equal(bits(s_gpu), bits(s_dev))
    assert np.array_equal(bits(smu_gpu), bits(smu_dev))
    # the same queries in small calls (parameter-space launches) give the same bits
    s_small = np.concatenate([a.alphaParameters(q[i:i + 100])[0] for i in range(0, len(q), 100)])
    assert np.array_equal(bits(s_small), bits(s_gpu))


def test_chisq_known_answer():
    """cpp_tests/testDenseGibbsSampler.cpp:11-35: A = P = 0, data(i,j) = i+j+1 on 25x50 => chiSq = 100*nRow*nCol."""
    import cogaps_b200 as cg
    from cogaps_b200._runhelp import make_params
    g, s, k = 25, 50, 7
    data = (np.add.outer(np.arange(g), np.arange(s)) + 1).astype(np.float32)
    params = make_params(nPatterns=k)
    rs = cg.GapsRandomState(123)
    a = cg.GibbsSampler(data, True, True, 0.01, 100.0, params, rs)
    p = cg.GibbsSampler(data, False, False, 0.01, 100.0, params, rs)
    a.sync(p)
    p.sync(a)
    a.extraInitialization()
    p.extraInitialization()
    assert a.chiSq() == pytest.approx(100.0 * g * s, rel=1e-6)
    assert p.chiSq() == pytest.approx(100.0 * g * s, rel=1e-6)
    assert a.nAtoms() == 0 and p.nAtoms() == 0


@pytest.mark.parametrize("name", ["modsim_async", "gist_async", "gist_transposed", "gist_uncertainty", "gist_pump",
                                  "gist_fixedP", "gist_fixedA", "gist_subset_genes", "gist_subset_samples",
                                  "syn_203x117", "syn_sparse"])
def test_run_matches_oracle(oracle, name):
    """gaps::run end to end: same seed, same data -> same chain as the oracle in device order."""
    import cogaps_b200 as cg
    data, unc, kw = case_inputs(name)
    g, s = dims(data, kw)
    opts = device_options(oracle, g, s)
    want = oracle.run(data, uncertainty=unc, snapshots=True, options=opts, **kw)
    got = cg.gaps_run(data, uncertainty=unc, snapshots=True, **kw)
    # integer bookkeeping: exact
    assert np.array_equal(got.atomHistoryA, want.atomHistoryA)
    assert np.array_equal(got.atomHistoryP, want.atomHistoryP)
    assert got.totalUpdates == want.totalUpdates
    assert np.float32(got.averageQueueLengthA) == np.float32(want.averageQueueLengthA)
    assert np.float32(got.averageQueueLengthP) == np.float32(want.averageQueueLengthP)
    # floating point: stated tolerances (in practice these are bit-identical; see the report test below)
    assert_close(got.chisqHistory, want.chisqHistory, RTOL_CHISQ, "chisqHistory")
    for f in ("Amean", "Asd", "Pmean", "Psd"):
        assert_close(getattr(got, f), getattr(want, f), RTOL_MEANS, f)
    if kw.get("whichMatrixFixed", "N") == "N":
        assert got.meanChiSq == pytest.approx(want.meanChiSq, rel=RTOL_CHISQ)
    else:
        assert got.meanChiSq == 0.0
    if kw.get("snapshotFrequency"):
        assert np.array_equal(bits(got.snapshotsA[-1]), bits(want.snapshotsA[-1]))
        assert np.array_equal(bits(got.snapshotsP[-1]), bits(want.snapshotsP[-1]))
    if RUN_CASES[name].get("pump"):
        assert np.array_equal(got.pumpMatrix, want.pumpMatrix)
        assert np.array_equal(got.meanPatternAssignment, want.meanPatternAssignment)


@pytest.mark.parametrize("shape", [(37, 23, 4), (64, 300, 7), (50, 700, 30), (12, 3000, 5)])
def test_sparse_alpha_parameters_lockstep(oracle, shape):
    """SparseNormalModel.cpp:153-292 (all three alphaParameters variants), generateLookupTables (:294-311) and
    chiSq (:39-60) through the sparse eval kernel, against the oracle in the device's summation order
    (bit-exact) and in the reference's own order (fp32 re-association only)."""
    import cogaps_b200 as cg
    from cogaps_b200._runhelp import make_params
    g, s, k = shape
    rng = np.random.default_rng(g * 1000 + s)
    data = rng.gamma(2.0, 1.0, (g, s)).astype(np.float32)
    data[rng.random((g, s)) < 0.8] = 0
    data[0, :] = 0          # an empty data row
    data[1, :] = 1.5        # and a full one (longer than one 256-wide group when s > 256)
    A = (rng.gamma(2.0, 0.5, (g, k)) * (rng.random((g, k)) < 0.6)).astype(np.float32)
    Pm = (rng.gamma(2.0, 0.5, (s, k)) * (rng.random((s, k)) < 0.6)).astype(np.float32)
    A[2, 0] = 5e-6          # below epsilon: kept in the row copy, dropped from the column copy
    Pm[3, 1] = 5e-6
    q = [(0, 0, 0, 0, 0, 0.0), (0, 1, 1, 1, 1, 0.0), (1, 1, 0, 1, 1, 0.0), (2, 1, 1, 1, 1, -0.3)]
    for _ in range(120):
        r1, r2 = rng.integers(0, g, 2)
        c1, c2 = rng.integers(0, k, 2)
        v = int(rng.integers(0, 3))
        if v == 1 and rng.random() < 0.6:
            r2 = r1
        q.append((v, r1, c1, r2, c2, -float(rng.random())))

    params = make_params(nPatterns=k, useSparseOptimization=1)
    rs = cg.GapsRandomState(1)
    a = cg.GibbsSampler(data, True, True, 0.01, 100.0, params, rs)
    p = cg.GibbsSampler(data, False, False, 0.01, 100.0, params, rs)
    a.setMatrix(A)
    p.setMatrix(Pm)
    a.sync(p)
    p.sync(a)
    a.extraInitialization()
    p.extraInitialization()
    s_gpu, smu_gpu = a.alphaParameters(q)
    opts = oracle.options(reduce="device", orderA=a.reductionOrder(), orderP=p.reductionOrder())
    s_dev, smu_dev = oracle.alpha_parameters_sparse(data, A, Pm, q, options=opts)
    assert np.array_equal(bits(s_gpu), bits(s_dev))
    assert np.array_equal(bits(smu_gpu), bits(smu_dev))
    s_ref, smu_ref = oracle.alpha_parameters_sparse(data, A, Pm, q)
    # s and s_mu are differences of large table terms and scan sums (times beta = 100): bound the error by the
    # magnitude of the pieces, not of the result
    scale = 100.0 * (np.abs(A).max() * np.abs(Pm).max() * k + 1.0) * np.abs(Pm).max() * s
    assert np.all(np.abs(s_gpu - s_ref) <= 1e-6 * scale)
    assert np.all(np.abs(smu_gpu - smu_ref) <= 1e-6 * scale)
    cs = oracle.chisq_sparse(data, A, Pm)
    assert a.chiSq() == pytest.approx(float(cs[0]), rel=RTOL_CHISQ)
    assert p.chiSq() == pytest.approx(float(cs[1]), rel=RTOL_CHISQ)
    assert p.dataSparsity() == pytest.approx(float(cs[2]), abs=1e-7)


def test_sparse_chisq_known_answer():
    """cpp_tests/testSparseGibbsSampler.cpp:13-33: A = P = 0, data(i,j) = i+j+1 on 25x50 => chiSq = 100*nRow*nCol."""
    import cogaps_b200 as cg
    from cogaps_b200._runhelp import make_params
    g, s, k = 25, 50, 7
    data = (np.add.outer(np.arange(g), np.arange(s)) + 1).astype(np.float32)
    params = make_params(nPatterns=k, useSparseOptimization=1)
    rs = cg.GapsRandomState(123)
    a = cg.GibbsSampler(data, True, True, 0.01, 100.0, params, rs)
    p = cg.GibbsSampler(data, False, False, 0.01, 100.0, params, rs)
    a.sync(p)
    p.sync(a)
    assert a.chiSq() == pytest.approx(100.0 * g * s, rel=1e-6)
    assert p.chiSq() == pytest.approx(100.0 * g * s, rel=1e-6)


@pytest.mark.parametrize("name", ["sparse_gist", "sparse_modsim", "sparse_120x90", "sparse_k30", "sparse_gist_fixedP"])
def test_sparse_run_matches_oracle(oracle, name):
    """gaps::run with sparseOptimization (SparseGibbsSampler): same seed, same data -> same chain as the oracle
    in device order; the oracle's sparse model is pinned bit-for-bit to the reference build."""
    import cogaps_b200 as cg
    data, unc, kw = case_inputs(name)
    g, s = dims(data, kw)
    opts = device_options(oracle, g, s)
    want = oracle.run(data, snapshots=True, options=opts, **kw)
    got = cg.gaps_run(data, snapshots=True, **kw)
    assert np.array_equal(got.atomHistoryA, want.atomHistoryA)
    assert np.array_equal(got.atomHistoryP, want.atomHistoryP)
    assert got.totalUpdates == want.totalUpdates
    assert np.float32(got.averageQueueLengthA) == np.float32(want.averageQueueLengthA)
    assert np.float32(got.averageQueueLengthP) == np.float32(want.averageQueueLengthP)
    assert_close(got.chisqHistory, want.chisqHistory, RTOL_CHISQ, "chisqHistory")
    for f in ("Amean", "Asd", "Pmean", "Psd"):
        assert_close(getattr(got, f), getattr(want, f), RTOL_MEANS, f)
    if kw.get("whichMatrixFixed", "N") == "N":
        assert got.meanChiSq == pytest.approx(want.meanChiSq, rel=RTOL_CHISQ)
    else:
        assert got.meanChiSq == 0.0


@pytest.mark.parametrize("name", ["modsim_seq", "gist_seq", "sparse_gist_seq"])
@pytest.mark.parametrize("resident", [True, False])
def test_sequential_sampler_matches_oracle(oracle, name, resident, monkeypatch):
    """asynchronousUpdates = FALSE: SingleThreadedGibbsSampler (gibbs_sampler/SingleThreadedGibbsSampler.h:94-257) —
    one rng stream shared by generation and evaluation, birth accepts mass > epsilon, same-bin exchanges ignored.
    This is the sampler config C1 names and the one distributed CoGAPS forces (R/DistributedCogaps.R:28-29)."""
    import cogaps_b200 as cg
    if not resident:
        monkeypatch.setenv("COGAPS_PERSISTENT", "0")
    data, unc, kw = case_inputs(name)
    if name != "modsim_seq":
        kw["nIterations"] = min(kw["nIterations"], 30)
    g, s = dims(data, kw)
    want = oracle.run(data, snapshots=True, options=device_options(oracle, g, s), **kw)
    got = cg.gaps_run(data, snapshots=True, **kw)
    assert np.array_equal(got.atomHistoryA, want.atomHistoryA)
    assert np.array_equal(got.atomHistoryP, want.atomHistoryP)
    assert got.totalUpdates == want.totalUpdates
    assert got.averageQueueLengthA == 0.0 and got.averageQueueLengthP == 0.0
    assert_close(got.chisqHistory, want.chisqHistory, RTOL_CHISQ, "chisqHistory")
    for f in ("Amean", "Asd", "Pmean", "Psd"):
        assert_close(getattr(got, f), getattr(want, f), RTOL_MEANS, f)
    assert got.meanChiSq == pytest.approx(want.meanChiSq, rel=RTOL_CHISQ)


def test_posterior_means_are_bit_identical(oracle):
    """Stronger than the stated tolerance: with the oracle in the device's reduction order the whole
    chain — every atom, every mass — is reproduced to the last bit."""
    import cogaps_b200 as cg
    data, unc, kw = case_inputs("gist_async", nIterations=60, snapshotFrequency=20)
    g, s = dims(data, kw)
    want = oracle.run(data, snapshots=True, options=device_options(oracle, g, s), **kw)
    got = cg.gaps_run(data, snapshots=True, **kw)
    for f in ("Amean", "Asd", "Pmean", "Psd", "snapshotsA", "snapshotsP"):
        assert np.array_equal(bits(getattr(got, f)), bits(getattr(want, f))), f


@pytest.mark.parametrize("name", ["gist_async", "syn_203x117", "sparse_120x90"])
def test_launch_per_batch_and_resident_grid_give_the_same_chain(name, monkeypatch):
    """cgb_sampler_set_persistent: the resident streaming grid and one kernel launch per conflict-free batch run
    the same device code on the same proposals, so every output bit agrees."""
    import cogaps_b200 as cg
    data, unc, kw = case_inputs(name, nIterations=40, snapshotFrequency=20)
    resident = cg.gaps_run(data, snapshots=True, **kw)
    monkeypatch.setenv("COGAPS_PERSISTENT", "0")
    launched = cg.gaps_run(data, snapshots=True, **kw)
    assert np.array_equal(resident.atomHistoryA, launched.atomHistoryA)
    assert launched.nBatchesA > 0 and resident.nBatchesA > 0
    for f in ("Amean", "Asd", "Pmean", "Psd", "snapshotsA", "snapshotsP", "chisqHistory"):
        assert np.array_equal(bits(getattr(resident, f)), bits(getattr(launched, f))), f


@pytest.mark.parametrize("knobs", [
    dict(COGAPS_FORCE_ROW_WAIT="1"),                                   # every task on a touched row spins on rowVersion
    dict(COGAPS_PERSISTENT_CLUSTERS="3"),                              # 2 worker clusters: record rings wrap inside a batch
    dict(COGAPS_CHUNK_PROPOSALS="7"),                                  # batches longer than a chunk: posted in several
    dict(COGAPS_PERSISTENT_CLUSTERS="4", COGAPS_CHUNK_PROPOSALS="40", COGAPS_FORCE_ROW_WAIT="1"),
])
@pytest.mark.parametrize("name", ["gist_async", "sparse_120x90"])
def test_rarely_taken_paths_of_the_resident_grid(name, knobs, monkeypatch):
    """The resident grid's fallbacks — rowVersion waits (normally skipped once the host has proof), ring-slot reuse
    (normally there are more slots than tasks in flight), chunked batches (normally a batch fits one chunk) — forced
    by environment knobs read at sampler creation; the chain must not change by a bit."""
    import cogaps_b200 as cg
    data, unc, kw = case_inputs(name, nIterations=40, snapshotFrequency=20)
    plain = cg.gaps_run(data, snapshots=True, **kw)
    for k, v in knobs.items():
        monkeypatch.setenv(k, v)
    forced = cg.gaps_run(data, snapshots=True, **kw)
    assert np.array_equal(plain.atomHistoryA, forced.atomHistoryA)
    for f in ("Amean", "Asd", "Pmean", "Psd", "snapshotsA", "snapshotsP", "chisqHistory"):
        assert np.array_equal(bits(getattr(plain, f)), bits(getattr(forced, f))), f


@pytest.mark.skipif(os.environ.get("COGAPS_RUN_STRESS", "0") != "1",
                    reason="opt-in stress test (COGAPS_RUN_STRESS=1): whole runs, setup and teardown included, from several threads")
def test_chains_sharing_the_device_do_not_disturb_each_other():
    """cgb_set_resident_share: three chains driven by three host threads at once give, each, the bits it gives alone."""
    import threading
    import cogaps_b200 as cg
    from cogaps_b200._lib import check
    cases = ["gist_async", "syn_203x117", "sparse_120x90"]
    inputs = [case_inputs(n, nIterations=40) for n in cases]
    alone = [cg.gaps_run(d, **kw) for d, _, kw in inputs]
    together = [None] * len(cases)

    def drive(i):
        d, _, kw = inputs[i]
        together[i] = cg.gaps_run(d, **kw)

    check(cg.lib().cgb_set_resident_share(len(cases)))
    try:
        threads = [threading.Thread(target=drive, args=(i,)) for i in range(len(cases))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    finally:
        check(cg.lib().cgb_set_resident_share(1))
    for a, b in zip(alone, together):
        assert b is not None
        assert np.array_equal(a.atomHistoryA, b.atomHistoryA)
        for f in ("Amean", "Pmean", "Asd", "Psd"):
            assert np.array_equal(bits(getattr(a, f)), bits(getattr(b, f))), f


def test_long_rows_use_clusters(oracle):
    """Row lengths beyond one segment: the scan is split over a thread-block cluster (DSMEM reduce)."""
    import cogaps_b200 as cg
    from cogaps_b200.sampler import reduction_order_for_length
    data = load_data("syn:40:14000:4:3")     # A rows have 14000 samples -> 3 segments
    assert reduction_order_for_length(14000)[2] > 1
    kw = dict(seed=5, nPatterns=4, nIterations=25, outputFrequency=5, maxThreads=1)
    want = oracle.run(data, options=device_options(oracle, 40, 14000), **kw)
    got = cg.gaps_run(data, **kw)
    assert np.array_equal(got.atomHistoryA, want.atomHistoryA)
    assert np.array_equal(got.atomHistoryP, want.atomHistoryP)
    for f in ("Amean", "Pmean"):
        assert_close(getattr(got, f), getattr(want, f), RTOL_MEANS, f)
    assert_close(got.chisqHistory, want.chisqHistory, RTOL_CHISQ, "chisqHistory")


def test_seed_consistency_and_statistical_agreement_with_reference(oracle, golden):
    """test_seed_consistency.R:13-21: same seed twice -> identical; and against the reference build's own
    golden chain (a different summation order, so a different but statistically equivalent chain):
    atoms and chi-square within the spread the reference's two builds show between themselves."""
    import cogaps_b200 as cg
    data, unc, kw = case_inputs("gist_async")
    r1 = cg.gaps_run(data, **kw)
    r2 = cg.gaps_run(data, **kw)
    assert np.array_equal(r1.atomHistoryA, r2.atomHistoryA)
    assert np.array_equal(bits(r1.Amean), bits(r2.Amean))
    ref = golden["scalar/gist_async/atomHistoryA"].astype(np.float64)
    assert abs(float(r1.atomHistoryA[-1]) - ref[-1]) / ref[-1] < 0.15
    refc = golden["scalar/gist_async/chisqHistory"].astype(np.float64)
    assert abs(float(r1.chisqHistory[-1]) - refc[-1]) / refc[-1] < 0.25


def test_user_api_mirrors_reference():
    """CoGAPS() / CogapsParams / CogapsResult (R/CoGAPS.R:90-156)."""
    import cogaps_b200 as cg
    data = load_data("gist")
    params = cg.CogapsParams(nPatterns=3, nIterations=40, seed=42)
    res = cg.CoGAPS(data, params, outputFrequency=10, messages=False)
    assert res.featureLoadings.shape == (1363, 3) and res.sampleFactors.shape == (9, 3)   # test_top_level.R:11-30
    assert not np.isnan(res.featureLoadings).any() and not np.isnan(res.sampleFactors).any()
    assert len(res.metadata["chisq"]) == 8 and len(set(res.metadata["chisq"].tolist())) == 8
    # test_chisq.R:1-17 — meanChiSq equals the recomputation from the returned means
    A, P = res.featureLoadings.astype(np.float64), res.sampleFactors.astype(np.float64)
    S = np.maximum(0.1 * data.astype(np.float64), 0.1)
    recomputed = (((data - A @ P.T) / S) ** 2).sum()
    assert res.getMeanChiSq() == pytest.approx(recomputed, rel=1e-4)
    with pytest.raises(ValueError):
        cg.CoGAPS(data, params, notAParameter=1)
    with pytest.raises(ValueError):
        cg.CoGAPS(-data, params)


def test_distributed_modes_single_process():
    """test_top_level.R:84-118 — genome-wide and single-cell distributed CoGAPS return full-size matrices with no
    NA; here every subset runs on this process's GPU (the multi-rank plumbing is covered over gloo on CPU)."""
    import cogaps_b200 as cg
    data = load_data("gist")
    gw = cg.CogapsParams(nPatterns=3, nIterations=60, seed=42, distributed="genome-wide")
    res = cg.CoGAPS(data, gw, messages=False, outputFrequency=30)
    assert res.featureLoadings.shape[0] == 1363 and res.sampleFactors.shape[0] == 9
    assert res.featureLoadings.shape[1] == res.sampleFactors.shape[1] >= 1
    assert not np.isnan(res.featureLoadings).any() and not np.isnan(res.sampleFactors).any()
    assert len(res.metadata["subsets"]) == 4 and res.metadata["meanChiSq"] == 0.0   # fixed-matrix runs report 0
    sc = cg.CogapsParams(nPatterns=2, nIterations=60, seed=42, distributed="single-cell")
    sc.setParam("nSets", 2)
    res = cg.CoGAPS(data, sc, messages=False, outputFrequency=30)
    assert res.featureLoadings.shape[0] == 1363 and res.sampleFactors.shape[0] == 9
    assert not np.isnan(res.sampleFactors).any()


def test_distributed_result_fits_the_data():
    """What the two-pass distributed driver returns is a factorisation of the WHOLE matrix (R/DistributedCogaps.R:48-119:
    per-subset runs, consensus matrix, second pass with it held fixed, rows stitched back in place): its chi-square against
    the data — with the default uncertainty max(0.1 D, 0.1) — must be of the order of a plain run's with the same number of
    patterns, and far below that of the best rank-0 model (every gene at its mean)."""
    import cogaps_b200 as cg
    from tests.cases import synthetic
    data = synthetic("syn:400:96:4:3")
    sd = np.maximum(0.1 * data, 0.1)

    def chisq(A, P):
        return float((((data - A.astype(np.float64) @ P.astype(np.float64).T) / sd) ** 2).sum())

    plain = cg.CoGAPS(data, cg.CogapsParams(nPatterns=4, nIterations=300, seed=7), messages=False, outputFrequency=150)
    base = chisq(plain.featureLoadings, plain.sampleFactors)
    null = float((((data - data.mean(axis=1, keepdims=True)) / sd) ** 2).sum())
    assert base < 0.2 * null
    for mode, nsets in (("genome-wide", 4), ("single-cell", 3)):
        p = cg.CogapsParams(nPatterns=4, nIterations=300, seed=7, distributed=mode)
        p.setParam("nSets", nsets)
        res = cg.CoGAPS(data, p, messages=False, outputFrequency=150)
        assert res.featureLoadings.shape == (400, res.sampleFactors.shape[1]) and res.sampleFactors.shape[0] == 96
        got = chisq(res.featureLoadings, res.sampleFactors)
        assert got < 3.0 * base and got < 0.3 * null, (mode, got, base, null)


@pytest.mark.parametrize("sparse", [0, 1])
def test_comm_allgather_returns_the_factor_rows(sparse):
    """cgb_comm_init + cgb_allgather_rows (the C ABI's NCCL communicator) on a one-rank communicator: the rows
    gathered from device memory are the values cgb_sampler_get_matrix returns — the dense model's matrix, the sparse model's
    row copy (not its column copy, in which values below epsilon are stored as 0).  tools/dist_check.py and bench.py's c5
    record cover 2 / 4 / 8 ranks (profiles/r2_dist_check_8gpu.log)."""
    import bench
    import cogaps_b200 as cg
    data = load_data("gist")
    chain = bench.Chain(data, 4, 17, sparse=bool(sparse), updateMode=1)
    chain.ramp(60)
    try:
        comm = cg.Comm(cg.Comm.unique_id(), 0, 1)
    except cg.CogapsError as e:
        pytest.skip("NCCL is not loadable here: %s" % e)
    for smp in (chain.P, chain.A):
        want = smp.getMatrix()
        got = comm.allgatherRows(smp, [want.shape[0]])
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert comm.last_ms >= 0.0
    assert (chain.P.getMatrix() > 0).any()


def test_full_size_invariants_20000x5000_k20():
    """BASELINE.json configs[2] at full size (the bench workload), through properties that need no oracle run:
    (1) the resident grid and one launch per batch give the same chain, bit for bit, with 3-CTA clusters on the
        P side (L = 20000);
    (2) the incrementally maintained AP equals A * P^T rebuilt from the factors: chi-square before and after
        extraInitialization() agree to fp32 round-off, and the A-side and P-side chi-square are the same number;
    (3) mass conservation: every factor element is the sum of the masses of the atoms in its bin
        (changeMatrix / safelyChangeMatrix, DenseNormalModel.cpp:110-123), and no element is negative."""
    import bench
    import cogaps_b200 as cg
    g, s, k = 20000, 5000, 20
    data = bench.make_data(g, s, k)

    def run(persistent, iters=26):
        chain = bench.Chain(data, k, 42)
        for smp in (chain.A, chain.P):
            smp.setPersistent(persistent)
        updates = 0
        for i in range(iters):
            temp = min(1.0, 2.0 * i / iters)
            chain.A.setAnnealingTemp(temp)
            chain.P.setAnnealingTemp(temp)
            updates += chain.step()
        return chain, updates

    resident, n1 = run(True)
    launched, n2 = run(False)
    assert resident.P.reductionOrder()[2] > 1                      # clusters on the long rows
    assert n1 == n2 and resident.A.nAtoms() == launched.A.nAtoms() and resident.P.nAtoms() == launched.P.nAtoms()
    assert resident.A.nAtoms() > 1000
    Ares, Pres = resident.A.getMatrix(), resident.P.getMatrix()
    assert np.array_equal(bits(Ares), bits(launched.A.getMatrix()))
    assert np.array_equal(bits(Pres), bits(launched.P.getMatrix()))
    csA, csP = resident.A.chiSq(), resident.P.chiSq()
    assert csA == launched.A.chiSq() and csP == launched.P.chiSq()
    assert csA == pytest.approx(csP, rel=RTOL_CHISQ)
    # (2) AP kept by ~10^5 rank-one commits vs rebuilt from the factors
    resident.A.extraInitialization()
    resident.P.extraInitialization()
    assert resident.A.chiSq() == pytest.approx(csA, rel=RTOL_CHISQ)
    assert resident.P.chiSq() == pytest.approx(csP, rel=RTOL_CHISQ)
    # (3) atoms <-> matrix
    for smp, M in ((resident.A, Ares), (resident.P, Pres)):
        assert (M >= 0).all()
        pos, mass = smp.atoms()
        nbins = M.shape[0] * k
        binlen = np.uint64(0xFFFFFFFFFFFFFFFF // nbins)
        bins = np.minimum((pos // binlen).astype(np.int64), nbins - 1)
        summed = np.bincount(bins, weights=mass.astype(np.float64), minlength=nbins).reshape(M.shape[0], k)
        assert np.allclose(summed, M.astype(np.float64), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------------
# BASELINE sizes against the oracle, and the reference's own chain over seeds ("close the parity gaps")
# ------------------------------------------------------------------------------------------------
def test_c3_full_size_run_matches_oracle(oracle):
    """BASELINE.json configs[2] at FULL SIZE (20000 x 5000, nPatterns = 20) through cgb_run against the oracle in device
    order: 10 + 10 iterations from zero atoms (tens of thousands of proposals; 3-CTA clusters on the P side's
    20000-long rows) — atom-count histories and update counts exact, final factor matrices bit for bit, posterior
    means / sds and chi-square within 1e-4."""
    import bench
    import cogaps_b200 as cg
    g, s, k = 20000, 5000, 20
    data = bench.make_data(g, s, k)
    kw = dict(seed=42, nPatterns=k, nIterations=10, outputFrequency=5, maxThreads=1, snapshotFrequency=10)
    want = oracle.run(data, snapshots=True, options=device_options(oracle, g, s), **kw)
    got = cg.gaps_run(data, snapshots=True, **kw)
    assert want.totalUpdates > 30000
    assert np.array_equal(got.atomHistoryA, want.atomHistoryA), (got.atomHistoryA, want.atomHistoryA)
    assert np.array_equal(got.atomHistoryP, want.atomHistoryP)
    assert got.totalUpdates == want.totalUpdates
    assert np.float32(got.averageQueueLengthA) == np.float32(want.averageQueueLengthA)
    assert np.float32(got.averageQueueLengthP) == np.float32(want.averageQueueLengthP)
    assert np.array_equal(bits(got.snapshotsA[-1]), bits(want.snapshotsA[-1]))
    assert np.array_equal(bits(got.snapshotsP[-1]), bits(want.snapshotsP[-1]))
    for f in ("Amean", "Asd", "Pmean", "Psd"):
        assert_close(getattr(got, f), getattr(want, f), RTOL_MEANS, f)
    # chi-square over 10^8 elements: the reference (and the oracle with it) keeps ONE fp32 running sum
    # (DenseNormalModel.cpp:56-68; GapsStatistics.cpp:63-87).  At this size that sum stalls — once it passes ~4e9 a term of
    # a few tens is below half an ulp and is dropped, and the early chi-square here is ~1e10 — so the oracle's number is
    # not a reference value (it is off by a factor, and so is the reference's).  The device accumulates in f64; it is
    # held to an f64 evaluation of sum(((D - A P^T) / S)^2) from the final factor matrices (the last report and the last
    # snapshot coincide), and meanChiSq to the same sum over the posterior means.
    def chi(A, P):
        A, P = A.astype(np.float64), P.astype(np.float64)
        total = 0.0
        for r0 in range(0, g, 2000):
            d = data[r0:r0 + 2000].astype(np.float64)
            sd = np.maximum(np.float32(0.1) * data[r0:r0 + 2000], np.float32(0.1)).astype(np.float64)
            total += float((((d - A[r0:r0 + 2000] @ P.T) / sd) ** 2).sum())
        return total
    assert got.chisqHistory[-1] == pytest.approx(chi(got.snapshotsA[-1], got.snapshotsP[-1]), rel=RTOL_CHISQ)
    assert got.meanChiSq == pytest.approx(chi(got.Amean, got.Pmean), rel=RTOL_CHISQ)
    assert abs(want.chisqHistory[-1] - got.chisqHistory[-1]) > 1e-3 * got.chisqHistory[-1]   # the stall, on record


@pytest.mark.parametrize("spec,k", [("spz:3000:30000:8:7:95", 50), ("spz:26000:2500:8:9:95", 50)])
def test_c4_shaped_sparse_run_matches_oracle(oracle, spec, k):
    """BASELINE.json configs[3]'s row shape through the sparse model against the oracle: 30000-long A rows at 95 % zeros
    (about 1500 non-zeros per row: more than one 1024-entry group per scan) and, transposed, 26000-long P rows, with
    nPatterns = 50 (k > 25: gaps::dot accumulates forwards, VectorMath.h:40-98).  The full 50000 x 30000 is the same
    code on more rows (bench.py --sparse); its oracle run would take tens of GB of host memory."""
    import cogaps_b200 as cg
    data = load_data(spec)
    g, s = data.shape
    kw = dict(seed=17, nPatterns=k, nIterations=8, outputFrequency=4, maxThreads=1, useSparseOptimization=1, snapshotFrequency=8)
    want = oracle.run(data, snapshots=True, options=device_options(oracle, g, s), **kw)
    got = cg.gaps_run(data, snapshots=True, **kw)
    assert want.totalUpdates > 3000
    assert np.array_equal(got.atomHistoryA, want.atomHistoryA), (got.atomHistoryA, want.atomHistoryA)
    assert np.array_equal(got.atomHistoryP, want.atomHistoryP)
    assert got.totalUpdates == want.totalUpdates
    assert np.array_equal(bits(got.snapshotsA[-1]), bits(want.snapshotsA[-1]))
    assert np.array_equal(bits(got.snapshotsP[-1]), bits(want.snapshotsP[-1]))
    assert_close(got.chisqHistory, want.chisqHistory, RTOL_CHISQ, "chisqHistory")
    for f in ("Amean", "Pmean"):
        assert_close(getattr(got, f), getattr(want, f), RTOL_MEANS, f)


@pytest.mark.parametrize("name,k,its", [("gist", 7, 600), ("syn:203:117:5:11", 5, 600)])
@pytest.mark.parametrize("mode", [0, 1])
def test_tier3_chains_agree_with_the_reference_over_seeds(name, k, its, mode):
    """SURVEY 7.4-2 Tier 3: free-running GPU chains (exact mode, and the row-parallel sweep) against the REFERENCE ITSELF
    (oracle/_ref, scalar build — whose own chain differs from any device order after a few iterations) over
    six seeds: the chi-square trajectories and meanChiSq agree inside the seed-to-seed spread (|difference of means| <= 3
    standard errors, floor 3 %), and they and the atom counts within the tolerance the reference sets for "the same result"
    (0.1 relative, tests/testthat/test_seed_consistency.R:13-21; atom counts 0.15, sweep 0.2) or one seed-to-seed standard
    deviation — from the end of the equilibration phase on."""
    import cogaps_b200 as cg
    from oracle.harness import RefLib
    if not RefLib.available("scalar"):
        pytest.skip("oracle/_ref (the compiled reference) did not travel to this box")
    ref = RefLib("scalar")
    data = load_data(name)
    seeds = [1, 3, 5, 7, 9, 11]           # the seeder ORs 1 into the seed: odd seeds are distinct chains
    rows_ref, rows_gpu = [], []
    for seed in seeds:
        kw = dict(seed=seed, nPatterns=k, nIterations=its, outputFrequency=its // 3)
        a = ref.run(data, **kw)
        b = cg.gaps_run(data, updateMode=mode, **kw)
        rows_ref.append(np.concatenate([a.atomHistoryA, a.atomHistoryP, a.chisqHistory, [a.meanChiSq]]).astype(np.float64))
        rows_gpu.append(np.concatenate([b.atomHistoryA, b.atomHistoryP, b.chisqHistory, [b.meanChiSq]]).astype(np.float64))
    R, Gm = np.array(rows_ref), np.array(rows_gpu)
    nh = (R.shape[1] - 1) // 3
    for j in range(R.shape[1]):
        if j % nh < 2 and j < 3 * nh:
            continue                       # the first two reports fall in the annealed transient (temperature < 1 until
                                           # half of the equilibration phase), where trajectories are steep
        mr, mg = R[:, j].mean(), Gm[:, j].mean()
        sd = np.sqrt(0.5 * (R[:, j].var(ddof=1) + Gm[:, j].var(ddof=1)))
        se = sd * np.sqrt(2.0 / len(seeds))
        # atom counts: 0.15 for the reference's own chain in another summation order; 0.2 for the sweep, whose frozen
        # birth / death balance shifts the count of a 9-row matrix by about a tenth
        tol = (0.2 if mode == 1 else 0.15) if j < 2 * nh else 0.1
        what = "column %d (%s)" % (j, "atoms" if j < 2 * nh else "chi-square"), mr, mg, sd
        if j >= 2 * nh:
            assert abs(mr - mg) <= max(3.0 * se, 0.03 * abs(mr)), what      # the fit: no significant difference
        assert abs(mr - mg) <= max(tol * abs(mr), sd), what


@pytest.mark.parametrize("mode", [0, 1])
def test_tier3_midsize_chains_agree_with_the_reference(mode):
    """The same comparison on a matrix between the reference's own data sets and the BASELINE shapes: 1200 x 500, 8
    patterns, 300 + 300 iterations, six seeds.  The reference's trajectories (scalar build, 20 s per seed on a host core)
    travel as tests/golden/tier3_midsize_ref.npz (generator: tests/golden/make_tier3_midsize.py); the free-running CUDA
    chains — exact mode and the row-parallel sweep — are held to them from the end of the annealed transient on: chi-square
    and meanChiSq inside the seed-to-seed spread (|difference of means| <= 3 standard errors, floor 3 %), atom counts
    within 0.15 (sweep 0.2) or one seed-to-seed standard deviation."""
    import cogaps_b200 as cg
    fix = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tier3_midsize_ref.npz"))
    data = load_data(str(fix["spec"]))
    k, its, freq = int(fix["k"]), int(fix["its"]), int(fix["freq"])
    R = fix["rows"].astype(np.float64)
    rows = []
    for seed in fix["seeds"]:
        b = cg.gaps_run(data, updateMode=mode, seed=int(seed), nPatterns=k, nIterations=its, outputFrequency=freq)
        rows.append(np.concatenate([b.atomHistoryA, b.atomHistoryP, b.chisqHistory, [b.meanChiSq]]).astype(np.float64))
    Gm = np.array(rows)
    assert Gm.shape == R.shape
    nh = (R.shape[1] - 1) // 3
    for j in range(R.shape[1]):
        if j % nh < 2 and j < 3 * nh:
            continue                       # annealed transient (temperature < 1 until half of the equilibration phase)
        mr, mg = R[:, j].mean(), Gm[:, j].mean()
        sd = np.sqrt(0.5 * (R[:, j].var(ddof=1) + Gm[:, j].var(ddof=1)))
        se = sd * np.sqrt(2.0 / R.shape[0])
        tol = (0.2 if mode == 1 else 0.15) if j < 2 * nh else 0.1
        what = "column %d (%s)" % (j, "atoms" if j < 2 * nh else "chi-square"), mr, mg, sd
        if j >= 2 * nh:
            assert abs(mr - mg) <= max(3.0 * se, 0.03 * abs(mr)), what
        assert abs(mr - mg) <= max(tol * abs(mr), sd), what


# ------------------------------------------------------------------------------------------------
# the reference-side binding, compiled (oracle/cuda_adapter.cpp -> oracle/_ref/libcogaps_ref_adapter.so)
# ------------------------------------------------------------------------------------------------
def _adapter():
    from oracle.harness import RefLib
    if not RefLib.available("adapter"):
        pytest.skip("oracle/_ref/libcogaps_ref_adapter.so (the reference's run loop + CudaGibbsSampler) did not travel to this box")
    return RefLib("adapter")


@pytest.mark.parametrize("name", ["gist_async", "gist_uncertainty", "gist_fixedP", "gist_subset_genes", "gist_pump", "syn_203x117"])
def test_reference_run_loop_drives_the_c_abi(name):
    """The reference's OWN runCoGAPSAlgorithm<> / runOnePhase / GapsStatistics (src/GapsRunner.cpp:272-327,381-503,
    GapsStatistics.h:129-202), instantiated with CudaGibbsSampler — the class INTEGRATION.md tells a maintainer to add,
    here compiled against the unmodified reference — drives the device through the C ABI and returns what cgb_run
    returns: the same chain (atom histories, update count, queue lengths) and, because the reference's statistics then
    run on the host over the same factor matrices in the same order, the same posterior means bit for bit."""
    import cogaps_b200 as cg
    adapter = _adapter()
    data, unc, kw = case_inputs(name)
    want = cg.gaps_run(data, uncertainty=unc, **kw)
    got = adapter.run(data, uncertainty=unc, **kw)
    assert np.array_equal(got.atomHistoryA, want.atomHistoryA), (got.atomHistoryA, want.atomHistoryA)
    assert np.array_equal(got.atomHistoryP, want.atomHistoryP)
    assert got.totalUpdates == want.totalUpdates
    assert np.float32(got.averageQueueLengthA) == np.float32(want.averageQueueLengthA)
    assert np.float32(got.averageQueueLengthP) == np.float32(want.averageQueueLengthP)
    assert_close(got.chisqHistory, want.chisqHistory, RTOL_CHISQ, "chisqHistory")
    for f in ("Amean", "Asd", "Pmean", "Psd"):
        assert np.array_equal(bits(getattr(got, f)), bits(getattr(want, f))), f
    if kw.get("whichMatrixFixed", "N") == "N":
        # the reference sums meanChiSq in one fp32 running sum on the host, the library in f64 on the device
        assert got.meanChiSq == pytest.approx(want.meanChiSq, rel=RTOL_CHISQ)
    if RUN_CASES[name].get("pump"):
        assert np.array_equal(got.pumpMatrix, want.pumpMatrix)


def test_reference_checkpoint_code_archives_the_cuda_sampler(tmp_path):
    """createCheckpoint / processCheckpoint of the reference itself (GapsRunner.cpp:224-270) over CudaGibbsSampler's
    `Archive <<` / `>>`: the file the reference's loop writes while it drives the device is byte for byte the file
    cgb_run_ex writes, and the reference's loop resumed from it ends where the uninterrupted run ends."""
    import cogaps_b200 as cg
    adapter = _adapter()
    data = load_data("gist")
    kw = dict(seed=42, nPatterns=5, nIterations=40, outputFrequency=10, maxThreads=1)
    f_ref, f_lib = tmp_path / "adapter.out", tmp_path / "library.out"
    a = adapter.run(data, checkpointInterval=25, checkpointOutFile=f_ref, **kw)
    b = cg.gaps_run(data, checkpointInterval=25, checkpointOutFile=f_lib, **kw)
    assert f_ref.read_bytes() == f_lib.read_bytes()
    assert np.array_equal(a.atomHistoryA, b.atomHistoryA) and np.array_equal(bits(a.Amean), bits(b.Amean))
    c = adapter.run(data, checkpointInFile=f_ref, **kw)
    assert np.array_equal(bits(c.Amean), bits(a.Amean)) and np.array_equal(bits(c.Pmean), bits(a.Pmean))


def test_reference_run_loop_in_sweep_mode(monkeypatch):
    """the same binding with the row-parallel sweep selected through the environment (an R caller cannot pass a new
    parameter): the reference's loop over the sweep equals cgb_run(updateMode = sweep)"""
    import cogaps_b200 as cg
    adapter = _adapter()
    data = load_data("gist")
    kw = dict(seed=7, nPatterns=5, nIterations=60, outputFrequency=20, maxThreads=1)
    want = cg.gaps_run(data, updateMode=1, **kw)
    monkeypatch.setenv("COGAPS_UPDATE_MODE", "1")
    got = adapter.run(data, **kw)
    assert np.array_equal(got.atomHistoryA, want.atomHistoryA) and np.array_equal(got.atomHistoryP, want.atomHistoryP)
    # the reference's loop counts the proposals it asked for (nA + nP), the sweep reports the proposals it made
    for f in ("Amean", "Pmean"):
        assert np.array_equal(bits(getattr(got, f)), bits(getattr(want, f))), f
