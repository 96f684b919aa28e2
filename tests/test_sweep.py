"""The row-parallel sweep (cgb_params.updateMode = CGB_UPDATE_SWEEP; cogaps_b200/csrc/sweep.cuh).

It is a different chain from the reference's for the same seed (north star: conflict-partitioned sweep with device-side
counter-based draws), so parity is defined in two steps:
  * CPU (`-m "not gpu"`): the oracle's restatement of the sweep (oracle/cogaps_oracle.c sweep_row) against the REFERENCE
    ITSELF (oracle/_ref scalar build, or the committed golden vectors of its runs) statistically — atom counts, chi-square
    and meanChiSq over seeds inside the reference's own seed-to-seed spread and its own tolerance for "same result"
    (0.1 relative, tests/testthat/test_seed_consistency.R:15-18); plus known answers for the Philox generator and
    the portable exp the sweep rests on;
  * GPU (`-m gpu`): the CUDA sweep through the C ABI against that restatement on the same seeded inputs: atom-count
    histories and update counts EXACT, chi-square rtol 1e-4, posterior means / sds rtol 1e-4 (north star tolerance; in
    practice bit-identical), and size-independent invariants at the BASELINE size.
"""
import ctypes as C
import os

import numpy as np
import pytest

from tests.cases import RUN_CASES, load_data

RTOL_MEANS = 1e-4
RTOL_CHISQ = 1e-4
SWEEP = 1  # CGB_UPDATE_SWEEP


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_close(a, b, rtol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max() / scale
    assert err <= rtol, "%s: max error %.3g of scale (tolerance %.1g)" % (what, err, rtol)


# ------------------------------------------------------------------------------------------------
# CPU: the restatement itself
# ------------------------------------------------------------------------------------------------
def test_philox_known_answers(oracle):
    """Philox4x32-10 against the known-answer vectors published with Random123 (kat_vectors: philox4x32 10)."""
    kats = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kats:
        out = (C.c_uint32 * 4)()
        oracle.lib.cogaps_oracle_philox((C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), out)
        assert tuple(out) == want


def test_portable_exp_is_exp(oracle):
    """the exp behind the same-bin exchange's truncGammaUpper (math/Random.cpp:194-200): within 1 ulp of the correctly
    rounded fp32 exp over the range the sweep uses it on (arguments -b/scale <= 0)"""
    oracle.lib.cogaps_oracle_portable_expf.restype = C.c_float
    oracle.lib.cogaps_oracle_portable_expf.argtypes = [C.c_float]
    rng = np.random.default_rng(3)
    xs = np.concatenate([-rng.random(20000) * 100.0, rng.random(2000) * 80.0, [0.0, -0.0, -1e-8, -87.0, -103.9]]).astype(np.float32)
    got = np.array([oracle.lib.cogaps_oracle_portable_expf(float(x)) for x in xs], np.float32)
    want = np.exp(xs.astype(np.float64)).astype(np.float32)
    ulp = np.spacing(np.maximum(np.abs(want), np.float32(1e-37)))
    assert (np.abs(got.astype(np.float64) - want.astype(np.float64)) <= ulp).all()


def test_sweep_oracle_is_deterministic_and_seed_dependent(oracle):
    data = load_data("gist")
    kw = dict(seed=42, nPatterns=5, nIterations=40, outputFrequency=10, updateMode=SWEEP)
    a = oracle.run(data, **kw)
    b = oracle.run(data, **kw)
    c = oracle.run(data, **dict(kw, seed=45))   # 43 == 42 | 1: the seeder ORs 1 in
    assert np.array_equal(a.atomHistoryA, b.atomHistoryA) and np.array_equal(bits(a.Amean), bits(b.Amean))
    assert a.totalUpdates == b.totalUpdates
    assert not np.array_equal(a.atomHistoryA, c.atomHistoryA)


def _stat_rows(run, data, k, its, seeds, **kw):
    rows = []
    for seed in seeds:
        r = run(data, seed=seed, nPatterns=k, nIterations=its, outputFrequency=its // 4, **kw)
        rows.append((r.atomHistoryA[-1], r.atomHistoryP[-1], r.chisqHistory[-1], r.meanChiSq))
    return np.array(rows, np.float64)


@pytest.mark.parametrize("name,k,its", [("gist", 7, 400), ("modsim", 3, 1500)])
def test_sweep_agrees_statistically_with_the_reference(oracle, name, k, its):
    """Tier 3 of SURVEY 7.4-2 for the sweep: over seeds, the sweep's atom counts, final chi-square and meanChiSq sit inside
    the reference's own seed-to-seed spread: |difference of means| <= 3 standard errors, and never more than the larger of
    10 % (what the reference calls "the same result") and one seed-to-seed standard deviation."""
    from oracle.harness import RefLib
    if not RefLib.available("scalar"):
        pytest.skip("oracle/_ref not built here (needs /root/reference); the committed goldens cover the exact path")
    data = load_data(name)
    seeds = list(range(1, 16, 2))     # the seeder ORs 1 into the seed (Random.cpp:221-229): odd seeds are distinct chains
    ref = _stat_rows(RefLib("scalar").run, data, k, its, seeds)
    swp = _stat_rows(oracle.run, data, k, its, seeds, updateMode=SWEEP)
    for j, what in enumerate(("atoms A", "atoms P", "chi-square", "meanChiSq")):
        mr, ms = ref[:, j].mean(), swp[:, j].mean()
        se = np.sqrt((ref[:, j].var(ddof=1) + swp[:, j].var(ddof=1)) / len(seeds))
        assert abs(mr - ms) <= max(3.0 * se, 0.02 * abs(mr)), "%s: reference %.1f vs sweep %.1f (se %.1f)" % (what, mr, ms, se)
        sd = np.sqrt(0.5 * (ref[:, j].var(ddof=1) + swp[:, j].var(ddof=1)))
        tol = 0.15 if what.startswith("atoms") else 0.1          # atom counts: the bar round 1 set for GPU vs reference
        assert abs(mr - ms) <= max(tol * abs(mr), sd), what      # and inside 10-15 % or one seed-to-seed standard deviation


def test_sweep_reconstruction_matches_the_exact_chain(oracle):
    """On a noisy low-rank matrix the two chains must find the same fit: A*P^T of the posterior means agree far inside
    the noise, and the factors hold the same total mass."""
    data = load_data("syn:120:90:4:21")
    kw = dict(seed=11, nPatterns=4, nIterations=600, outputFrequency=200)
    e = oracle.run(data, **kw)
    s = oracle.run(data, updateMode=SWEEP, **kw)
    re_, rs = e.Amean @ e.Pmean.T, s.Amean @ s.Pmean.T
    noise = np.abs(data - re_).mean()
    assert np.abs(re_ - rs).mean() <= 0.5 * noise
    assert s.chisqHistory[-1] == pytest.approx(e.chisqHistory[-1], rel=0.05)
    assert int(s.atomHistoryA[-1]) == pytest.approx(int(e.atomHistoryA[-1]), rel=0.15)


@pytest.mark.parametrize("name,k,its", [("gist", 7, 400), ("spz:120:90:4:3:80", 4, 400)])
def test_sparse_sweep_agrees_statistically_with_the_reference(oracle, name, k, its):
    """the same Tier-3 comparison for the sweep over the sparse model, against the reference's own SparseNormalModel chain"""
    from oracle.harness import RefLib
    if not RefLib.available("scalar"):
        pytest.skip("oracle/_ref not built here (needs /root/reference)")
    data = load_data(name)
    seeds = list(range(1, 12, 2))
    ref = _stat_rows(RefLib("scalar").run, data, k, its, seeds, useSparseOptimization=1)
    swp = _stat_rows(oracle.run, data, k, its, seeds, updateMode=SWEEP, useSparseOptimization=1)
    for j, what in enumerate(("atoms A", "atoms P", "chi-square", "meanChiSq")):
        mr, ms = ref[:, j].mean(), swp[:, j].mean()
        sd = np.sqrt(0.5 * (ref[:, j].var(ddof=1) + swp[:, j].var(ddof=1)))
        se = sd * np.sqrt(2.0 / len(seeds))
        if j >= 2:
            assert abs(mr - ms) <= max(3.0 * se, 0.03 * abs(mr)), "%s: reference %.1f vs sweep %.1f (se %.1f)" % (what, mr, ms, se)
        assert abs(mr - ms) <= max((0.2 if j < 2 else 0.1) * abs(mr), sd), what


def test_sweep_keeps_mass_and_atoms_in_step(oracle):
    """changeMatrix / safelyChangeMatrix bookkeeping of the sweep: snapshots are non-negative and zero exactly where a
    fixed matrix must not move"""
    data = load_data("gist")
    rows = data.shape[1]
    fixed = np.random.default_rng(7).gamma(2.0, 0.5, (rows, 3)).astype(np.float32)
    r = oracle.run(data, seed=5, nPatterns=3, nIterations=60, outputFrequency=20, updateMode=SWEEP,
                   whichMatrixFixed="P", fixedPatterns=fixed, snapshots=True, snapshotFrequency=30)
    assert np.array_equal(bits(r.snapshotsP[-1]), bits(fixed))      # test_fixed_matrix.R: the fixed matrix is untouched
    assert (r.snapshotsA[-1] >= 0).all() and r.atomHistoryP[-1] == 0


# ------------------------------------------------------------------------------------------------
# GPU: the CUDA sweep against the restatement
# ------------------------------------------------------------------------------------------------
def sweep_options(oracle, nGenes, nSamples):
    from cogaps_b200.sampler import sweep_reduction_order_for_length
    return oracle.options(reduce="device", math="portable", orderA=sweep_reduction_order_for_length(nSamples),
                          orderP=sweep_reduction_order_for_length(nGenes))


def case_inputs(name, **extra):
    case = RUN_CASES[name]
    data = load_data(case["data"])
    kw = dict(case["params"])
    kw.update(extra)
    unc = np.maximum(0.15 * data, 0.2).astype(np.float32) if case.get("uncertainty") else None
    if case.get("fixed"):
        rows = data.shape[1] if kw["whichMatrixFixed"] == "P" else data.shape[0]
        kw["fixedPatterns"] = np.random.default_rng(7).gamma(2.0, 0.5, (rows, kw["nPatterns"])).astype(np.float32)
    return data, unc, kw


def check_run(got, want, kw):
    assert np.array_equal(got.atomHistoryA, want.atomHistoryA), (got.atomHistoryA, want.atomHistoryA)
    assert np.array_equal(got.atomHistoryP, want.atomHistoryP), (got.atomHistoryP, want.atomHistoryP)
    assert got.totalUpdates == want.totalUpdates
    assert_close(got.chisqHistory, want.chisqHistory, RTOL_CHISQ, "chisqHistory")
    for f in ("Amean", "Asd", "Pmean", "Psd"):
        assert_close(getattr(got, f), getattr(want, f), RTOL_MEANS, f)
    if kw.get("whichMatrixFixed", "N") == "N":
        assert got.meanChiSq == pytest.approx(want.meanChiSq, rel=RTOL_CHISQ)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["modsim_async", "gist_async", "gist_transposed", "gist_uncertainty", "gist_pump",
                                  "gist_fixedP", "gist_fixedA", "gist_subset_genes", "gist_subset_samples",
                                  "syn_203x117", "syn_sparse"])
def test_sweep_run_matches_oracle(oracle, name):
    """gaps::run in sweep mode: same seed, same data -> the chain the restatement computes, snapshot bits included"""
    import cogaps_b200 as cg
    data, unc, kw = case_inputs(name, updateMode=SWEEP)
    g, s = (data.shape[1], data.shape[0]) if kw.get("transposeData") else data.shape
    if kw.get("subsetIndices") is not None:
        g, s = (len(kw["subsetIndices"]), s) if kw.get("subsetGenes") else (g, len(kw["subsetIndices"]))
    want = oracle.run(data, uncertainty=unc, snapshots=True, options=sweep_options(oracle, g, s), **kw)
    got = cg.gaps_run(data, uncertainty=unc, snapshots=True, **kw)
    check_run(got, want, kw)
    if kw.get("snapshotFrequency"):
        assert np.array_equal(bits(got.snapshotsA[-1]), bits(want.snapshotsA[-1]))
        assert np.array_equal(bits(got.snapshotsP[-1]), bits(want.snapshotsP[-1]))
    if RUN_CASES[name].get("pump"):
        assert np.array_equal(got.pumpMatrix, want.pumpMatrix)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sparse_gist", "sparse_modsim", "sparse_120x90", "sparse_k30", "sparse_gist_fixedP"])
def test_sparse_sweep_run_matches_oracle(oracle, name):
    """the sweep over the SparseNormalModel (sweep_sparse_kernel + the sparse transport): same chain as the oracle's sweep
    on its sparse model — no AP line here, the row's factor row in both copies lives in shared memory instead"""
    import cogaps_b200 as cg
    data, unc, kw = case_inputs(name, updateMode=SWEEP)
    want = oracle.run(data, snapshots=True, options=sweep_options(oracle, data.shape[0], data.shape[1]), **kw)
    got = cg.gaps_run(data, snapshots=True, **kw)
    check_run(got, want, kw)
    if kw.get("snapshotFrequency"):
        assert np.array_equal(bits(got.snapshotsA[-1]), bits(want.snapshotsA[-1]))
        assert np.array_equal(bits(got.snapshotsP[-1]), bits(want.snapshotsP[-1]))


@pytest.mark.gpu
@pytest.mark.parametrize("spec,k", [("spz:3000:30000:8:7:95", 50), ("spz:26000:2500:8:9:95", 50)])
def test_sparse_sweep_c4_shaped_rows(oracle, spec, k):
    """BASELINE configs[3]'s row shape in sweep mode: 30000-long rows at 95 % zeros (more than one 1024-entry group per
    scan), nPatterns = 50, both orientations"""
    import cogaps_b200 as cg
    data = load_data(spec)
    kw = dict(seed=17, nPatterns=k, nIterations=8, outputFrequency=4, maxThreads=1, useSparseOptimization=1, snapshotFrequency=8,
              updateMode=SWEEP)
    want = oracle.run(data, snapshots=True, options=sweep_options(oracle, data.shape[0], data.shape[1]), **kw)
    got = cg.gaps_run(data, snapshots=True, **kw)
    assert np.array_equal(got.atomHistoryA, want.atomHistoryA), (got.atomHistoryA, want.atomHistoryA)
    assert np.array_equal(got.atomHistoryP, want.atomHistoryP)
    assert got.totalUpdates == want.totalUpdates
    assert np.array_equal(bits(got.snapshotsA[-1]), bits(want.snapshotsA[-1]))
    assert np.array_equal(bits(got.snapshotsP[-1]), bits(want.snapshotsP[-1]))


@pytest.mark.gpu
@pytest.mark.parametrize("spec,k,its", [("syn:60:5000:4:3", 4, 30), ("syn:30:20000:3:5", 3, 20), ("syn:12:70000:3:9", 3, 12)])
def test_sweep_long_rows(oracle, spec, k, its):
    """row lengths of the BASELINE shapes (5000: 256 threads per row, 20000: 512, both staged in shared memory) and one
    that no longer fits there (70000 floats x 2 lines > 227 KB: the row is scanned and committed through L2)"""
    import cogaps_b200 as cg
    data = load_data(spec)
    kw = dict(seed=5, nPatterns=k, nIterations=its, outputFrequency=5, maxThreads=1, updateMode=SWEEP)
    want = oracle.run(data, options=sweep_options(oracle, data.shape[0], data.shape[1]), **kw)
    got = cg.gaps_run(data, **kw)
    check_run(got, want, kw)


@pytest.mark.gpu
def test_sweep_rows_in_shared_memory_or_not_give_the_same_chain(monkeypatch):
    """what a row keeps in shared memory (D and AP lines / the AP line only / nothing) and the order rows are handed to
    CTAs in (longest chain first / by index) are scheduling choices: every output bit is the same"""
    import cogaps_b200 as cg
    data = load_data("syn:203:117:5:11")
    kw = dict(seed=123, nPatterns=5, nIterations=60, outputFrequency=10, updateMode=SWEEP)
    a = cg.gaps_run(data, snapshots=True, snapshotFrequency=20, **kw)
    for knob, value in (("COGAPS_SWEEP_STAGE", "0"), ("COGAPS_SWEEP_STAGE", "1"), ("COGAPS_SWEEP_STAGE", "2"), ("COGAPS_SWEEP_ORDER", "0")):
        monkeypatch.setenv(knob, value)
        b = cg.gaps_run(data, snapshots=True, snapshotFrequency=20, **kw)
        monkeypatch.delenv(knob)
        assert np.array_equal(a.atomHistoryA, b.atomHistoryA) and np.array_equal(a.atomHistoryP, b.atomHistoryP), (knob, value)
        assert np.array_equal(bits(a.snapshotsA), bits(b.snapshotsA)) and np.array_equal(bits(a.snapshotsP), bits(b.snapshotsP)), (knob, value)
        assert np.array_equal(bits(a.Amean), bits(b.Amean)), (knob, value)


@pytest.mark.gpu
def test_switching_modes_keeps_the_atoms():
    """exact -> sweep -> exact through cgb_sampler_set_update_mode: the atoms (position, mass) survive both conversions,
    each mode's update runs from the other's state, and every factor element stays the sum of the masses in its bin"""
    import bench
    data = load_data("gist")
    k = 5
    pair = bench.Chain(data, k, 9)
    for _ in range(15):
        pair.step()
    before = {}
    for name, smp in (("A", pair.A), ("P", pair.P)):
        pos, mass = smp.atoms()
        order = np.argsort(pos)
        before[name] = (pos[order], mass[order])
        smp.setUpdateMode(SWEEP)
        pos2, mass2 = smp.atoms()
        assert np.array_equal(pos2, before[name][0]) and np.array_equal(bits(mass2), bits(before[name][1]))
        assert smp.nAtoms() == pos.size
    for _ in range(10):
        pair.step()
    for smp in (pair.A, pair.P):
        smp.setUpdateMode(0)
    for _ in range(10):
        pair.step()
    for smp in (pair.A, pair.P):
        M = smp.getMatrix()
        pos, mass = smp.atoms()
        nbins = M.shape[0] * k
        binlen = np.uint64(0xFFFFFFFFFFFFFFFF // nbins)
        b = np.minimum((pos // binlen).astype(np.int64), nbins - 1)
        summed = np.bincount(b, weights=mass.astype(np.float64), minlength=nbins).reshape(M.shape[0], k)
        assert np.allclose(summed, M.astype(np.float64), rtol=1e-4, atol=1e-5)
        assert smp.nAtoms() == pos.size and pos.size > 0


@pytest.mark.gpu
def test_sweep_full_size_invariants_20000x5000_k20():
    """BASELINE.json configs[2] at full size in sweep mode, through properties that need no oracle run: the AP kept by
    in-shared-memory rank-one commits equals A * P^T rebuilt from the factors (chi-square before / after
    extraInitialization, A side == P side), every factor element is the sum of the atom masses of its bin, atoms of a
    row stay inside the row's segment in ascending order, and the reported atom total is the store's."""
    import bench
    g, s, k = 20000, 5000, 20
    data = bench.make_data(g, s, k)
    chain = bench.Chain(data, k, 42, updateMode=SWEEP)
    iters = 30
    for i in range(iters):
        temp = min(1.0, 2.0 * i / iters)
        chain.A.setAnnealingTemp(temp)
        chain.P.setAnnealingTemp(temp)
        chain.step()
    assert chain.A.nAtoms() > 10000 and chain.P.nAtoms() > 2000
    csA, csP = chain.A.chiSq(), chain.P.chiSq()
    assert csA == pytest.approx(csP, rel=RTOL_CHISQ)
    for smp in (chain.A, chain.P):
        M = smp.getMatrix()
        assert (M >= 0).all()
        pos, mass = smp.atoms()
        assert pos.size == smp.nAtoms()
        assert (np.diff(pos.astype(np.float64)) > 0).all()          # ascending over rows and within rows
        nbins = M.shape[0] * k
        binlen = np.uint64(0xFFFFFFFFFFFFFFFF // nbins)
        b = np.minimum((pos // binlen).astype(np.int64), nbins - 1)
        summed = np.bincount(b, weights=mass.astype(np.float64), minlength=nbins).reshape(M.shape[0], k)
        assert np.allclose(summed, M.astype(np.float64), rtol=1e-4, atol=1e-5)
    chain.A.extraInitialization()
    chain.P.extraInitialization()
    assert chain.A.chiSq() == pytest.approx(csA, rel=RTOL_CHISQ)
    assert chain.P.chiSq() == pytest.approx(csP, rel=RTOL_CHISQ)
