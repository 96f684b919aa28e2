"""Pins the CPU oracle (oracle/cogaps_oracle.c) to the reference.

Two sources of truth:
  * golden vectors in tests/golden/ref_golden.npz, produced by tests/golden/make_fixtures.py from the
    UNMODIFIED reference compiled in place (oracle/_ref, scalar and AVX builds) — always checked;
  * the live oracle/_ref libraries, when present (they are built by __graft_entry__.build() wherever
    /root/reference exists and travel to the GPU box) — RNG streams, lookup tables, scan probes.
Everything is compared BIT-FOR-BIT: the oracle in reduce mode "scalar"/"avx8" is the same arithmetic
as the corresponding reference build.
"""
import numpy as np
import pytest

from tests.cases import RUN_CASES, load_data

VARIANTS = [("scalar", "scalar"), ("avx", "avx8")]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def run_case(lib, name, options=None, **extra):
    case = RUN_CASES[name]
    data = load_data(case["data"])
    kw = dict(case["params"])
    kw.update(extra)
    unc = np.maximum(0.15 * data, 0.2).astype(np.float32) if case.get("uncertainty") else None
    if case.get("fixed"):
        rows = data.shape[1] if kw["whichMatrixFixed"] == "P" else data.shape[0]
        rng = np.random.default_rng(7)
        kw["fixedPatterns"] = rng.gamma(2.0, 0.5, (rows, kw["nPatterns"])).astype(np.float32)
    if options is not None:
        return lib.run(data, uncertainty=unc, snapshots=True, options=options, **kw)
    return lib.run(data, uncertainty=unc, snapshots=True, **kw)


def ref_or_skip(variant):
    from oracle.harness import RefLib
    if not RefLib.available(variant):
        pytest.skip("oracle/_ref not built here (needs /root/reference); golden vectors cover this")
    return RefLib(variant)


@pytest.mark.parametrize("variant,mode", VARIANTS)
@pytest.mark.parametrize("name", sorted(RUN_CASES))
def test_run_matches_golden(oracle, golden, name, variant, mode):
    res = run_case(oracle, name, options=oracle.options(reduce=mode))
    pre = "%s/%s/" % (variant, name)
    # integer bookkeeping: exact (test_seed_consistency.R:13-21 demands exact atom histories)
    assert np.array_equal(res.atomHistoryA, golden[pre + "atomHistoryA"])
    assert np.array_equal(res.atomHistoryP, golden[pre + "atomHistoryP"])
    sc = golden[pre + "scalars"]
    assert res.totalUpdates == int(sc[0])
    # floating point: bit-exact, same arithmetic in the same order
    for f in ("chisqHistory", "Amean", "Asd", "Pmean", "Psd"):
        assert np.array_equal(bits(getattr(res, f)), bits(golden[pre + f])), f
    assert np.float32(res.meanChiSq) == np.float32(sc[1])
    assert np.float32(res.averageQueueLengthA) == np.float32(sc[2])
    assert np.float32(res.averageQueueLengthP) == np.float32(sc[3])
    if RUN_CASES[name].get("pump"):
        assert np.array_equal(bits(res.pumpMatrix), bits(golden[pre + "pumpMatrix"]))
        assert np.array_equal(bits(res.meanPatternAssignment), bits(golden[pre + "meanPatternAssignment"]))
    if RUN_CASES[name]["params"].get("snapshotFrequency"):
        assert np.array_equal(bits(res.snapshotsA[-1]), bits(golden[pre + "snapshotsA_last"]))
        assert np.array_equal(bits(res.snapshotsP[-1]), bits(golden[pre + "snapshotsP_last"]))


@pytest.mark.parametrize("variant", ["scalar", "avx"])
def test_tables(oracle, golden, variant):
    """Random.cpp:269-295 — erf / erfinv / qgamma tables, bit-exact (Boost is restated, see oracle header)."""
    erf, erfinv, qgamma = oracle.tables()
    assert np.array_equal(bits(erf), bits(golden[variant + "/tables/erf"]))
    assert np.array_equal(bits(erfinv), bits(golden[variant + "/tables/erfinv"]))
    assert np.array_equal(bits(qgamma), bits(golden[variant + "/tables/qgamma"]))
    # cpp_tests/testRandom.cpp:54-86 accuracy bound: table vs exact within 0.03
    from math import erf as erf_exact
    x = np.arange(3001) / 1000.0
    assert np.max(np.abs(erf - np.array([erf_exact(v) for v in x]))) < 1e-6


def test_tables_live(oracle):
    ref = ref_or_skip("scalar")
    for a, b in zip(ref.tables(), oracle.tables()):
        assert np.array_equal(bits(a), bits(b))


RNG_KINDS = [
    (0, {}), (1, {}), (2, dict(a=0, b=9)), (2, dict(a=5, b=5)), (2, dict(a=0, b=4000000000)),
    (3, dict(a=1, b=18446744073709551600)), (3, dict(a=123456789, b=123456789012345)),
    (3, dict(a=7, b=7)), (4, {}), (5, dict(lam=0.5)), (5, dict(lam=4.9)), (5, dict(lam=10.0)),
    (5, dict(lam=3400.0)), (5, dict(lam=215000.0)), (6, dict(f=(0.37, 0, 0, 0))),
    (7, dict(f=(0.0, 5.0, 1.2, 0.7))), (7, dict(f=(0.0, 50.0, -3.0, 0.5))), (7, dict(f=(-1.5, 2.5, 0.2, 3.0))),
    (7, dict(f=(0.0, 50.0, 80.0, 1.0))), (8, dict(f=(3.0, 1.7, 0, 0))), (8, dict(f=(0.02, 0.4, 0, 0))),
]


@pytest.mark.parametrize("kind,kw", RNG_KINDS)
def test_rng_streams_live(oracle, kind, kw):
    """math/Random.cpp:32-260 — xoroshiro128+ seeder, PCG32, ranges, Poisson, truncated draws."""
    ref = ref_or_skip("scalar")
    for seed in (1, 42, 969, 4294967295):
        a = ref.rng_stream(seed, kind, 400, **kw)
        b = oracle.rng_stream(seed, kind, 400, **kw)
        assert np.array_equal(a, b)


def test_rng_uniform_range(oracle):
    """uniform() = u32 / float(UINT32_MAX) lies in [0, 1] inclusive (Random.cpp:63-66)."""
    u = oracle.rng_stream(42, 4, 20000).astype(np.uint32).view(np.float32)
    assert 0.0 <= u.min() and u.max() <= 1.0 and abs(u.mean() - 0.5) < 0.01


@pytest.mark.parametrize("variant,mode", VARIANTS)
def test_alpha_parameters_live(oracle, variant, mode):
    """DenseNormalModel.cpp:162-240 — the three PERFORMANCE CRITICAL scans, bit-exact per build."""
    ref = ref_or_skip(variant)
    rng = np.random.default_rng(3)
    for (g, s, k) in ((37, 23, 4), (64, 40, 3), (50, 129, 6)):
        data = rng.gamma(2.0, 1.0, (g, s)).astype(np.float32)
        data[rng.random((g, s)) < 0.2] = 0
        A = (rng.gamma(2.0, 0.5, (g, k)) * (rng.random((g, k)) < 0.6)).astype(np.float32)
        Pm = (rng.gamma(2.0, 0.5, (s, k)) * (rng.random((s, k)) < 0.6)).astype(np.float32)
        q = []
        for _ in range(60):
            r1, r2 = rng.integers(0, g, 2)
            c1, c2 = rng.integers(0, k, 2)
            v = rng.integers(0, 3)
            if v == 1 and rng.random() < 0.5:
                r2 = r1
            q.append((v, r1, c1, r2, c2, -float(rng.random())))
        for unc in (None, np.maximum(0.2 * data, 0.3).astype(np.float32)):
            s_ref, smu_ref, ap_ref = ref.alpha_parameters(data, A, Pm, q, uncertainty=unc, want_ap=True)
            s_or, smu_or, ap_or = oracle.alpha_parameters(data, A, Pm, q, uncertainty=unc, want_ap=True,
                                                         options=oracle.options(reduce=mode))
            assert np.array_equal(bits(ap_ref), bits(ap_or))
            assert np.array_equal(bits(s_ref), bits(s_or))
            assert np.array_equal(bits(smu_ref), bits(smu_or))


@pytest.mark.parametrize("variant,mode", VARIANTS)
def test_sparse_alpha_parameters_live(oracle, variant, mode):
    """SparseNormalModel.cpp:153-311 — alphaParameters(row,col), (r1,c1,r2,c2), WithChange and the Z1/Z2 tables behind
    them — and chiSq (:39-60), oracle against the compiled reference, bit-exact per build.  Includes k > 25, where
    gaps::dot switches its accumulation order (math/VectorMath.h:40-98), values below epsilon (kept by the row copy of
    HybridMatrix, dropped by the column copy) and an empty and a full data row."""
    ref = ref_or_skip(variant)
    rng = np.random.default_rng(4)
    for (g, s, k) in ((37, 23, 4), (64, 300, 7), (30, 90, 30)):
        data = rng.gamma(2.0, 1.0, (g, s)).astype(np.float32)
        data[rng.random((g, s)) < 0.8] = 0
        data[0, :] = 0
        data[1, :] = 1.5
        A = (rng.gamma(2.0, 0.5, (g, k)) * (rng.random((g, k)) < 0.6)).astype(np.float32)
        Pm = (rng.gamma(2.0, 0.5, (s, k)) * (rng.random((s, k)) < 0.6)).astype(np.float32)
        A[2, 0] = 5e-6
        Pm[3, 1] = 5e-6
        q = [(0, 0, 0, 0, 0, 0.0), (0, 1, 1, 1, 1, 0.0), (1, 1, 0, 1, 1, 0.0), (2, 1, 1, 1, 1, -0.3), (1, 2, 0, 5, 1, 0.0)]
        for _ in range(80):
            r1, r2 = rng.integers(0, g, 2)
            c1, c2 = rng.integers(0, k, 2)
            v = int(rng.integers(0, 3))
            if v == 1 and rng.random() < 0.6:
                r2 = r1
            q.append((v, r1, c1, r2, c2, -float(rng.random())))
        s_ref, smu_ref = ref.alpha_parameters_sparse(data, A, Pm, q)
        s_or, smu_or = oracle.alpha_parameters_sparse(data, A, Pm, q, options=oracle.options(reduce=mode))
        assert np.array_equal(bits(s_ref), bits(s_or))
        assert np.array_equal(bits(smu_ref), bits(smu_or))
        cs_ref, cs_or = ref.chisq_sparse(data, A, Pm), oracle.chisq_sparse(data, A, Pm) if mode == "scalar" else None
        if cs_or is not None:
            assert np.array_equal(bits(cs_ref), bits(cs_or))


def test_chisq_known_answer(oracle):
    """cpp_tests/testDenseGibbsSampler.cpp:11-35: A = P = 0, default uncertainty, data(i,j) = i+j+1 on
    25x50  =>  chiSq == 100 * nRow * nCol (S = 0.1 D so every term is exactly 100)."""
    g, s, k = 25, 50, 7
    data = (np.add.outer(np.arange(g), np.arange(s)) + 1).astype(np.float32)
    out = oracle.chisq(data, np.zeros((g, k), np.float32), np.zeros((s, k), np.float32))
    assert out[0] == pytest.approx(100.0 * g * s, rel=1e-6)
    assert out[1] == pytest.approx(100.0 * g * s, rel=1e-6)


def test_chisq_live(oracle):
    ref = ref_or_skip("scalar")
    rng = np.random.default_rng(5)
    data = rng.gamma(2.0, 1.0, (41, 29)).astype(np.float32)
    A = rng.gamma(2.0, 0.5, (41, 4)).astype(np.float32)
    Pm = rng.gamma(2.0, 0.5, (29, 4)).astype(np.float32)
    a, b = ref.chisq(data, A, Pm), oracle.chisq(data, A, Pm)
    assert np.array_equal(bits(a[:2]), bits(b[:2]))


@pytest.mark.parametrize("name", ["gist_async", "modsim_async", "syn_203x117"])
def test_run_live_and_thread_invariance(oracle, name):
    """test_seed_consistency.R:41-69 — same seed, nThreads in {1,3,6}: identical results; and the oracle
    equals that common result."""
    ref = ref_or_skip("scalar")
    base = run_case(oracle, name)
    for threads in (1, 3, 6):
        r = run_case(ref, name, maxThreads=threads)
        assert np.array_equal(r.atomHistoryA, base.atomHistoryA)
        assert np.array_equal(r.atomHistoryP, base.atomHistoryP)
        for f in ("chisqHistory", "Amean", "Asd", "Pmean", "Psd"):
            assert np.array_equal(bits(getattr(r, f)), bits(getattr(base, f))), f


def test_portable_log_is_correctly_rounded(oracle):
    """The f64-series log shared with the device: equals the correctly rounded fp32 log (f64 log rounded
    once) on every sample, hence within 1 ulp of glibc logf (which is only 0.818-ulp accurate)."""
    import ctypes
    libm = ctypes.CDLL("libm.so.6")
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.random(20000), 2.0 ** rng.uniform(-126, 0, 2000), [1.0, 0.5, 2.0 ** -149]]).astype(np.float32)
    exact = np.log(xs.astype(np.float64)).astype(np.float32)
    differs_from_libm = 0
    for x, e in zip(xs, exact):
        a, b = oracle.portable_logf(x), libm.logf(float(x))
        assert np.float32(a) == e
        if a != b:
            differs_from_libm += 1
            assert abs(a - b) <= abs(b) * 2.0 ** -23
    assert differs_from_libm < 0.02 * len(xs)
    assert oracle.portable_logf(0.0) == -np.inf
    assert oracle.portable_logf(1.0) == 0.0


def test_oracle_device_order_is_a_valid_chain(oracle):
    """Reduce mode 'device' changes only the association of two sums: the chain must stay close to the
    scalar chain over a short horizon and identical in RNG/bookkeeping until the first near-tie."""
    opts = oracle.options(reduce="device", math="portable", orderA=(256, 4, 1, 5120), orderP=(256, 4, 1, 5120))
    a = run_case(oracle, "modsim_async", options=opts)
    b = run_case(oracle, "modsim_async")
    assert a.atomHistoryA[0] == b.atomHistoryA[0]
    assert abs(float(a.chisqHistory[-1]) - float(b.chisqHistory[-1])) / float(b.chisqHistory[-1]) < 0.5


def test_dense_and_sparse_models_agree_where_their_assumptions_coincide(oracle):
    """The reference's own (disabled) consistency check, cpp_tests/testSparseGibbsSampler.cpp:41-249: on data whose
    non-zeros are >= 1 the sparse model's uncertainty (0.1 on zeros, 0.1 d elsewhere — SparseNormalModel.h:58-66) IS the
    dense model's default max(0.1 d, 0.1), so all three alphaParameters variants and chiSq must agree to fp32 rounding
    (the reference asks for 1e-3 relative on 100 x 75, half zeros, values 1..14)."""
    rng = np.random.default_rng(123)
    g, s, k = 100, 75, 6
    data = (rng.integers(1, 15, (g, s)) * (rng.random((g, s)) >= 0.5)).astype(np.float32)
    A = (rng.gamma(2.0, 0.5, (g, k)) * (rng.random((g, k)) < 0.7)).astype(np.float32)
    Pm = (rng.gamma(2.0, 0.5, (s, k)) * (rng.random((s, k)) < 0.7)).astype(np.float32)
    q = []
    for _ in range(300):
        r1, r2 = rng.integers(0, g, 2)
        c1, c2 = rng.integers(0, k, 2)
        v = int(rng.integers(0, 3))
        if v == 1 and rng.random() < 0.6:
            r2 = r1
        if v == 1 and r1 == r2 and c1 == c2:
            c2 = (c1 + 1) % k
        q.append((v, r1, c1, r2, c2, -float(rng.random())))
    s_d, smu_d = oracle.alpha_parameters(data, A, Pm, q)
    s_s, smu_s = oracle.alpha_parameters_sparse(data, A, Pm, q)
    scale = np.maximum(np.abs(s_d), 1.0)
    assert np.all(np.abs(s_d - s_s) <= 1e-3 * scale)
    # s_mu is a difference of large terms: compare against the magnitude of what is being summed
    assert np.all(np.abs(smu_d - smu_s) <= 1e-3 * np.maximum(np.abs(smu_d), scale))
    cd, cs = oracle.chisq(data, A, Pm), oracle.chisq_sparse(data, A, Pm)
    assert cs[0] == pytest.approx(cd[0], rel=1e-3) and cs[1] == pytest.approx(cd[1], rel=1e-3)


def test_fixed_matrix_behaviour_of_the_reference_suite(oracle):
    """tests/testthat/test_fixed_matrix.R: with P (or A) fixed to a previous run's result the other factor is recovered
    better than with a random fixed matrix, and the fixed factor's reported mean is all zeros (only the free factor's
    statistics are accumulated, GapsRunner.cpp:301-308)."""
    data = load_data("gist")
    kw = dict(seed=42, nPatterns=5, nIterations=100, outputFrequency=0)
    res1 = oracle.run(data, **kw)
    rng = np.random.default_rng(1)
    for which, free, fixed, rows in (("P", "Amean", "Pmean", data.shape[1]), ("A", "Pmean", "Amean", data.shape[0])):
        res2 = oracle.run(data, whichMatrixFixed=which, fixedPatterns=getattr(res1, fixed), **kw)
        res3 = oracle.run(data, whichMatrixFixed=which, fixedPatterns=rng.uniform(1, 10, (rows, 5)).astype(np.float32), **kw)
        assert abs((getattr(res1, free) - getattr(res2, free)).sum()) < abs((getattr(res1, free) - getattr(res3, free)).sum())
        assert getattr(res3, fixed).min() == 0.0 and getattr(res3, fixed).max() == 0.0
        assert res3.meanChiSq == 0.0                                    # GapsRunner.cpp:478-485


def test_reported_mean_chisq_matches_the_recomputation(oracle):
    """tests/testthat/test_chisq.R: meanChiSq == sum(((D - Amean Pmean^T) / unc)^2) to float precision times the size of
    the matrix, with an explicit uncertainty of 0.1 D."""
    data = load_data("gist")
    unc = (0.1 * data).astype(np.float32)
    assert unc.min() > 0
    res = oracle.run(data, uncertainty=unc, seed=1, nPatterns=3, nIterations=300, outputFrequency=0)
    M = res.Amean.astype(np.float64) @ res.Pmean.astype(np.float64).T
    calculated = (((data.astype(np.float64) - M) / unc.astype(np.float64)) ** 2).sum()
    assert res.meanChiSq == pytest.approx(calculated, rel=1e-7 * data.size)
