"""Shared test-case definitions: the reference's own fixtures (modsimdata 25x20, GIST 1363x9) and
seeded synthetic matrices, with the parameter sets mirrored from the reference's testthat suite
(tests/testthat/test_seed_consistency.R, test_fixed_matrix.R, test_top_level.R)."""
import os
import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def synthetic(spec):
    """'syn:G:S:k:seed' -> SURVEY 8(d) recipe: noisy non-negative low-rank matrix, max < 50."""
    _, g, s, k, seed = spec.split(":")
    g, s, k, seed = int(g), int(s), int(k), int(seed)
    rng = np.random.default_rng(seed)
    a0 = rng.gamma(2.0, 0.5, (g, k)) * (rng.random((g, k)) < 0.3)
    p0 = rng.gamma(2.0, 0.5, (s, k)) * (rng.random((s, k)) < 0.3)
    m = a0 @ p0.T
    d = np.maximum(m * (1 + 0.1 * rng.standard_normal(m.shape)), 0)
    d *= 40.0 / max(d.max(), 1e-9)
    return d.astype(np.float32)


def synthetic_sparse(spec):
    """'spz:G:S:k:seed:zeroPercent' -> the synthetic recipe with a fraction of entries zeroed uniformly at
    random (single-cell-like matrices for the sparse model)."""
    _, g, s, k, seed, z = spec.split(":")
    d = synthetic("syn:%s:%s:%s:%s" % (g, s, k, seed))
    rng = np.random.default_rng(int(seed) + 1)
    d[rng.random(d.shape) < int(z) / 100.0] = 0
    return d


def load_data(name):
    if name.startswith("spz:"):
        return synthetic_sparse(name)
    if name.startswith("syn:"):
        return synthetic(name)
    return np.load(os.path.join(GOLDEN, name + ".npy"))


def P(**kw):
    base = dict(seed=42, nPatterns=3, nIterations=100, outputFrequency=10, maxThreads=1)
    base.update(kw)
    return base


# name -> {data, params, flags}
RUN_CASES = {
    # BASELINE.json configs[0]: modsimdata 25x20, nPatterns=3, seed=42, single-thread sequential sampler
    "modsim_seq": dict(data="modsim", params=P(asynchronousUpdates=0, nIterations=500, outputFrequency=50)),
    "modsim_async": dict(data="modsim", params=P(nIterations=500, outputFrequency=50, snapshotFrequency=100)),
    # BASELINE.json configs[1]: GIST 1363x9, nPatterns=7 (test_seed_consistency.R:41-69 uses seed 42, 100 it)
    "gist_async": dict(data="gist", params=P(nPatterns=7, nIterations=100, outputFrequency=10)),
    "gist_seq": dict(data="gist", params=P(nPatterns=7, nIterations=60, outputFrequency=10, asynchronousUpdates=0)),
    "gist_transposed": dict(data="gist", params=P(nPatterns=4, nIterations=60, transposeData=1)),
    "gist_uncertainty": dict(data="gist", params=P(nPatterns=5, nIterations=60), uncertainty=True),
    "gist_pump": dict(data="gist", params=P(nPatterns=5, nIterations=60, takePumpSamples=1), pump=True),
    # test_fixed_matrix.R
    "gist_fixedP": dict(data="gist", params=P(nPatterns=3, nIterations=80, whichMatrixFixed="P"), fixed=True),
    "gist_fixedA": dict(data="gist", params=P(nPatterns=3, nIterations=80, whichMatrixFixed="A"), fixed=True),
    # test_subset_data.R: explicit subsets of genes / samples (1-based indices)
    "gist_subset_genes": dict(data="gist", params=P(nPatterns=3, nIterations=60, subsetGenes=1,
                                                    subsetIndices=list(range(5, 1300, 7)))),
    "gist_subset_samples": dict(data="gist", params=P(nPatterns=3, nIterations=60, subsetGenes=0,
                                                      subsetIndices=[9, 2, 4, 5, 7])),
    # a shape where both row lengths exceed one SIMD/warp width and are not multiples of 8
    "syn_203x117": dict(data="syn:203:117:5:11", params=P(nPatterns=5, nIterations=60, seed=123)),
    # SparseNormalModel (sparseOptimization=TRUE; test_seed_consistency.R:55-69 runs the same checks on it)
    "sparse_gist": dict(data="gist", params=P(nPatterns=7, nIterations=100, outputFrequency=10, useSparseOptimization=1)),
    "sparse_modsim": dict(data="modsim", params=P(nIterations=300, outputFrequency=50, useSparseOptimization=1,
                                                   snapshotFrequency=100)),
    "sparse_120x90": dict(data="spz:120:90:4:3:80", params=P(nPatterns=4, nIterations=80, seed=7, useSparseOptimization=1)),
    # k > 25: gaps::dot leaves its fall-through switch and accumulates forwards (VectorMath.h:40-98)
    "sparse_k30": dict(data="spz:60:200:5:9:90", params=P(nPatterns=30, nIterations=40, seed=9, useSparseOptimization=1)),
    "sparse_gist_seq": dict(data="gist", params=P(nPatterns=5, nIterations=50, useSparseOptimization=1,
                                                  asynchronousUpdates=0)),
    "sparse_gist_fixedP": dict(data="gist", params=P(nPatterns=3, nIterations=60, whichMatrixFixed="P",
                                                     useSparseOptimization=1), fixed=True),
    "syn_sparse": dict(data="syn:90:70:4:5", params=P(nPatterns=4, nIterations=80, seed=9, alphaA=0.05, alphaP=0.02,
                                                      maxGibbsMassA=50.0, maxGibbsMassP=75.0)),
}
