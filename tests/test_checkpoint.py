"""Checkpoints: the reference's Archive wire format (utils/Archive.h:16-87) and its
createCheckpoint / processCheckpoint semantics (GapsRunner.cpp:99-105,224-270).

Chain of evidence:
  reference-written files (tests/golden/ref_checkpoint_*.bin, and live oracle/_ref where present)
    == files the oracle writes for the same run                                    byte for byte   (CPU)
    -> the oracle resumes from them exactly as the reference does                  bit for bit     (CPU)
    -> the library's reader/writer reproduces them through its own images          byte for byte   (CPU)
    -> cgb_run_ex on the GPU writes the file the oracle writes in device order     byte for byte   (GPU)
    -> a GPU run resumed from a file equals the uninterrupted GPU run              bit for bit     (GPU)
tests/checkpoint_format.py is an independent reader of the format used to look inside the files.
"""
import os

import numpy as np
import pytest

from tests.cases import load_data
from tests.checkpoint_format import parse_file
from tests.golden.make_checkpoint_fixtures import CHECKPOINT_CASES, FIELDS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STATS = ("Amean", "Asd", "Pmean", "Psd")


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def read(path):
    with open(path, "rb") as f:
        return f.read()


@pytest.fixture(scope="module")
def ckgolden():
    return np.load(os.path.join(GOLDEN, "ref_checkpoint_golden.npz"))


def golden_file(name):
    return os.path.join(GOLDEN, "ref_checkpoint_%s.bin" % name)


def assert_same_result(res, golden, prefix):
    for f in FIELDS:
        want = golden[prefix + f]
        got = getattr(res, f)
        if want.dtype == np.float32:
            assert np.array_equal(bits(got), bits(want)), f
        else:
            assert np.array_equal(got, want), f
    sc = golden[prefix + "scalars"]
    assert res.totalUpdates == int(sc[0])
    assert np.float32(res.meanChiSq) == np.float32(sc[1])
    assert np.float32(res.averageQueueLengthA) == np.float32(sc[2])
    assert np.float32(res.averageQueueLengthP) == np.float32(sc[3])


# ------------------------------------------------------------------------------------------------
# CPU: the oracle and the library's host-side reader / writer against reference-written files
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", sorted(CHECKPOINT_CASES))
def test_oracle_writes_the_reference_file_and_resumes_like_it(oracle, ckgolden, name, tmp_path):
    dataset, interval, kw = CHECKPOINT_CASES[name]
    data = load_data(dataset)
    out = tmp_path / "oracle.out"
    full = oracle.run(data, options=oracle.options(checkpointInterval=interval, checkpointOutFile=out), **kw)
    assert_same_result(full, ckgolden, name + "/full/")
    assert read(out) == read(golden_file(name))
    assert not os.path.exists(str(out) + ".backup")                 # GapsRunner.cpp:243
    resumed = oracle.run(data, options=oracle.options(checkpointInFile=golden_file(name),
                                                      checkpointOutFile=tmp_path / "again.out"), **kw)
    assert_same_result(resumed, ckgolden, name + "/resumed/")
    # what a checkpoint is for: the statistics of the resumed run are those of the uninterrupted one
    for f in STATS:
        assert np.array_equal(bits(getattr(resumed, f)), bits(getattr(full, f))), f
    assert np.float32(resumed.meanChiSq) == np.float32(full.meanChiSq)
    # what it does not restore (the reference does not archive them): history and update counts restart
    c = parse_file(golden_file(name))
    assert c["phase"] == 2
    assert len(resumed.chisqHistory) == kw["nIterations"] // kw["outputFrequency"] - c["iter"] // kw["outputFrequency"]
    assert resumed.totalUpdates < full.totalUpdates


def test_taking_checkpoints_changes_the_chain_as_in_the_reference(oracle, ckgolden):
    """createCheckpoint ends with extraInitialization (GapsRunner.cpp:246-255): AP is rebuilt from the factors, which
    is not bit-identical to the incrementally updated AP, so a run with checkpoints is a (slightly) different chain."""
    dataset, interval, kw = CHECKPOINT_CASES["dense"]
    plain = oracle.run(load_data(dataset), **kw)
    assert not np.array_equal(bits(plain.Amean), bits(ckgolden["dense/full/Amean"]))
    np.testing.assert_allclose(plain.atomHistoryA[:2], ckgolden["dense/full/atomHistoryA"][:2])  # same until the first one


@pytest.mark.parametrize("name", sorted(CHECKPOINT_CASES))
def test_library_reads_and_rewrites_reference_files_byte_for_byte(name, tmp_path):
    import cogaps_b200 as cg
    src = golden_file(name)
    dst = tmp_path / "copy.out"
    cg.checkpoint_rewrite(src, dst)
    assert read(dst) == read(src)
    info = cg.checkpoint_info(src)
    c = parse_file(src)
    for f in ("seed", "nGenes", "nSamples", "nPatterns", "nIterations", "checkpointInterval"):
        assert info[f] == c["params"][f], f
    for f in ("alphaA", "alphaP", "maxGibbsMassA", "maxGibbsMassP"):
        assert np.float32(info[f]) == np.float32(c["params"][f]), f
    assert bool(info["useSparseOptimization"]) == c["params"]["useSparseOptimization"]
    assert (info["phase"], info["iter"]) == (c["phase"], c["iter"])
    assert (info["nAtomsA"], info["nAtomsP"]) == (c["A"]["pos"].size, c["P"]["pos"].size)
    assert info["statUpdates"] == c["statUpdates"]
    assert info["fileBytes"] == os.path.getsize(src)
    dataset, interval, kw = CHECKPOINT_CASES[name]
    assert (info["seed"], info["nPatterns"], info["checkpointInterval"]) == (kw["seed"], kw["nPatterns"], interval)


def test_library_rejects_damaged_files(tmp_path):
    import cogaps_b200 as cg
    raw = read(golden_file("sparse"))
    bad = tmp_path / "bad.out"
    for cut in (0, 3, 4, 30, 60, len(raw) // 3, len(raw) // 2, len(raw) - 1):
        bad.write_bytes(raw[:cut])
        with pytest.raises(cg.CogapsError) as e:
            cg.checkpoint_info(bad)
        assert e.value.code == -1
    bad.write_bytes(raw + b"\0")
    with pytest.raises(cg.CogapsError, match="trailing bytes"):
        cg.checkpoint_info(bad)
    flipped = bytearray(raw)
    flipped[1] ^= 0x40                                   # magic number (Archive.h:16,33-36)
    bad.write_bytes(bytes(flipped))
    with pytest.raises(cg.CogapsError, match="incompatible checkpoint file"):
        cg.checkpoint_info(bad)
    huge = bytearray(raw)
    huge[4 + 41 + 16:4 + 41 + 20] = (0xFFFFFFF0).to_bytes(4, "little")   # nRows of the A matrix
    bad.write_bytes(bytes(huge))
    with pytest.raises(cg.CogapsError, match="does not fit the file"):
        cg.checkpoint_info(bad)
    both = bytearray(raw)                                # rows x columns that would wrap a 64-bit product
    both[4 + 41 + 16:4 + 41 + 24] = (0xFFFFFFFF).to_bytes(4, "little") * 2
    bad.write_bytes(bytes(both))
    with pytest.raises(cg.CogapsError, match="does not fit the file"):
        cg.checkpoint_info(bad)
    with pytest.raises(cg.CogapsError, match="cannot open"):
        cg.checkpoint_info(tmp_path / "missing.out")


def test_sequential_sampler_has_no_checkpoints(oracle):
    """The reference's SingleThreadedGibbsSampler does not archive its rng and cannot read its own archive
    (SingleThreadedGibbsSampler.h:260-273); the oracle and the library refuse instead of inventing a format."""
    import cogaps_b200 as cg
    data = load_data("modsim")
    with pytest.raises(RuntimeError, match="-5"):
        oracle.run(data, options=oracle.options(checkpointInterval=10), seed=1, nPatterns=3, nIterations=20,
                   asynchronousUpdates=0)
    with pytest.raises(cg.CogapsError) as e:
        cg.gaps_run(data, checkpointInterval=10, seed=1, nPatterns=3, nIterations=20, asynchronousUpdates=0)
    assert e.value.code == -5                            # before any device is touched
    with pytest.raises(ValueError, match="asynchronousUpdates"):
        cg.CoGAPS(data, nPatterns=3, checkpointInterval=10, asynchronousUpdates=False, messages=False)


def test_user_api_checkpoint_arguments_are_validated_like_the_reference():
    """CoGAPS(checkpointInterval=, checkpointOutFile=, checkpointInFile=) — R/CoGAPS.R:90-156; distributed runs refuse
    checkpoints (R/HelperFunctions.R:237-238)."""
    import cogaps_b200 as cg
    data = load_data("modsim")
    params = cg.CogapsParams(nPatterns=3, distributed="genome-wide")
    with pytest.raises(ValueError, match="distributed"):
        cg.CoGAPS(data, params, checkpointInFile=golden_file("dense"), messages=False)
    with pytest.raises(ValueError, match="distributed"):
        cg.CoGAPS(data, params, checkpointInterval=100, messages=False)


def test_resuming_with_other_npatterns_is_refused(tmp_path):
    import cogaps_b200 as cg
    with pytest.raises(cg.CogapsError, match="nPatterns differs"):
        cg.gaps_run(load_data("modsim"), checkpointInFile=golden_file("dense"), seed=42, nPatterns=5, nIterations=60)


# live reference (oracle/_ref travels to the GPU box; absent on a bare checkout)
def ref_or_skip():
    from oracle.harness import RefLib
    if not RefLib.available("scalar"):
        pytest.skip("oracle/_ref not built here (needs /root/reference); the golden files cover this")
    return RefLib("scalar")


LIVE_CASES = {
    "gist_dense": ("gist", 15, dict(seed=42, nPatterns=3, nIterations=40, outputFrequency=10)),
    "gist_sparse": ("gist", 7, dict(seed=11, nPatterns=5, nIterations=30, outputFrequency=10, useSparseOptimization=1)),
    "gist_transposed_unc": ("gist", 9, dict(seed=3, nPatterns=4, nIterations=30, outputFrequency=5, transposeData=1)),
    "syn_small": ("syn:60:45:3:5", 10, dict(seed=5, nPatterns=3, nIterations=30, outputFrequency=10)),
    # processFixedMatrix runs before processCheckpoint (GapsRunner.cpp:410,440): the archived matrix wins
    "gist_fixedP": ("gist", 10, dict(seed=8, nPatterns=3, nIterations=30, outputFrequency=10, whichMatrixFixed="P",
                                     fixedPatterns=np.random.default_rng(7).gamma(2.0, 0.5, (9, 3)).astype(np.float32))),
}


@pytest.mark.parametrize("name", sorted(LIVE_CASES))
def test_oracle_and_live_reference_agree_on_files_and_resumes(oracle, name, tmp_path):
    ref = ref_or_skip()
    dataset, interval, kw = LIVE_CASES[name]
    data = load_data(dataset)
    unc = np.maximum(0.15 * data, 0.2).astype(np.float32) if name.endswith("_unc") else None
    rfile, ofile = tmp_path / "ref.out", tmp_path / "oracle.out"
    rfull = ref.run(data, uncertainty=unc, checkpointInterval=interval, checkpointOutFile=rfile, **kw)
    ofull = oracle.run(data, uncertainty=unc, options=oracle.options(checkpointInterval=interval, checkpointOutFile=ofile), **kw)
    assert read(rfile) == read(ofile)
    for f in FIELDS:
        assert np.array_equal(getattr(rfull, f), getattr(ofull, f)), f
    # an equilibration-phase file: the oracle stops right after its first checkpoint (the reference has no such knob)
    eq = tmp_path / "eq.out"
    assert oracle.run(data, uncertainty=unc, options=oracle.options(checkpointInterval=interval, checkpointOutFile=eq,
                                                                    stopAfterCheckpoints=1), **kw) is None
    c = parse_file(eq)
    assert (c["phase"], c["iter"]) == (1, interval - 1)
    for src in (rfile, eq):
        rres = ref.run(data, uncertainty=unc, checkpointInFile=src, checkpointOutFile=tmp_path / "r2.out", **kw)
        ores = oracle.run(data, uncertainty=unc, options=oracle.options(checkpointInFile=src, checkpointOutFile=tmp_path / "o2.out"), **kw)
        for f in FIELDS:
            assert np.array_equal(getattr(rres, f), getattr(ores, f)), f
        assert rres.totalUpdates == ores.totalUpdates and rres.meanChiSq == ores.meanChiSq
        assert read(tmp_path / "r2.out") == read(tmp_path / "o2.out")     # the resumed runs checkpoint again
        for f in STATS:
            assert np.array_equal(bits(getattr(ores, f)), bits(getattr(ofull, f))), f


def test_checkpoint_system_as_the_reference_suite_tests_it(oracle, tmp_path):
    """tests/testthat/test_checkpoints.R: GIST, checkpointInterval=51, nIterations=100, seed 22; the second run resumes
    with ANOTHER seed (33) and must reproduce the first run's matrices — seed and random state come from the file."""
    data = load_data("gist")
    ck = tmp_path / "test.out"
    kw = dict(nPatterns=7, nIterations=100, outputFrequency=0)
    run1 = oracle.run(data, options=oracle.options(checkpointInterval=51, checkpointOutFile=ck), seed=22, **kw)
    run2 = oracle.run(data, options=oracle.options(checkpointInFile=ck, checkpointOutFile=tmp_path / "again.out"), seed=33, **kw)
    assert np.array_equal(bits(run1.Amean), bits(run2.Amean)) and np.array_equal(bits(run1.Pmean), bits(run2.Pmean))
    assert run2.seed == 22
    from oracle.harness import RefLib
    if RefLib.available("scalar"):
        ref = RefLib("scalar")
        r1 = ref.run(data, checkpointInterval=51, checkpointOutFile=tmp_path / "ref.out", seed=22, **kw)
        r2 = ref.run(data, checkpointInFile=tmp_path / "ref.out", checkpointOutFile=tmp_path / "ref2.out", seed=33, **kw)
        assert np.array_equal(bits(r1.Amean), bits(r2.Amean)) and np.array_equal(bits(r1.Amean), bits(run1.Amean))
        assert read(tmp_path / "ref.out") == read(ck)


def test_what_the_archive_leaves_out_restarts_like_in_the_reference(oracle, tmp_path):
    """Pump statistics, the chi-square / atom histories, totalUpdates and the queue-length averages are not archived
    (GapsStatistics.cpp:164-169 writes the four sums only): after a resume they restart, identically in the reference
    and the oracle."""
    ref = ref_or_skip()
    data = load_data("gist")
    kw = dict(seed=21, nPatterns=4, nIterations=30, outputFrequency=10, takePumpSamples=1)
    ck = tmp_path / "pump.out"
    rfull = ref.run(data, checkpointInterval=20, checkpointOutFile=ck, **kw)
    rres = ref.run(data, checkpointInFile=ck, checkpointOutFile=tmp_path / "r2.out", **kw)
    ores = oracle.run(data, options=oracle.options(checkpointInFile=ck, checkpointOutFile=tmp_path / "o2.out"), **kw)
    assert np.array_equal(rres.pumpMatrix, ores.pumpMatrix)
    assert np.array_equal(rres.meanPatternAssignment, ores.meanPatternAssignment)
    assert not np.array_equal(rres.pumpMatrix, rfull.pumpMatrix)           # ten samples instead of thirty
    assert np.array_equal(bits(rres.Amean), bits(rfull.Amean))             # while the archived sums carry on exactly
    assert rres.totalUpdates == ores.totalUpdates < rfull.totalUpdates
    assert np.float32(rres.averageQueueLengthA) == np.float32(ores.averageQueueLengthA)


# ------------------------------------------------------------------------------------------------
# GPU: cgb_run_ex and the sampler-level Archive<< / >> against the oracle in device order
# ------------------------------------------------------------------------------------------------
GPU_CASES = {
    "modsim_dense": ("modsim", 25, dict(seed=42, nPatterns=3, nIterations=60, outputFrequency=10)),
    "gist_dense": ("gist", 15, dict(seed=42, nPatterns=5, nIterations=40, outputFrequency=10)),
    "gist_sparse": ("gist", 12, dict(seed=11, nPatterns=4, nIterations=30, outputFrequency=10, useSparseOptimization=1)),
    "syn_wide": ("syn:40:700:4:9", 8, dict(seed=9, nPatterns=4, nIterations=20, outputFrequency=5)),
}


def device_options(oracle, data, kw, **ck):
    from cogaps_b200.sampler import reduction_order_for_length
    g, s = (data.shape[1], data.shape[0]) if kw.get("transposeData") else data.shape
    return oracle.options(reduce="device", math="portable", orderA=reduction_order_for_length(s),
                          orderP=reduction_order_for_length(g), **ck)


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(GPU_CASES))
def test_gpu_checkpoint_file_is_the_oracles_and_resume_is_exact(oracle, name, tmp_path):
    import cogaps_b200 as cg
    dataset, interval, kw = GPU_CASES[name]
    data = load_data(dataset)
    gfile, ofile = tmp_path / "gpu.out", tmp_path / "oracle.out"
    want = oracle.run(data, options=device_options(oracle, data, kw, checkpointInterval=interval, checkpointOutFile=ofile), **kw)
    full = cg.gaps_run(data, checkpointInterval=interval, checkpointOutFile=gfile, **kw)
    # same chain as the oracle with checkpoints on (the AP rebuild at every checkpoint included) ...
    assert np.array_equal(full.atomHistoryA, want.atomHistoryA)
    assert np.array_equal(full.atomHistoryP, want.atomHistoryP)
    assert full.totalUpdates == want.totalUpdates
    for f in STATS:
        assert np.array_equal(bits(getattr(full, f)), bits(getattr(want, f))), f
    # ... and the same file: factor matrices, atoms in pick order, queue and rng states, statistics sums
    assert read(gfile) == read(ofile)
    assert not os.path.exists(str(gfile) + ".backup")
    # resume on the GPU from the GPU's file, and from the oracle's (they are the same bytes, but go through both names)
    for src in (gfile, ofile):
        again = tmp_path / "again.out"
        resumed = cg.gaps_run(data, checkpointInFile=src, checkpointOutFile=again, **kw)
        for f in STATS:
            assert np.array_equal(bits(getattr(resumed, f)), bits(getattr(full, f))), f
        assert np.float32(resumed.meanChiSq) == np.float32(full.meanChiSq)
        c = parse_file(src)
        assert c["phase"] == 2
        n = kw["nIterations"] // kw["outputFrequency"] - c["iter"] // kw["outputFrequency"]   # reports left in the phase
        assert np.array_equal(resumed.atomHistoryA, full.atomHistoryA[-n:])
        assert np.array_equal(resumed.atomHistoryP, full.atomHistoryP[-n:])
        assert np.array_equal(bits(resumed.chisqHistory), bits(full.chisqHistory[-n:]))
    # and the oracle resumes from the GPU's file to the oracle's own end state
    ores = oracle.run(data, options=device_options(oracle, data, kw, checkpointInFile=gfile, checkpointOutFile=tmp_path / "o2.out"), **kw)
    for f in STATS:
        assert np.array_equal(bits(getattr(ores, f)), bits(getattr(want, f))), f


@pytest.mark.gpu
def test_gpu_interrupt_then_resume_from_the_equilibration_phase(oracle, tmp_path):
    """gaps_check_interrupt (GapsRunner.cpp:280) through the callback; the file left behind is from the equilibration
    phase, so the resumed run also restores the annealing schedule position."""
    import cogaps_b200 as cg
    dataset, interval, kw = GPU_CASES["gist_dense"]
    data = load_data(dataset)
    full = cg.gaps_run(data, checkpointInterval=interval, checkpointOutFile=tmp_path / "full.out", **kw)
    polls = []

    def interrupt():
        polls.append(1)
        return len(polls) > interval + 2      # two iterations after the first checkpoint (iteration interval-1)

    part = tmp_path / "part.out"
    with pytest.raises(cg.CogapsError) as e:
        cg.gaps_run(data, checkpointInterval=interval, checkpointOutFile=part, interrupt=interrupt, **kw)
    assert e.value.code == -7
    assert len(polls) == interval + 3         # polled once per iteration, stopped at the first true
    info = cg.checkpoint_info(part)
    assert (info["phase"], info["iter"]) == (1, interval - 1)
    eq = tmp_path / "eq.out"
    assert oracle.run(data, options=device_options(oracle, data, kw, checkpointInterval=interval, checkpointOutFile=eq,
                                                   stopAfterCheckpoints=1), **kw) is None
    assert read(part) == read(eq)
    resumed = cg.gaps_run(data, checkpointInFile=part, checkpointOutFile=tmp_path / "again.out", **kw)
    for f in STATS:
        assert np.array_equal(bits(getattr(resumed, f)), bits(getattr(full, f))), f
    assert np.float32(resumed.meanChiSq) == np.float32(full.meanChiSq)


@pytest.mark.gpu
@pytest.mark.parametrize("sparse", [0, 1])
def test_gpu_sampler_archive_round_trip(sparse):
    """`Archive << sampler` / `>>` of the Sampler concept at the sampler level: a second pair of samplers restored from
    the bytes continues exactly like the first."""
    import cogaps_b200 as cg
    from cogaps_b200._runhelp import make_params
    data = load_data("gist")
    k = 4

    def build(seed):
        params = make_params(nPatterns=k, seed=seed, useSparseOptimization=sparse)
        rs = cg.GapsRandomState(seed)
        a = cg.GibbsSampler(data, True, True, 0.01, 100.0, params, rs)
        p = cg.GibbsSampler(data, False, False, 0.01, 100.0, params, rs)
        return rs, a, p

    def settle(a, p):
        a.sync(p)
        p.sync(a)
        a.extraInitialization()
        p.extraInitialization()

    def steps(a, p, n):
        for _ in range(n):
            a.update(200)
            p.sync(a)
            p.update(40)
            a.sync(p)

    rs1, a1, p1 = build(5)
    settle(a1, p1)
    steps(a1, p1, 6)
    blobA, blobP, seeder = a1.serialize(), p1.serialize(), rs1.getState()
    settle(a1, p1)                           # what createCheckpoint does after writing (GapsRunner.cpp:254-255)

    # the bytes are what they claim to be
    from tests.checkpoint_format import _Cursor, _sampler
    imgA = _sampler(_Cursor(blobA), bool(sparse))
    pos, mass = a1.atoms()
    assert np.array_equal(imgA["pos"], pos) and np.array_equal(bits(imgA["mass"]), bits(mass))
    assert np.array_equal(bits(imgA["matrix"]), bits(a1.getMatrix()))
    assert imgA["minAtoms"] == imgA["maxAtoms"] == pos.size

    rs2, a2, p2 = build(99)                  # another seed: everything that matters must come from the bytes
    rs2.setState(seeder)
    a2.deserialize(blobA)
    p2.deserialize(blobP)
    settle(a2, p2)
    assert np.array_equal(bits(a2.getMatrix()), bits(a1.getMatrix()))
    assert a2.serialize() == blobA and p2.serialize() == blobP
    steps(a1, p1, 4)
    steps(a2, p2, 4)
    for x, y in ((a1, a2), (p1, p2)):
        assert np.array_equal(bits(x.getMatrix()), bits(y.getMatrix()))
        px, mx = x.atoms()
        py, my = y.atoms()
        assert np.array_equal(px, py) and np.array_equal(bits(mx), bits(my))
    assert a1.chiSq() == a2.chiSq()

    # set_atoms alone: same atoms in the same order -> same bytes; a position used twice is refused
    a2.setAtoms(*a1.atoms())
    assert a2.serialize() == a1.serialize()
    with pytest.raises(cg.CogapsError):
        a2.setAtoms(np.array([5, 5], np.uint64), np.array([1.0, 2.0], np.float32))
    # a dense archive does not go into a sparse sampler of the same shape, nor A's into P
    with pytest.raises(cg.CogapsError):
        p2.deserialize(blobA)


def test_reader_survives_random_damage(tmp_path):
    """A checkpoint is an input file: whatever bytes it holds, the reader either accepts a structurally valid archive
    or reports CGB_EINVAL — it never crashes, hangs or allocates from a corrupt header (sizes are checked against the
    bytes that are left before anything is resized)."""
    import cogaps_b200 as cg
    rng = np.random.default_rng(2026)
    path = tmp_path / "fuzz.out"
    accepted = 0
    for name in sorted(CHECKPOINT_CASES):
        raw = read(golden_file(name))
        for trial in range(150):
            b = bytearray(raw)
            kind = trial % 3
            if kind == 0:                                   # flip a few bits anywhere
                for _ in range(int(rng.integers(1, 4))):
                    b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:                                 # overwrite an aligned word with an extreme value
                off = int(rng.integers(0, len(b) - 8))
                b[off:off + 4] = int(rng.choice([0, 1, 0x7FFFFFFF, 0xFFFFFFFF, 0x80000000])).to_bytes(4, "little")
            else:                                           # cut the tail and / or splice a random block
                cut = int(rng.integers(1, len(b)))
                b = b[:cut] + bytearray(rng.integers(0, 256, int(rng.integers(0, 64)), dtype=np.uint8).tobytes())
            path.write_bytes(bytes(b))
            try:
                info = cg.checkpoint_info(path)
                cg.checkpoint_rewrite(path, tmp_path / "fuzz_copy.out")
                assert read(tmp_path / "fuzz_copy.out") == bytes(b) or name == "sparse"   # flag words are re-derived
                assert info["fileBytes"] == len(b)
                accepted += 1
            except cg.CogapsError as e:
                assert e.code == -1
    assert accepted > 0      # damage inside a float payload is still a valid archive


@pytest.mark.gpu
def test_resume_refuses_atom_counts_that_disagree_with_the_atoms(oracle, tmp_path):
    """a damaged or crafted checkpoint whose archived proposal-queue atom counts (minAtoms / maxAtoms) do not
    match the archived atom list must be refused at resume — it would let the first update pick from atoms that do not
    exist — and so must a negative or non-finite atom mass.  The reader alone (cgb_checkpoint_rewrite) accepts such a
    file: the counts are only meaningful against the sampler."""
    import struct
    import cogaps_b200 as cg
    from cogaps_b200 import CogapsError
    dataset, interval, kw = GPU_CASES["gist_dense"]
    data = load_data(dataset)
    good = tmp_path / "good.out"
    oracle.run(data, options=device_options(oracle, data, kw, checkpointInterval=interval, checkpointOutFile=good), **kw)
    raw = bytearray(good.read_bytes())
    c = parse_file(good)
    assert c["A"]["minAtoms"] == c["A"]["maxAtoms"] == c["A"]["pos"].size
    # (1) the A sampler's queue record: rng, minAtoms, maxAtoms follow its atoms
    key = struct.pack("<QQQ", c["A"]["rng"], c["A"]["minAtoms"], c["A"]["maxAtoms"])
    at = bytes(raw).find(key)
    assert at > 0 and bytes(raw).find(key, at + 1) < 0
    for lo, hi in ((c["A"]["minAtoms"] + 3, c["A"]["maxAtoms"] + 3), (c["A"]["minAtoms"] - 1, c["A"]["maxAtoms"])):
        bad = bytearray(raw)
        bad[at + 8:at + 24] = struct.pack("<QQ", lo, hi)
        path = tmp_path / "counts.out"
        path.write_bytes(bytes(bad))
        with pytest.raises(CogapsError) as err:
            cg.gaps_run(data, checkpointInFile=path, checkpointOutFile=tmp_path / "unused.out", **kw)
        assert err.value.code == -1 and "atom counts" in str(err.value)
    # (2) a negative and a NaN mass in the A sampler's first atom (8-byte position, 4-byte mass)
    first = struct.pack("<Qf", int(c["A"]["pos"][0]), float(c["A"]["mass"][0]))
    at = bytes(raw).find(first)
    assert at > 0
    for mass in (-1.0, float("nan")):
        bad = bytearray(raw)
        bad[at + 8:at + 12] = struct.pack("<f", mass)
        path = tmp_path / "mass.out"
        path.write_bytes(bytes(bad))
        with pytest.raises(CogapsError) as err:
            cg.gaps_run(data, checkpointInFile=path, checkpointOutFile=tmp_path / "unused.out", **kw)
        assert err.value.code == -1 and "mass" in str(err.value)
    # the untouched file still resumes
    cg.gaps_run(data, checkpointInFile=good, checkpointOutFile=tmp_path / "unused.out", **kw)


@pytest.mark.gpu
def test_set_atoms_keeps_the_generator_in_step():
    """cgb_sampler_set_atoms used to leave the proposal queue's atom counts stale; an update right after it
    must run (it asserts min == max == domain size on entry, ProposalQueue.cpp:59-60)."""
    import bench
    data = load_data("gist")
    chain = bench.Chain(data, 4, 3)
    for _ in range(5):
        chain.step()
    pos, mass = chain.A.atoms()
    keep = pos.size // 2
    chain.A.setAtoms(pos[:keep], mass[:keep])
    assert chain.A.nAtoms() == keep
    # matrix and atoms are the caller's to keep consistent: every element = the sum of the masses in its bin
    nbins = chain.A.nRows * 4
    binlen = np.uint64(0xFFFFFFFFFFFFFFFF // nbins)
    b = np.minimum((pos[:keep] // binlen).astype(np.int64), nbins - 1)
    M = np.bincount(b, weights=mass[:keep].astype(np.float64), minlength=nbins).reshape(chain.A.nRows, 4).astype(np.float32)
    chain.A.setMatrix(M)
    chain.P.sync(chain.A)
    chain.A.extraInitialization()
    for _ in range(3):
        chain.step()
    assert chain.A.nAtoms() > 0
