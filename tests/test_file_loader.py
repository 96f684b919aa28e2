"""The path overload of gaps::run (src/GapsRunner.h:19-24): data files are read with the reference's parsing rules
(src/file_parser/).  The loader is host code: these tests need no GPU.  Where oracle/_ref exists the same files go
through the reference's own FileParser + Matrix(path, ...) and must give the same bits."""
import os

import numpy as np
import pytest

from oracle.harness import RefLib


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def _matrix(seed, shape):
    rng = np.random.default_rng(seed)
    m = rng.gamma(2.0, 1.5, shape).astype(np.float32)
    m[rng.random(shape) < 0.3] = 0
    return m


def _sci(v, digits=6):
    """scientific notation the reference accepts: its number test allows digits, '.', '-' only — no '+'
    (MatrixElement.cpp:10-13), so 3.2e+00 is an invalid entry there and here"""
    return (("%." + str(digits) + "e") % v).replace("e+", "e")


def _write_delimited(path, m, sep, row_names, quoted=False, crlf=False):
    eol = "\r\n" if crlf else "\n"
    q = (lambda s: '"%s"' % s) if quoted else (lambda s: s)
    with open(path, "w", newline="") as f:
        header = [q("s%d" % j) for j in range(m.shape[1])]
        f.write(sep.join(([q("")] if row_names else []) + header) + eol)
        for i in range(m.shape[0]):
            cells = [repr(float(v)) if j % 3 else (_sci(v) if v else "0") for j, v in enumerate(m[i])]
            f.write(sep.join(([q("g%d" % i)] if row_names else []) + cells) + eol)


def _expected(m):
    # what the written text parses to: repr() round-trips fp32 exactly; "%.6e" goes through base * powf(10, exp)
    out = m.copy()
    for i in range(m.shape[0]):
        for j in range(m.shape[1]):
            if j % 3 == 0 and m[i, j]:
                base, exp = _sci(m[i, j]).split("e")
                out[i, j] = np.float32(base) * np.float32(np.power(np.float32(10.0), np.float32(float(exp))))
    return out


@pytest.mark.parametrize("kind", ["csv_names", "csv_plain", "tsv_names", "csv_quoted_crlf"])
def test_delimited_files(tmp_path, kind):
    import cogaps_b200 as cg
    m = _matrix(1, (17, 6))
    ext, sep = (".tsv", "\t") if kind.startswith("tsv") else (".csv", ",")
    path = str(tmp_path / ("data" + ext))
    _write_delimited(path, m, sep, row_names=kind != "csv_plain", quoted="quoted" in kind, crlf="crlf" in kind)
    got = cg.read_matrix_file(path)
    assert got.shape == m.shape
    assert np.allclose(got, _expected(m), rtol=2e-7, atol=0)
    if RefLib.available("scalar"):
        assert np.array_equal(bits(got), bits(RefLib("scalar").read_file(path)))


def test_matrix_market(tmp_path):
    import cogaps_b200 as cg
    m = _matrix(2, (23, 9))
    path = str(tmp_path / "data.mtx")
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n% a comment\n")
        nz = np.argwhere(m != 0)
        f.write("%d %d %d\n" % (m.shape[0], m.shape[1], len(nz)))
        for i, j in nz:
            f.write("%d %d %s\n" % (i + 1, j + 1, repr(float(m[i, j])) if (i + j) % 2 else _sci(m[i, j], 5)))
    got = cg.read_matrix_file(path)
    assert got.shape == m.shape
    assert np.array_equal(got == 0, m == 0)
    assert np.allclose(got, m, rtol=2e-5)   # "%.5e" keeps six significant digits
    if RefLib.available("scalar"):
        assert np.array_equal(bits(got), bits(RefLib("scalar").read_file(path)))


def test_gct(tmp_path):
    import cogaps_b200 as cg
    m = _matrix(3, (8, 5))
    path = str(tmp_path / "data.gct")
    with open(path, "w") as f:
        f.write("#1.2\n%d\t%d\n" % m.shape)
        f.write("\t".join(["NAME", "Description"] + ["s%d" % j for j in range(m.shape[1])]) + "\n")
        for i in range(m.shape[0]):
            f.write("\t".join(["g%d" % i, "desc"] + [repr(float(v)) for v in m[i]]) + "\n")
    got = cg.read_matrix_file(path)
    assert np.array_equal(bits(got), bits(m))
    if RefLib.available("scalar"):
        assert np.array_equal(bits(got), bits(RefLib("scalar").read_file(path)))


def test_reference_gist_csv_matches_the_golden_matrix():
    """inst/extdata/GIST.csv as committed under tests/golden (the matrix the reference's own tests factorise)."""
    import cogaps_b200 as cg
    ref_csv = "/root/reference/inst/extdata/GIST.csv"
    if not os.path.exists(ref_csv):
        pytest.skip("reference tree not present on this machine")
    from tests.cases import load_data
    got = cg.read_matrix_file(ref_csv)
    assert np.array_equal(bits(got), bits(load_data("gist")))


def test_bad_files_are_errors(tmp_path):
    import cogaps_b200 as cg
    from cogaps_b200._lib import CogapsError
    with pytest.raises(CogapsError):
        cg.read_matrix_file(str(tmp_path / "missing.csv"))
    bad = tmp_path / "data.txt"
    bad.write_text("1,2\n3,4\n")
    with pytest.raises(CogapsError):
        cg.read_matrix_file(str(bad))              # FileParser.cpp:69-86: unknown extension
    for cell in ("abc", "3.2e+00"):
        junk = tmp_path / "junk.csv"
        junk.write_text(",a,b\nr1,1.0,%s\n" % cell)
        with pytest.raises(CogapsError):
            cg.read_matrix_file(str(junk))         # MatrixElement.cpp:25-29: invalid entry


@pytest.mark.gpu
def test_run_from_file_equals_run_from_memory(tmp_path):
    """gaps::run(path) == gaps::run(matrix) on the matrix the file holds; subsets read from a file are taken in
    increasing index order (Matrix.cpp:113-131)."""
    import cogaps_b200 as cg
    m = _matrix(5, (60, 14))
    path = str(tmp_path / "data.csv")
    _write_delimited(path, m, ",", row_names=True)
    mem = cg.read_matrix_file(path)
    kw = dict(seed=3, nPatterns=3, nIterations=40, outputFrequency=10, maxThreads=1)
    a = cg.gaps_run_file(path, **kw)
    b = cg.gaps_run(mem, **kw)
    assert np.array_equal(bits(a.Amean), bits(b.Amean)) and np.array_equal(bits(a.Pmean), bits(b.Pmean))
    sub = [9, 2, 30, 4, 17, 55, 41]
    a = cg.gaps_run_file(path, subsetGenes=1, subsetIndices=sub, **kw)
    b = cg.gaps_run(mem, subsetGenes=1, subsetIndices=sorted(sub), **kw)
    assert np.array_equal(bits(a.Amean), bits(b.Amean)) and np.array_equal(bits(a.Pmean), bits(b.Pmean))


def _write_mtx(path, m, rng, duplicates=0, explicit_zeros=0, shuffle=True):
    """Matrix-Market text of m with the entries in random order, `duplicates` cells listed twice (the later value is
    the one m holds) and `explicit_zeros` zero cells listed as 0."""
    nz = [(i, j, float(m[i, j])) for i, j in np.argwhere(m != 0)]
    if shuffle:
        rng.shuffle(nz)
    lines = ["%d %d %s" % (i + 1, j + 1, repr(v)) for i, j, v in nz]
    for i, j, v in nz[:duplicates]:
        lines.insert(int(rng.integers(0, lines.index("%d %d %s" % (i + 1, j + 1, repr(v))) + 1)),
                     "%d %d %s" % (i + 1, j + 1, repr(v + 1.5)))          # an earlier, overwritten value
    zeros = np.argwhere(m == 0)
    for i, j in zeros[rng.permutation(len(zeros))[:explicit_zeros]]:
        lines.insert(int(rng.integers(0, len(lines) + 1)), "%d %d 0" % (i + 1, j + 1))
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n")
        f.write("%d %d %d\n" % (m.shape[0], m.shape[1], len(lines)))
        f.write("\n".join(lines) + "\n")


def _csr_of(dense):
    ptr, idx, val = [0], [], []
    for row in dense:
        cols = np.nonzero(row > 0)[0]
        idx.extend(cols)
        val.extend(row[cols])
        ptr.append(len(idx))
    return np.array(ptr, np.uint32), np.array(idx, np.uint32), np.array(val, np.float32)


@pytest.mark.parametrize("shape,seed", [((23, 9), 2), ((7, 64), 3), ((1, 5), 4), ((40, 1), 5)])
def test_matrix_market_straight_to_compressed_rows(tmp_path, shape, seed):
    """with sparseOptimization a .mtx file goes straight to the compressed rows of both samplers
    (SparseMatrix(path), SparseMatrix.cpp:52-107).  They must be what the dense route derives from the same file."""
    import cogaps_b200 as cg
    rng = np.random.default_rng(seed)
    m = _matrix(seed, shape)
    m[0, :] = 0                                            # an empty row, and maybe an empty column
    path = str(tmp_path / "data.mtx")
    _write_mtx(path, m, rng, duplicates=min(3, int((m != 0).sum())), explicit_zeros=min(4, int((m == 0).sum())))
    dense = cg.read_matrix_file(path)
    assert np.array_equal(bits(dense), bits(m))            # the dense reader: last entry for a cell wins
    for by_rows, want in ((True, dense), (False, dense.T)):
        nrow, ncol, ptr, idx, val = cg.read_matrix_csr(path, by_rows=by_rows)
        assert (nrow, ncol) == shape
        wptr, widx, wval = _csr_of(want)
        assert np.array_equal(ptr, wptr) and np.array_equal(idx, widx) and np.array_equal(bits(val), bits(wval))


def test_matrix_market_compressed_rows_edge_cases(tmp_path):
    import cogaps_b200 as cg
    from cogaps_b200._lib import CogapsError
    path = tmp_path / "data.mtx"
    path.write_text("%%MatrixMarket matrix coordinate real general\n3 4 0\n")            # no entries at all
    nrow, ncol, ptr, idx, val = cg.read_matrix_csr(path)
    assert (nrow, ncol, idx.size) == (3, 4, 0) and np.array_equal(ptr, np.zeros(4, np.uint32))
    path.write_text("%%MatrixMarket matrix coordinate real general\n3 4 2\n1 1 2.5\n2 3 -1.0\n")
    with pytest.raises(CogapsError) as e:                                                  # lambda sums negatives too:
        cg.read_matrix_csr(path)                                                           # only the dense route carries them
    assert e.value.code == -5
    path.write_text("%%MatrixMarket matrix coordinate real general\n3 4 1\n4 1 2.5\n")   # outside the declared shape
    with pytest.raises(CogapsError):
        cg.read_matrix_csr(path)
    with pytest.raises(CogapsError):
        cg.read_matrix_csr(tmp_path / "data.csv")


@pytest.mark.gpu
@pytest.mark.parametrize("transpose", [0, 1])
def test_sparse_run_from_matrix_market_never_needs_the_dense_matrix(tmp_path, transpose, monkeypatch):
    """cgb_run_file on a .mtx with sparseOptimization: compressed rows straight from the triplets, the dense copy the
    chi-square kernels read rebuilt on the device (csr_scatter_kernel).  Every output bit equals the dense route's and
    the in-memory run's."""
    import cogaps_b200 as cg
    rng = np.random.default_rng(11)
    m = _matrix(11, (90, 37))
    m[rng.random(m.shape) < 0.5] = 0
    m[3, :] = 0
    path = str(tmp_path / "data.mtx")
    _write_mtx(path, m, rng, duplicates=5, explicit_zeros=6)
    kw = dict(seed=4, nPatterns=4, nIterations=40, outputFrequency=10, useSparseOptimization=1, transposeData=transpose)
    straight = cg.gaps_run_file(path, **kw)
    monkeypatch.setenv("COGAPS_MTX_DENSE", "1")
    dense_route = cg.gaps_run_file(path, **kw)
    monkeypatch.delenv("COGAPS_MTX_DENSE")
    memory = cg.gaps_run(cg.read_matrix_file(path), **kw)
    for other in (dense_route, memory):
        for f in ("Amean", "Asd", "Pmean", "Psd", "chisqHistory"):
            assert np.array_equal(bits(getattr(straight, f)), bits(getattr(other, f))), f
        assert np.array_equal(straight.atomHistoryA, other.atomHistoryA)
        assert np.array_equal(straight.atomHistoryP, other.atomHistoryP)
        assert straight.totalUpdates == other.totalUpdates
        assert np.float32(straight.meanChiSq) == np.float32(other.meanChiSq)


def test_number_tokens_parse_like_the_reference(tmp_path):
    """MatrixElement.cpp:10-46: a token of digits, '.', '-' goes through the stream extractor (longest valid prefix, 0
    when there is none); base 'e' exponent is base * powf(10, exponent).  Odd but legal tokens, against the reference's
    own parser where it is built, and against Python's correctly rounded float32 otherwise."""
    import cogaps_b200 as cg
    tokens = ["7", "007", "3.", ".25", "-0", "-.5", "0.1", "16777217", "123456789.125", "0.000001", "1e3", "2.5e-3", "-1e2",
              "1.17549435e-38", "3.4028234e38", "0.30000001192092896", "4.35", "1e0", "9.999999e-1"]
    path = tmp_path / "tokens.csv"
    path.write_text(",v\n" + "".join("r%d,%s\n" % (i, t) for i, t in enumerate(tokens)))
    got = cg.read_matrix_file(path)[:, 0]
    plain = [i for i, t in enumerate(tokens) if "e" not in t]
    assert np.array_equal(bits(got[plain]), bits(np.array([float(tokens[i]) for i in plain], np.float32)))
    if RefLib.available("scalar"):
        assert np.array_equal(bits(got), bits(RefLib("scalar").read_file(str(path))[:, 0]))
    # the same tokens as Matrix-Market values, through the triplet scanner
    mtx = tmp_path / "tokens.mtx"
    mtx.write_text("%%MatrixMarket matrix coordinate real general\n%d 1 %d\n" % (len(tokens), len(tokens))
                   + "".join("%d 1 %s\n" % (i + 1, t) for i, t in enumerate(tokens)))
    dense = cg.read_matrix_file(mtx)[:, 0]
    assert np.array_equal(bits(dense), bits(got))
    pos = [i for i in range(len(tokens)) if got[i] > 0]
    nonneg = tmp_path / "nonneg.mtx"
    nonneg.write_text("%%MatrixMarket matrix coordinate real general\n%d 1 %d\n" % (len(tokens), len(pos))
                      + "".join("%d  1\t%s\r\n" % (i + 1, tokens[i]) for i in pos))        # odd spacing, CRLF
    nrow, ncol, ptr, idx, val = cg.read_matrix_csr(nonneg, by_rows=False)
    assert np.array_equal(idx, np.array(pos, np.uint32)) and np.array_equal(bits(val), bits(got[pos]))


def test_result_matrices_are_written_like_the_reference(tmp_path):
    """FileParser::writeToCsv (FileParser.h:59-89) / GapsResult::writeToFile (GapsResult.cpp:27-35): byte for byte against
    the reference's own writer where it is built, and readable by the csv reader either way."""
    import ctypes as C
    import cogaps_b200 as cg
    from cogaps_b200._runhelp import make_params, ResultArrays
    rng = np.random.default_rng(8)
    m = (rng.gamma(2.0, 1.5, (19, 5)) * 10.0 ** rng.integers(-9, 9, (19, 5))).astype(np.float32)
    m[0, 0], m[1, 1], m[2, 2], m[3, 3] = 0.0, 1.0, 123456.0, 1234567.0        # %g: 0, 1, 123456, 1.23457e+06
    ours = tmp_path / "ours.csv"
    cg.write_matrix_csv(ours, m)
    text = ours.read_text()
    assert text.splitlines()[0] == '"",' + ",".join('"Col%d"' % j for j in range(5))
    assert text.splitlines()[2].startswith('"Row1",')
    if RefLib.available("scalar"):
        theirs = tmp_path / "theirs.csv"
        RefLib("scalar").write_csv(theirs, m)
        assert ours.read_bytes() == theirs.read_bytes()
    # six significant digits come back; "1.23457e+06" is not a number the reference's reader accepts ('+'), so stay below
    small = np.where(np.abs(m) < 1e5, m, np.float32(2.5)).astype(np.float32)
    small[np.abs(small) < 1e-4] = np.float32(0.125)
    cg.write_matrix_csv(ours, small)
    assert np.allclose(cg.read_matrix_file(ours), small, rtol=1e-5)
    # the four result files
    p = make_params(nPatterns=3)
    res = ResultArrays(p, 7, 4)
    for name, arr in (("Amean", res.Amean), ("Asd", res.Asd), ("Pmean", res.Pmean), ("Psd", res.Psd)):
        arr[...] = rng.random(arr.shape, dtype=np.float32)
    cg.write_result_files(tmp_path / "run", res)
    for name, arr in (("Amean", res.Amean), ("Asd", res.Asd), ("Pmean", res.Pmean), ("Psd", res.Psd)):
        back = cg.read_matrix_file(tmp_path / ("run_3_%s.csv" % name))
        assert back.shape == arr.shape and np.allclose(back, arr, rtol=1e-5, atol=1e-9)
    with pytest.raises(cg.CogapsError, match="must be a csv"):
        cg.write_matrix_csv(tmp_path / "x.tsv", m)


def test_file_info_like_the_reference(tmp_path):
    """getFileInfo_cpp (src/Cogaps.cpp:245-256): dimensions, rowNames (never filled by any parser of the reference),
    colNames (header cells of .csv / .tsv)."""
    import cogaps_b200 as cg
    m = _matrix(4, (6, 4))
    files = {}
    for kind, ext, sep, names, quoted in (("csv_names", ".csv", ",", True, False), ("csv_plain", ".csv", ",", False, False),
                                          ("tsv_quoted", ".tsv", "\t", True, True)):
        path = tmp_path / (kind + ext)
        _write_delimited(str(path), m, sep, row_names=names, quoted=quoted)
        files[kind] = path
    mtx = tmp_path / "data.mtx"
    _write_mtx(str(mtx), m, np.random.default_rng(1))
    files["mtx"] = mtx
    gct = tmp_path / "data.gct"
    with open(gct, "w") as f:
        f.write("#1.2\n%d\t%d\n" % m.shape)
        f.write("\t".join(["NAME", "Description"] + ["s%d" % j for j in range(m.shape[1])]) + "\n")
        for i in range(m.shape[0]):
            f.write("\t".join(["g%d" % i, "desc"] + [repr(float(v)) for v in m[i]]) + "\n")
    files["gct"] = gct
    for kind, path in files.items():
        info = cg.getFileInfo(path)
        assert info["dimensions"] == (6, 4), kind
        assert info["rowNames"] == []
        want = ["s%d" % j for j in range(4)] if kind.startswith(("csv", "tsv")) else []
        assert info["colNames"] == want, kind
        if RefLib.available("scalar"):
            dims, rows, cols = RefLib("scalar").file_info(path)
            assert (dims, rows, cols) == (info["dimensions"], info["rowNames"], info["colNames"]), kind


def test_triplet_scanner_agrees_with_the_dense_reader_on_random_files(tmp_path):
    """Differential test of the two Matrix-Market readers (the stream-extractor one that mirrors MtxParser.cpp and the
    hand-rolled scanner behind the compressed-row route): random files with comments, ragged whitespace, CRLF, duplicate
    cells, explicit zeros, scientific notation and trailing junk must give the same matrix or the same kind of error."""
    import cogaps_b200 as cg
    from cogaps_b200._lib import CogapsError
    rng = np.random.default_rng(77)
    seps = [" ", "  ", "\t", " \t "]
    for trial in range(120):
        nrow, ncol = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        n = int(rng.integers(0, 2 * nrow * ncol + 1))
        lines = ["%%MatrixMarket matrix coordinate real general"]
        for _ in range(int(rng.integers(0, 3))):
            lines.append("% comment " + "x" * int(rng.integers(0, 5)))
        lines.append("%d %d %d" % (nrow, ncol, n))
        for _ in range(n):
            i, j = int(rng.integers(1, nrow + 1)), int(rng.integers(1, ncol + 1))
            kind = int(rng.integers(0, 6))
            v = float(np.float32(rng.gamma(2.0, 2.0)))
            tok = {0: repr(v), 1: "0", 2: "%d" % int(v), 3: _sci(v, 4), 4: "%.3f" % v, 5: ".5"}[kind]
            lines.append(seps[int(rng.integers(0, 4))].join([str(i), str(j), tok]) + ("  " if rng.random() < 0.2 else ""))
        if trial % 10 == 7:
            lines.append("%d %d 1.0" % (nrow + 1, 1))               # outside the declared shape: an error in both
        if trial % 10 == 3:
            lines.append("junk that ends the triplets")              # both readers stop here without complaint
        text = ("\r\n" if trial % 4 == 1 else "\n").join(lines) + ("\n" if trial % 3 else "")
        path = tmp_path / ("f%d.mtx" % trial)
        path.write_text(text)
        try:
            dense = cg.read_matrix_file(path)
        except CogapsError:
            with pytest.raises(CogapsError):
                cg.read_matrix_csr(path)
            continue
        for by_rows, want in ((True, dense), (False, dense.T)):
            r, c, ptr, idx, val = cg.read_matrix_csr(path, by_rows=by_rows)
            wptr, widx, wval = _csr_of(want)
            assert (r, c) == dense.shape
            assert np.array_equal(ptr, wptr) and np.array_equal(idx, widx) and np.array_equal(bits(val), bits(wval)), trial


def test_overflowing_token_saturates_like_the_stream_extractor(tmp_path):
    """a numeric token too large for a float is +/-FLT_MAX through `stringstream >> float` (what the reference's
    parser uses), not +/-inf as strtof returns — an inf would turn D, lambda and chi-square into NaN."""
    import cogaps_b200 as cg
    path = tmp_path / "big.csv"
    path.write_text(',a,b\nr1,1' + '0' * 45 + ',2\nr2,-9' + '9' * 50 + ',0.5\n')
    m = cg.read_matrix_file(path)
    fmax = np.finfo(np.float32).max
    assert m.shape == (2, 2) and m[0, 0] == fmax and m[1, 0] == -fmax and m[0, 1] == 2 and m[1, 1] == 0.5


def test_mtx_dimensions_beyond_32_bits_are_refused(tmp_path):
    import cogaps_b200 as cg
    from cogaps_b200 import CogapsError
    path = tmp_path / "huge.mtx"
    path.write_text("%%MatrixMarket matrix coordinate real general\n5000000000 3 1\n1 1 2.0\n")
    with pytest.raises(CogapsError):
        cg.read_matrix_file(path)
