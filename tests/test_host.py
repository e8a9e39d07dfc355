"""Host C code (libsqt_host.so: deck, text-dump reader, normalisation, writers) against the files the
UNMODIFIED reference wrote (tests/golden/).  No GPU: the raw sums fed to the writers come from the oracle."""
import ctypes as C
import filecmp
import os
import re
import shutil

import numpy as np
import pytest

import pyoracle
from golden_utils import dt_rows_without_uninitialised_columns, CASES, case_dir, case_trajectories, load_case, unpack_dumps
from lmp2sqt_b200 import Deck, Results, _lib, synth
from lmp2sqt_b200._lib import SqtDeck, SqtFrameHeader

libc = C.CDLL(None)
libc.fopen.restype = C.c_void_p
libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
libc.fclose.argtypes = [C.c_void_p]


def test_c_abi_exports_every_declared_symbol():
    """Every function include/sqt.h and include/sqt_host.h declare is exported by the built libraries
    (load + symbol lookup only: no compute call without a GPU)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for header, lib, table in (("sqt.h", _lib.engine_lib(), _lib.ENGINE_API), ("sqt_host.h", _lib.host_lib(), _lib.HOST_API)):
        text = open(os.path.join(root, "include", header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared = set(re.findall(r"\b(sqt_[a-z0-9_]+)\s*\(", text))
        bound = {n for n, _, _ in table}
        assert declared == bound, (header, declared ^ bound)
        for n in declared:
            assert hasattr(lib, n), n


def test_engine_fails_loudly_without_gpu_or_with_bad_args():
    import torch
    from lmp2sqt_b200 import SpacetimeCorr, SqtError
    if not torch.cuda.is_available():
        with pytest.raises(SqtError, match="no usable CUDA device"):
            SpacetimeCorr(Deck(nCSpatial=4, nCTime=10, nCBuffer=5))
    with pytest.raises(SqtError):
        SpacetimeCorr(Deck(nCSpatial=0, nCTime=10, nCBuffer=5))


@pytest.mark.parametrize("name", CASES)
def test_deck_read_and_rewrite_match_reference(name, tmp_path):
    """GetNameList / UpdateNameList: reading the deck the reference was given and rewriting it with the
    limitCorrAv main() computed gives byte-for-byte the deck the reference left behind."""
    c = load_case(name)
    d = SqtDeck()
    ok = _lib.host_lib().sqt_deck_read(os.path.join(case_dir(name), "in.lmp2sqt").encode(), C.byref(d))
    assert (d.nCSpatial, d.nCTime, d.nCBuffer) == (c["S"], c["T"], min(c["B"], c["T"]))
    assert d.limitCorrAv == c["limit"] and d.nCSkip == (c["skip"] or 0)
    assert ok == 0   # the "NameList -- data" and "----" lines match no name => "not ok"; main() ignores it (lmp2sqt.c:199)
    B = d.nCBuffer
    for t in case_trajectories(name):
        d.limitCorrAv += (t.n_frames - d.nCTime) // (d.nCTime // B)
    out = tmp_path / "deck"
    assert _lib.host_lib().sqt_deck_write(str(out).encode(), C.byref(d)) == 0
    assert out.read_bytes() == open(os.path.join(case_dir(name), "in.lmp2sqt.after"), "rb").read()


def test_deck_missing_file_and_template(tmp_path):
    d = SqtDeck()
    assert _lib.host_lib().sqt_deck_read(str(tmp_path / "nope").encode(), C.byref(d)) == -1
    assert _lib.host_lib().sqt_deck_write_template(str(tmp_path / "tmpl").encode()) == 0
    txt = (tmp_path / "tmpl").read_text()
    assert txt.startswith("kVal\t\t0.00000 \n") and txt.endswith("nCSkip\t\t0 \n----\n")


def test_reference_sample_decks_parse():
    """The reference's shipped decks (values quoted in SURVEY.md section 8d) in the NameList syntax."""
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "in.lmp2sqt.150")
        with open(p, "w") as f:
            f.write("NameList -- data\nkVal\t\t0.0981748 \ndeltaT\t\t0.0600000 \nmass\t\t1.00000 \nrVal\t\t0.200000 \n"
                    "limitCorrAv\t14825 \nnCBuffer\t150 \nnCSpatial\t30 \nnCTime\t\t150 \n----\n")
        d = Deck.read(p)
        assert (d.kVal, d.deltaT, d.rVal, d.limitCorrAv, d.nCBuffer, d.nCSpatial, d.nCTime, d.nCSkip) == \
               (0.0981748, 0.06, 0.2, 14825, 150, 30, 150, 0)


def _read_all_frames(path):
    h = _lib.host_lib()
    fp = libc.fopen(path.encode(), b"r")
    frames = []
    hd = SqtFrameHeader()
    while h.sqt_dump_read_header(fp, C.byref(hd)):
        xyz = np.empty((3, hd.n_atoms))
        assert h.sqt_dump_read_atoms(fp, C.byref(hd), xyz[0].ctypes.data, xyz[1].ctypes.data, xyz[2].ctypes.data)
        frames.append((hd.timestep, hd.n_atoms, (hd.xlo, hd.xhi, hd.ylo, hd.yhi, hd.zlo, hd.zhi), xyz.copy()))
    libc.fclose(fp)
    return frames


def test_dump_reader_matches_atof_bit_for_bit(tmp_path):
    t = synth.make_trajectory(n_cells=3, n_atoms=77, n_frames=7, seed=5, box_jitter=0.02)
    # make the numbers nasty: big unwrapped coordinates, negative values, many digits
    t.positions[1] *= -1234.5
    t.positions[2] += 9.87654321e5
    p = tmp_path / "d.lammpstrj"
    synth.write_text_dump(t, str(p))
    frames = _read_all_frames(str(p))
    assert len(frames) == 7
    lines = p.read_text().split("\n")
    per = 9 + 77
    for f, (ts, n, box, xyz) in enumerate(frames):
        assert ts == t.timesteps[f] and n == 77
        ref = np.array([[float(v) for v in ln.split()[2:5]] for ln in lines[f * per + 9:(f + 1) * per]])  # Python float() == strtod
        assert np.array_equal(xyz.T, ref)
        assert box[1] == float(lines[f * per + 5].split()[1])


def test_dump_reader_header_variants_and_garbage(tmp_path):
    h = _lib.host_lib()
    good = ("ITEM: TIMESTEP\n5\nITEM: NUMBER OF ATOMS\n2\nITEM: BOX BOUNDS pp pp ff\n0 1.5\n0 2\n-1 3e0\n"
            "ITEM: ATOMS id type xu yu zu vx vy vz\n1 1 0.5 1e-3 -2.25 9 9 9\n2 1 1 2 3 9 9 9\n")
    (tmp_path / "a").write_text(good + "ITEM: TIMESTEP\n6\nITEM: NUMBER OF ATOMS\n2\nITEM: BOX BOUNDS xx yy zz\n")
    fr = _read_all_frames(str(tmp_path / "a"))
    assert len(fr) == 1 and fr[0][2] == (0.0, 1.5, 0.0, 2.0, -1.0, 3.0)
    assert np.array_equal(fr[0][3], [[0.5, 1.0], [1e-3, 2.0], [-2.25, 3.0]])
    (tmp_path / "b").write_text("hello\n")
    assert _read_all_frames(str(tmp_path / "b")) == []
    (tmp_path / "c").write_text(good[:-12])  # truncated last atom line still parses what is there; truncated file -> stops
    fp = libc.fopen(str(tmp_path / "c").encode(), b"r")
    hd = SqtFrameHeader()
    assert h.sqt_dump_read_header(fp, C.byref(hd)) == 1
    libc.fclose(fp)


def _results_from(raw: dict, S, T) -> Results:
    r = Results(S, T)
    m = {"avF_qq2": "avF", "StdDevF_qq2": "StdDevF", "avF_s_qq2": "avF_s", "StdDevF_s_qq2": "StdDevF_s",
         "avF_d_qq2": "avF_d", "StdDevF_d_qq2": "StdDevF_d", "avDrTable": "avDrTable",
         "rrMSDAv": "rrMSDAv", "rrMQDAv": "rrMQDAv", "rrMSDCMAv": "rrMSDCMAv"}
    for k, v in m.items():
        r.array(k)[...] = raw[v]
    for k in ("countCorrAv", "limitCorrAv", "nPtls", "g_Vol", "kVal", "first_Vol", "first_limitCorrAv"):
        setattr(r.c, k, raw[k])
    return r


@pytest.mark.parametrize("name", CASES)
def test_writers_reproduce_reference_files_byte_for_byte(name, tmp_path):
    """PrintProcess: feed the raw sums (from the oracle, at the moment the reference printed) through
    sqt_host_print_process and compare output.txt / output.binary / SSF / Dt / output.log with the reference's."""
    c = load_case(name)
    trajs = case_trajectories(name)
    printed, raw, log = pyoracle.run_port(None, c["S"], c["T"], c["B"], nCSkip=c["skip"] or 0, limitCorrAv=c["limit"],
                                          files=trajs, flags=c["flags"])
    rp = printed["_raw_at_print"]
    res = _results_from(rp, c["S"], c["T"])
    deck = Deck.read(os.path.join(case_dir(name), "in.lmp2sqt"))
    res.print_process(str(tmp_path), deck, c["flags"], date="GOLD")
    g = case_dir(name)
    has_f, has_t = "f" in c["flags"].lower(), "t" in c["flags"]   # which files exist follows the flags (lmp2sqt.c:347-365, 927, 1033)
    assert filecmp.cmp(tmp_path / "output.log", os.path.join(g, "output.log"), shallow=False)
    for fn in ("output.binary", "output.txt"):
        assert (tmp_path / fn).exists() == has_f == os.path.exists(os.path.join(g, fn))
        if has_f:
            assert filecmp.cmp(tmp_path / fn, os.path.join(g, fn), shallow=False)
    assert (tmp_path / "SSF000.info.GOLD").exists() == has_f == os.path.exists(os.path.join(g, "SSF000.info"))
    if has_f:
        assert (tmp_path / "SSF000.info.GOLD").read_text() == open(os.path.join(g, "SSF000.info")).read()
    assert (tmp_path / "Dt000.info.GOLD").exists() == has_t == os.path.exists(os.path.join(g, "Dt000.info"))
    if has_t:
        got_dt, ref_dt = (tmp_path / "Dt000.info.GOLD").read_text(), open(os.path.join(g, "Dt000.info")).read()
        if all(ln.split()[2:5] == ["0", "0", "0"] for ln in ref_dt.strip().split("\n")[1:]):
            assert got_dt == ref_dt                              # byte for byte when the reference's memory happened to be zero
        assert dt_rows_without_uninitialised_columns(got_dt) == dt_rows_without_uninitialised_columns(ref_dt)
    # the normalised arrays equal what the reference left in its globals
    got = res.to_dict()
    assert np.array_equal(got["avF_qq2"], printed["avF"]) and np.array_equal(got["rrMSDAv"], printed["rrMSDAv"])
    assert np.array_equal(got["avDrTable"], printed["avDrTable"])
    if not has_t:
        assert not got["avDrTable"].any()                        # only -t fills the van Hove table (lmp2sqt.c:1370, 1428)
    # second print in the same directory picks the next free index, probed on the Dt name only (lmp2sqt.c:886-902)
    res2 = _results_from(rp, c["S"], c["T"])
    res2.print_process(str(tmp_path), deck, c["flags"], date="GOLD")
    if has_t:
        assert (tmp_path / "Dt001.info.GOLD").exists() and (tmp_path / "SSF001.info.GOLD").exists() == has_f


def test_factor_dr_matches_oracle():
    t = synth.make_trajectory(n_cells=2, n_frames=30, seed=3)
    S, T, B = 6, 5, 5
    o = pyoracle.OraclePort(S, T, B, limitCorrAv=77)
    o.begin_file()
    for f in range(t.n_frames):
        o.push_frame(t.positions[f], 0.0, t.box[f, 0, 1])
    o.finish()
    res = _results_from(o.printed_raw(), S, T)
    fd = np.zeros(S)
    _lib.host_lib().sqt_host_factor_dr(C.byref(res.c), 0.2, fd.ctypes.data)
    assert np.array_equal(fd, o.factorDr())


@pytest.fixture(autouse=True)
def _threaded_parser_for_small_test_frames(monkeypatch):
    """The reader forks its parser only above 131 072 lines per thread; the test frames are small, so lower the bar to
    keep the threaded path under test."""
    monkeypatch.setenv("SQT_READER_LINES_PER_THREAD", "1000")


def _read_all_frames_mmap(path, threads):
    h = _lib.host_lib()
    r = h.sqt_reader_open(path.encode(), threads)
    assert r
    frames = []
    hd = SqtFrameHeader()
    while h.sqt_reader_next_header(r, C.byref(hd)):
        xyz = np.empty((3, hd.n_atoms))
        if not h.sqt_reader_read_atoms(r, C.byref(hd), xyz[0].ctypes.data, xyz[1].ctypes.data, xyz[2].ctypes.data):
            break
        frames.append((hd.timestep, hd.n_atoms, (hd.xlo, hd.xhi, hd.ylo, hd.yhi, hd.zlo, hd.zhi), xyz.copy()))
    h.sqt_reader_close(r)
    return frames


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_mmap_reader_equals_stream_reader(tmp_path, threads):
    """The threaded, memory-mapped reader must return exactly what the line-by-line reader returns."""
    t = synth.make_trajectory(n_cells=9, n_atoms=2777, n_frames=5, seed=8, box_jitter=0.02)   # > 2048 atoms: several threads
    t.positions[1] *= -321.5
    t.positions[3] += 1.234567e6
    p = tmp_path / "d.lammpstrj"
    synth.write_text_dump(t, str(p))
    # exponent notation and a Windows line ending exercise the strtod fallback
    txt = p.read_text().split("\n")
    txt[9] = "1 1 1.5e+01 -2.5E-3 3.25e0"
    txt[10] = txt[10] + "\r"
    p.write_text("\n".join(txt))
    a, b = _read_all_frames(str(p)), _read_all_frames_mmap(str(p), threads)
    assert len(a) == len(b) == 5
    for (ts1, n1, box1, x1), (ts2, n2, box2, x2) in zip(a, b):
        assert (ts1, n1, box1) == (ts2, n2, box2)
        assert np.array_equal(x1, x2)
    assert a[0][3][0, 0] == 15.0 and a[0][3][1, 0] == -2.5e-3
    # count pass (ReadDumpForOnlyCheck analogue)
    h = _lib.host_lib()
    r = h.sqt_reader_open(str(p).encode(), 1)
    hd = SqtFrameHeader()
    n = 0
    while h.sqt_reader_next_header(r, C.byref(hd)) and h.sqt_reader_skip_atoms(r, C.byref(hd)):
        n += 1
    h.sqt_reader_rewind(r)
    assert h.sqt_reader_next_header(r, C.byref(hd)) == 1 and hd.timestep == 0
    h.sqt_reader_close(r)
    assert n == 5


def test_mmap_reader_truncated_and_missing(tmp_path):
    h = _lib.host_lib()
    assert not h.sqt_reader_open(str(tmp_path / "nope").encode(), 1)
    t = synth.make_trajectory(n_cells=2, n_frames=3, seed=1)
    p = tmp_path / "d"
    synth.write_text_dump(t, str(p))
    data = p.read_bytes()
    (tmp_path / "cut").write_bytes(data[: len(data) - 200])      # last frame loses atom lines
    assert len(_read_all_frames_mmap(str(tmp_path / "cut"), 2)) == 2
    (tmp_path / "empty").write_bytes(b"")
    assert _read_all_frames_mmap(str(tmp_path / "empty"), 2) == []


def test_raw_binary_frame_file_equals_text_dump(tmp_path):
    """Extension pinned by self-consistency: the raw binary frame file carries exactly the coordinates the text dump
    carries (those are rounded to %.6f by synth), and both come out of the same reader interface identically."""
    t = synth.make_trajectory(n_cells=3, n_atoms=77, n_frames=6, seed=4, box_jitter=0.01)
    synth.write_text_dump(t, str(tmp_path / "d.lammpstrj"))
    synth.write_raw_dump(t, str(tmp_path / "d.sqtraw"))
    a = _read_all_frames_mmap(str(tmp_path / "d.lammpstrj"), 2)
    b = _read_all_frames_mmap(str(tmp_path / "d.sqtraw"), 2)
    assert len(a) == len(b) == 6
    for (ts1, n1, box1, x1), (ts2, n2, box2, x2) in zip(a, b):
        assert (ts1, n1, box1) == (ts2, n2, box2) and np.array_equal(x1, x2)
    h = _lib.host_lib()
    r = h.sqt_reader_open(str(tmp_path / "d.sqtraw").encode(), 1)
    assert h.sqt_reader_is_raw(r) == 1
    hd = SqtFrameHeader()
    n = 0
    while h.sqt_reader_next_header(r, C.byref(hd)) and h.sqt_reader_skip_atoms(r, C.byref(hd)):
        n += 1
    h.sqt_reader_close(r)
    assert n == 6
    data = (tmp_path / "d.sqtraw").read_bytes()
    (tmp_path / "cut.sqtraw").write_bytes(data[:-100])          # truncated last frame is not delivered
    assert len(_read_all_frames_mmap(str(tmp_path / "cut.sqtraw"), 1)) == 5


@pytest.mark.parametrize("revision,chunks,extra", [(2, 3, 0), (2, 1, 3), (1, 2, 0), (0, 2, 0)])
def test_lammps_native_binary_dump_equals_text_dump(tmp_path, revision, chunks, extra):
    """SURVEY.md 8f.1, extension pinned by self-consistency: LAMMPS' native binary dump layout (current and pre-2020
    headers, several per-rank chunks, extra per-atom columns) delivers exactly what the text dump delivers."""
    t = synth.make_trajectory(n_cells=3, n_atoms=91, n_frames=6, seed=14, box_jitter=0.01)
    synth.write_text_dump(t, str(tmp_path / "d.lammpstrj"))
    synth.write_lammps_binary(t, str(tmp_path / "d.bin"), revision=revision, chunks=chunks, extra_cols=extra)
    a = _read_all_frames_mmap(str(tmp_path / "d.lammpstrj"), 2)
    b = _read_all_frames_mmap(str(tmp_path / "d.bin"), 2)
    assert len(a) == len(b) == 6
    for (ts1, n1, box1, x1), (ts2, n2, box2, x2) in zip(a, b):
        assert (ts1, n1, box1) == (ts2, n2, box2) and np.array_equal(x1, x2)
    h = _lib.host_lib()
    r = h.sqt_reader_open(str(tmp_path / "d.bin").encode(), 1)
    assert h.sqt_reader_is_raw(r) == 2
    hd = SqtFrameHeader()
    n = 0
    while h.sqt_reader_next_header(r, C.byref(hd)) and h.sqt_reader_skip_atoms(r, C.byref(hd)):
        n += 1
    h.sqt_reader_rewind(r)
    assert h.sqt_reader_next_header(r, C.byref(hd)) == 1 and hd.timestep == 0 and hd.n_atoms == 91 and hd.pbcType == 7
    h.sqt_reader_close(r)
    assert n == 6
    data = (tmp_path / "d.bin").read_bytes()
    (tmp_path / "cut.bin").write_bytes(data[:-100])                # truncated last frame is not delivered
    assert len(_read_all_frames_mmap(str(tmp_path / "cut.bin"), 1)) == 5


def test_lammps_native_binary_rejects_what_the_text_reader_rejects(tmp_path):
    """Triclinic boxes, `dump atom` files and other column orders stop the file like an unknown text header does
    (snapshot.c:296-331)."""
    t = synth.make_trajectory(n_cells=2, n_frames=3, seed=1)
    synth.write_lammps_binary(t, str(tmp_path / "tri.bin"), triclinic=True)
    synth.write_lammps_binary(t, str(tmp_path / "cols.bin"), columns="id type x y z")
    assert _read_all_frames_mmap(str(tmp_path / "tri.bin"), 1) == []
    assert _read_all_frames_mmap(str(tmp_path / "cols.bin"), 1) == []
    data = (tmp_path / "tri.bin").read_bytes()
    (tmp_path / "atom.bin").write_bytes(data.replace(b"DUMPCUSTOM", b"DUMPATOMXX"))
    assert _read_all_frames_mmap(str(tmp_path / "atom.bin"), 1) == []


@pytest.mark.parametrize("threads", [1, 3])
def test_mmap_reader_fast_and_general_line_paths_agree_with_strtod(tmp_path, threads):
    """The single-scan fast path of the mmap reader takes plain decimal lines; everything else (exponents, long
    mantissas, '+', CRLF, extra columns, a short last line, lines near the end of the mapping) must fall back to the
    general path -- and every value must equal strtod's (Python float) whichever path parsed it."""
    rng = np.random.default_rng(5)
    n, frames = 2500, 3
    txt, want = [], []
    for f in range(frames):
        txt += ["ITEM: TIMESTEP", str(f * 10), "ITEM: NUMBER OF ATOMS", str(n), "ITEM: BOX BOUNDS pp pp pp",
                "0.000000 16.796000", "0.0 16.796", "0 16.796", "ITEM: ATOMS id type xu yu zu vx vy vz"]
        rows = []
        for i in range(n):
            x, y, z = rng.normal(0, 50, 3)
            kind = i % 9
            if kind == 0:
                s = ["%.6f" % x, "%.3e" % y, "%.6f" % z]                 # exponent -> general path
            elif kind == 1:
                s = ["%+.6f" % x, "%.6f" % y, "%.17g" % z]               # '+' sign, 17 significant digits
            elif kind == 2:
                s = ["%d" % int(x), "%.1f" % y, "-0.000001"]             # integers, tiny numbers
            elif kind == 3:
                s = ["%.6f" % x, " %.6f" % y, "%.6f" % z]                # double space between tokens
            elif kind == 4:
                s = ["000012.500", "%.6f" % y, "%.6f" % z]               # leading zeros
            else:
                s = ["%.6f" % x, "%.6f" % y, "%.6f" % z]
            tail = " 0.100000 -0.200000 0.300000" + ("\r" if kind == 5 else "")
            txt.append("%d 1 %s %s %s%s" % (i + 1, s[0], s[1], s[2], tail))
            rows.append([float(v) for v in s])
        want.append(np.array(rows))
    p = tmp_path / "mixed.lammpstrj"
    p.write_text("\n".join(txt))                                          # no trailing newline: the last line ends the mapping
    got = _read_all_frames_mmap(str(p), threads)
    assert len(got) == frames
    for (ts, na, box, xyz), w in zip(got, want):
        assert na == n and np.array_equal(xyz.T, w)
    ref = _read_all_frames(str(p))                                        # the stream reader (sqt_dump.c) sees the same
    for (a, b) in zip(got, ref):
        assert np.array_equal(a[3], b[3])


def test_mmap_reader_decimal_fast_path_is_correctly_rounded(tmp_path):
    """20 000 random decimal strings (1..15 significant digits, 0..12 decimals, both signs, leading zeros): the reader's
    one-division fast path must give exactly strtod's double for every one of them."""
    rng = np.random.default_rng(99)
    n = 20000
    toks = []
    for i in range(n * 3):
        nd = int(rng.integers(1, 16))
        digits = "".join(str(int(d)) for d in rng.integers(0, 10, nd))
        dec = int(rng.integers(0, min(nd, 12) + 1))
        s = digits[:nd - dec] + ("." + digits[nd - dec:] if dec else "")
        if s.startswith("."):
            s = "0" + s
        if rng.random() < 0.1 and not s.startswith("-"):
            s = "00" + s
        toks.append(("-" if rng.random() < 0.5 else "") + s)
    lines = ["ITEM: TIMESTEP", "0", "ITEM: NUMBER OF ATOMS", str(n), "ITEM: BOX BOUNDS pp pp pp", "0 1", "0 1", "0 1",
             "ITEM: ATOMS id type xu yu zu"]
    lines += ["%d 1 %s %s %s" % (i + 1, toks[3 * i], toks[3 * i + 1], toks[3 * i + 2]) for i in range(n)]
    p = tmp_path / "dec.lammpstrj"
    p.write_text("\n".join(lines) + "\n" + "x" * 600)      # padding: every atom line is far from the end of the mapping
    (ts, na, box, xyz), = _read_all_frames_mmap(str(p), 2)
    want = np.array([float(t) for t in toks]).reshape(n, 3)
    assert np.array_equal(xyz.T, want)


def _read_all_frames_mmap_two_pass(path, threads):
    """What the command-line driver does: a count pass (sqt_reader_skip_atoms, which indexes the atom blocks), rewind,
    then the parse pass on the same reader."""
    h = _lib.host_lib()
    r = h.sqt_reader_open(path.encode(), threads)
    assert r
    hd = SqtFrameHeader()
    n_frames = 0
    while h.sqt_reader_next_header(r, C.byref(hd)) and h.sqt_reader_skip_atoms(r, C.byref(hd)):
        n_frames += 1
    h.sqt_reader_rewind(r)
    frames = []
    for _ in range(n_frames):
        assert h.sqt_reader_next_header(r, C.byref(hd))
        xyz = np.empty((3, hd.n_atoms))
        assert h.sqt_reader_read_atoms(r, C.byref(hd), xyz[0].ctypes.data, xyz[1].ctypes.data, xyz[2].ctypes.data)
        frames.append((hd.timestep, hd.n_atoms, (hd.xlo, hd.xhi, hd.ylo, hd.yhi, hd.zlo, hd.zhi), xyz.copy()))
    h.sqt_reader_close(r)
    return frames


@pytest.mark.parametrize("threads", [1, 3, 8])
@pytest.mark.parametrize("trailing_newline", [True, False])
def test_mmap_reader_indexed_parse_pass_equals_plain_parse(tmp_path, threads, trailing_newline):
    """The count pass records the start of every 1024-th line of an atom block; the parse pass then starts its threads
    from those marks instead of scanning for newlines again.  Same frames either way, for blocks that are not a multiple of
    the mark spacing, for an unterminated last line, and for a truncated file (the incomplete frame is not counted)."""
    t = synth.make_trajectory(n_cells=11, n_atoms=5123, n_frames=3, seed=21, box_jitter=0.01)
    p = tmp_path / "d.lammpstrj"
    synth.write_text_dump(t, str(p))
    if not trailing_newline:
        p.write_bytes(p.read_bytes()[:-1])
    a = _read_all_frames_mmap(str(p), threads)
    b = _read_all_frames_mmap_two_pass(str(p), threads)
    c = _read_all_frames(str(p))
    assert len(a) == len(b) == len(c) == 3
    for fa, fb, fc in zip(a, b, c):
        assert fa[:3] == fb[:3] == fc[:3]
        assert np.array_equal(fa[3], fb[3]) and np.array_equal(fa[3], fc[3])
        assert np.array_equal(fa[3].T, t.positions[a.index(fa)])
    cut = tmp_path / "cut.lammpstrj"
    cut.write_bytes(p.read_bytes()[:-4000])
    assert len(_read_all_frames_mmap_two_pass(str(cut), threads)) == 2


def test_integration_snippet_compiles_and_links(tmp_path):
    """INTEGRATION.md section 2 is real code: the documented call sequence compiles against include/*.h and links against
    the two shared libraries (not run here: it needs a GPU)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "lmp2sqt_b200")
    exe = tmp_path / "snippet"
    env = {k: v for k, v in os.environ.items() if k != "CC"}
    r = subprocess.run(["gcc", "-std=gnu99", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                        os.path.join(root, "tests", "data", "integration_snippet.c"), "-L", pkg, "-lsqt_b200", "-lsqt_host", "-lm",
                        "-Wl,-rpath," + pkg, "-o", str(exe)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 2                                 # usage path: the binary loads both libraries and starts


# ---- -V / -M / -s files: C2_* tables, Stress file, extra Dt columns ----------------------------------
def _results_from_x(raw: dict, S, T, flags) -> Results:
    r = _results_from(raw, S, T)
    n2, n1 = pyoracle.names_for(flags)
    for _, k in n2[7:] + n1[3:]:
        r.array(k)[...] = raw[k]
    if "s" in flags:
        for k in ("rrMSR1_R_Av", "rrMSR2_VR_Av"):
            r.array(k)[...] = raw[k]
    return r


@pytest.mark.parametrize("name", ["v1_velocity", "m1_magnet", "s1_stress", "s2_stress_alone"])
def test_writers_reproduce_reference_velocity_magnet_stress_files_byte_for_byte(name, tmp_path):
    """PrintProcess with -V / -M / -s: every file the reference binary wrote for the fixture, byte for byte, from the raw sums
    the oracle holds at the moment of the print (lmp2sqt.c:933-993 C2_* tables, :1006-1026 Stress, :1030-1076 Dt columns)."""
    import gzip
    import json
    from golden_utils import GOLDEN, read_dump_text
    g = os.path.join(GOLDEN, "velocity", name)
    c = json.load(open(os.path.join(g, "case.json")))
    with gzip.open(os.path.join(g, "dump0.lammpstrj.gz"), "rt") as f:
        traj = read_dump_text(f.read())
    S, T, flags = c["S"], c["T"], c["flags"]
    printed, raw, log = pyoracle.run_port(None, S, T, c["B"], nCSkip=c["skip"] or 0, limitCorrAv=c["limit"], files=[traj], flags=flags)
    res = _results_from_x(printed["_raw_at_print"], S, T, flags)
    deck = Deck.read(os.path.join(g, "in.lmp2sqt"))
    res.print_process(str(tmp_path), deck, flags, date="GOLD")
    want = sorted(f for f in os.listdir(g) if f.endswith(".info") or f.startswith("output."))
    assert len(want) >= 3
    for fn in want:
        ours = fn.replace(".info", ".info.GOLD") if fn.endswith(".info") else fn
        assert (tmp_path / ours).exists(), fn
        a, b = (tmp_path / ours).read_bytes(), open(os.path.join(g, fn), "rb").read()
        if fn.startswith("Dt") and "s" not in flags:
            assert dt_rows_without_uninitialised_columns(a.decode()) == dt_rows_without_uninitialised_columns(b.decode())
        else:
            assert a == b, fn
    made = sorted(f.replace(".GOLD", "") for f in os.listdir(tmp_path))
    assert made == want                                              # and nothing the reference does not write
    got = res.to_dict()                                              # normalised in place like the reference's globals
    if "s" in flags:
        assert np.array_equal(got["rrMSR2_VR_Av"], printed["rrMSR2_VR_Av"]) and np.array_equal(got["rrMSR1_R_Av"], printed["rrMSR1_R_Av"])
    if "M" in flags:
        assert np.array_equal(got["avC2_mu_mu"], printed["avC2_mu_mu"])


def test_mmap_reader_dipole_and_velocity_columns(tmp_path):
    """-M: mux muy muz are normalised to unit vectors with the reference's own operation order (make_atom_dipole,
    snapshot.c:440-459); the 11-column style yields both dipoles and velocities; asking a style for what it does not carry
    fails the frame."""
    t = synth.make_trajectory(n_cells=2, n_frames=3, seed=2)
    h = _lib.host_lib()

    def read(path, want_v, want_m):
        r = h.sqt_reader_open(str(path).encode(), 2)
        hd = SqtFrameHeader()
        out = []
        while h.sqt_reader_next_header(r, C.byref(hd)):
            n = hd.n_atoms
            a = np.full((9, n), np.nan)
            p = [a[i].ctypes.data for i in range(9)]
            ok = h.sqt_reader_read_atoms_vm(r, C.byref(hd), p[0], p[1], p[2], *(p[3:6] if want_v else [None] * 3), *(p[6:9] if want_m else [None] * 3))
            if not ok:
                break
            out.append(a.copy())
        h.sqt_reader_close(r)
        return out

    synth.write_text_dump(t, str(tmp_path / "mu"), dipoles=True)
    synth.write_text_dump(t, str(tmp_path / "all"), dipoles=True, velocities=True)
    synth.write_text_dump(t, str(tmp_path / "vel"), velocities=True)
    fm = read(tmp_path / "mu", False, True)
    assert len(fm) == 3
    for f in range(3):
        assert np.array_equal(fm[f][:3], t.positions[f].T)
        assert np.array_equal(fm[f][6:9], t.dipoles[f].T)            # bit for bit: same sqrt, same divisions
        np.testing.assert_allclose(np.linalg.norm(fm[f][6:9], axis=0), 1.0, rtol=1e-15)
    fa = read(tmp_path / "all", True, True)
    assert len(fa) == 3 and np.array_equal(fa[1][3:6], t.velocities[1].T) and np.array_equal(fa[1][6:9], t.dipoles[1].T)
    assert np.array_equal(read(tmp_path / "all", False, True)[2][6:9], t.dipoles[2].T)   # the reference's own reading of this style
    assert read(tmp_path / "all", True, False) == []                 # velocities alone: the 8-column velocity style only
    assert read(tmp_path / "vel", False, True) == [] and read(tmp_path / "mu", True, True) == []
    assert np.array_equal(read(tmp_path / "vel", True, False)[0][3:6], t.velocities[0].T)
    (tmp_path / "tiny").write_text((tmp_path / "mu").read_text().replace(" mux muy muz\n1 1 ", " mux muy muz\n1 1 ", 1))
    # a dipole shorter than 1e-3 is left unset by the reference; zero here
    txt = (tmp_path / "mu").read_text().split("\n")
    txt[9] = " ".join(txt[9].split()[:5] + ["0.000100", "0.000200", "0.000100"])
    (tmp_path / "tiny").write_text("\n".join(txt))
    assert not read(tmp_path / "tiny", False, True)[0][6:9, 0].any()


# ---- the writers against the reference's own PrintProcess run in-process, every flag combination (incl. -V -M, which no dump
# file can exercise through the reference's reader) ------------------------------------------------------------------------------
PRINT_FLAGS = ["FtVMs", "FtVM", "ftV", "fM", "ts", "FtMs", "fVM"]
FROZEN_PRINT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "print_process_flags.pkl")


def _print_case(flags):
    traj = synth.make_trajectory(n_cells=2, n_frames=40, seed=5 + len(flags), box_jitter=0.004)
    return traj, 5, 8, 4


def _check_print_files(flags, ref_files, tmp_path):
    traj, S, T, B = _print_case(flags)
    extra = (traj.n_frames - T) // (T // B)
    printed, raw, log = pyoracle.run_port(traj, S, T, B, limitCorrAv=10**9 - extra, flags=flags)   # same limitCorrAv as the reference run
    res = _results_from_x(printed["_raw_at_print"], S, T, flags)
    deck = Deck(limitCorrAv=0, nCBuffer=B, nCSpatial=S, nCTime=T)
    res.print_process(str(tmp_path), deck, flags, date="GOLD")
    made = sorted(os.listdir(tmp_path))
    assert made == sorted(ref_files), (flags, made, sorted(ref_files))          # the same set of files ...
    for fn in made:
        a, b = (tmp_path / fn).read_bytes(), ref_files[fn]
        if fn.startswith("Dt") and "s" not in flags:                            # (msdx msdy msdz: unset memory in the reference without -s)
            assert dt_rows_without_uninitialised_columns(a.decode()) == dt_rows_without_uninitialised_columns(b.decode()), (flags, fn)
        elif fn == "output.log":
            assert a.split()[:4] == b.split()[:4]                               # (the last field is limitCorrAv: 10^9 there, as given here)
        else:
            assert a == b, (flags, fn)                                          # ... byte for byte


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("flags", PRINT_FLAGS)
def test_writers_equal_reference_printprocess_in_process(flags, tmp_path):
    traj, S, T, B = _print_case(flags)
    ref = pyoracle.run_reference([traj], S, T, B, flags=flags, print_files=True, limit=10**9)
    _check_print_files(flags, ref["files"], tmp_path)


@pytest.mark.parametrize("flags", PRINT_FLAGS)
def test_writers_equal_frozen_reference_printprocess_files(flags, tmp_path):
    import pickle
    with open(FROZEN_PRINT, "rb") as f:
        frozen = pickle.load(f)
    _check_print_files(flags, frozen[flags], tmp_path)


if __name__ == "__main__":                                                      # regenerate the frozen files (needs oracle/_ref)
    import pickle
    out = {}
    for fl in PRINT_FLAGS:
        traj, S, T, B = _print_case(fl)
        out[fl] = pyoracle.run_reference([traj], S, T, B, flags=fl, print_files=True, limit=10**9)["files"]
    with open(FROZEN_PRINT, "wb") as f:
        pickle.dump(out, f)


def test_per_type_normalisation_and_types_file(tmp_path):
    """Host side of the per-type partials (an extension): normalisation identities and the
    types%03d.info.<date> file, from raw sums of the numpy statement (oracle/pytypes.py) -- no GPU needed."""
    import pytypes
    traj = synth.make_trajectory(n_cells=2, n_frames=40, seed=21, box_jitter=0.003)
    ty = (np.arange(traj.n_atoms) % 3 == 0).astype(np.int32) + 1
    S, T, B = 5, 8, 4
    tp = pytypes.run_types([traj], ty, S, T, B)
    printed, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=10**8, flags="Ft")
    rp = printed["_raw_at_print"]
    res = Results(S, T, 2)
    for k, v in {"avF_qq2": "avF", "StdDevF_qq2": "StdDevF", "avF_s_qq2": "avF_s", "StdDevF_s_qq2": "StdDevF_s", "avF_d_qq2": "avF_d",
                 "StdDevF_d_qq2": "StdDevF_d", "avDrTable": "avDrTable", "rrMSDAv": "rrMSDAv", "rrMQDAv": "rrMQDAv", "rrMSDCMAv": "rrMSDCMAv"}.items():
        res.array(k)[...] = rp[v]
    for k in ("countCorrAv", "limitCorrAv", "nPtls", "g_Vol", "kVal", "first_Vol", "first_limitCorrAv"):
        setattr(res.c, k, rp[k])
    nt = 2
    for name, shape in (("avF_ab", (nt, nt, S, T)), ("avF_s_a", (nt, S, T)), ("avDrTable_a", (nt, S, T)), ("rrMSD_a", (nt, T)), ("rrMQD_a", (nt, T))):
        np.ctypeslib.as_array(getattr(res.c, name), shape=shape)[...] = tp[name]
    for a in range(nt):
        res.c.nPtls_a[a] = int(tp["nPtls_a"][a])
    assert tp["countCorrAv"] == rp["countCorrAv"]
    deck = Deck(limitCorrAv=0, nCBuffer=B, nCSpatial=S, nCTime=T)
    res.normalise(deck, "Ft")
    fl = _lib.SqtFlags(1, 1, 1, 0, 0, 0)
    dk = deck.to_c()
    assert _lib.host_lib().sqt_host_print_types(str(tmp_path).encode(), C.byref(res.c), C.byref(dk), C.byref(fl), b"GOLD") == 0
    g = res.to_dict()
    Na = g["nPtls_a"].astype(float)
    np.testing.assert_allclose(g["avF_s_a"][:, :, 0], 1.0, rtol=1e-12)                                   # F_s,a(q, 0) = 1
    np.testing.assert_allclose(g["avF_ab"].sum(axis=(0, 1)), g["avF_qq2"], rtol=1e-10, atol=1e-13)      # sum_ab F_ab = F
    np.testing.assert_allclose((Na[:, None, None] * g["avF_s_a"]).sum(axis=0), Na.sum() * g["avF_s_qq2"], rtol=1e-10)
    np.testing.assert_allclose((Na[:, None] * g["rrMSD_a"]).sum(axis=0), Na.sum() * g["rrMSDAv"], rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose((Na[:, None, None] * g["avDrTable_a"]).sum(axis=0)[:, 1:], Na.sum() * g["avDrTable"][:, 1:], rtol=1e-10)
    txt = (tmp_path / "types000.info.GOLD").read_text()
    assert txt.startswith("# per-type partials: n_types 2, atoms %d %d," % (Na[0], Na[1]))
    assert [ln for ln in txt.split("\n") if ln.startswith("#")][1:] == ["# F_ab 1 1", "# F_ab 1 2", "# F_ab 2 1", "# F_ab 2 2", "# F_s_a 1", "# F_s_a 2",
                                                                     "# self-vanHove_a 1", "# self-vanHove_a 2", "# time, then MSD_a MQD_a per species"]
    blk = txt.split("# F_ab 1 2\n")[1].split("\n\n")[0].split("\n")
    assert len(blk) == T and blk[3].split()[1:] == [("%8.4e" % g["avF_ab"][0, 1, nk, 3]).strip() for nk in range(S)]
    # a second print picks the next index
    assert _lib.host_lib().sqt_host_print_types(str(tmp_path).encode(), C.byref(res.c), C.byref(dk), C.byref(fl), b"GOLD") == 0
    assert (tmp_path / "types001.info.GOLD").exists()


def _prefetch_frames(path, want_v, want_m, depth, threads, max_frames):
    h = _lib.host_lib()
    r = h.sqt_reader_open(str(path).encode(), 1)
    hd = SqtFrameHeader()
    n_snap = 0
    while h.sqt_reader_next_header(r, C.byref(hd)) and hd.n_atoms > 0 and h.sqt_reader_skip_atoms(r, C.byref(hd)):
        n_snap += 1
    assert h.sqt_reader_indexed_frames(r) == n_snap
    h.sqt_reader_rewind(r)
    p = h.sqt_prefetch_start(r, int(want_v), int(want_m), depth, threads, min(max_frames, n_snap))
    frames = []
    if p:
        soa = C.c_void_p()
        ncomp = 3 + 3 * want_v + 3 * want_m
        while h.sqt_prefetch_next(p, C.byref(hd), C.byref(soa)):
            a = np.ctypeslib.as_array(C.cast(soa, C.POINTER(C.c_double)), shape=(ncomp, hd.n_atoms)).copy()
            frames.append((hd.timestep, hd.n_atoms, (hd.xlo, hd.xhi), a))
        h.sqt_prefetch_stop(p)
    h.sqt_reader_close(r)
    return frames, n_snap


@pytest.mark.parametrize("threads,depth", [(1, 2), (3, 4), (4, 9)])
def test_frame_parallel_read_ahead_equals_the_sequential_reader(tmp_path, threads, depth):
    """sqt_prefetch_*: whole frames parsed ahead by a few threads, handed out in file order -- the same frames, the same
    numbers, the same stop at the first bad frame as sqt_reader_read_atoms_vm called frame after frame."""
    t = synth.make_trajectory(n_cells=3, n_atoms=77, n_frames=23, seed=8)
    h = _lib.host_lib()
    for name, kw, want_v, want_m in (("plain", {}, False, False), ("vel", {"velocities": True}, True, False), ("all", {"velocities": True, "dipoles": True}, True, True)):
        p = tmp_path / (name + ".lammpstrj")
        synth.write_text_dump(t, str(p), **kw)
        frames, n_snap = _prefetch_frames(p, want_v, want_m, depth, threads, 10**9)
        assert n_snap == 23 and len(frames) == 23
        for f, (ts, n, box, a) in enumerate(frames):
            assert ts == t.timesteps[f] and n == 77 and box == (0.0, t.box[f, 0, 1])
            assert np.array_equal(a[:3], t.positions[f].T)
            if want_v:
                assert np.array_equal(a[3:6], t.velocities[f].T)
            if want_m:
                assert np.array_equal(a[6:9], t.dipoles[f].T)
        assert len(_prefetch_frames(p, want_v, want_m, depth, threads, 7)[0]) == 7           # max_frames
    # a frame that fails in the parse pass (a short line in frame 5, not the last line) ends the stream there
    txt = (tmp_path / "plain.lammpstrj").read_text().split("\n")
    per = 9 + 77
    txt[5 * per + 9 + 3] = "4 1 0.5"
    (tmp_path / "bad.lammpstrj").write_text("\n".join(txt))
    frames, n_snap = _prefetch_frames(tmp_path / "bad.lammpstrj", False, False, depth, threads, 10**9)
    assert n_snap == 23 and len(frames) == 5
    assert h.sqt_prefetch_start(None, 0, 0, 4, 2, 10) is None
