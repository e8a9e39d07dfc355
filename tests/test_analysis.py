"""Post-processing pipeline (reference: analysis_binary.c) against the files the UNMODIFIED
reference binary wrote for the same inputs (tests/golden/analysis, made by make_golden_analysis.py), and live against
oracle/_ref when it is built.  CPU only: the arithmetic is host C (csrc/host/sqt_analysis.c)."""
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

import pyoracle
from lmp2sqt_b200 import Deck, _lib, synth
from lmp2sqt_b200.analysis import BLOCKS, Analysis

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden", "analysis")
MODES = {"t": [], "t_d2u05": ["-d", "2", "-u", "0.5"], "f": ["-f"], "fw": ["-f", "-w"]}
CASES = [("g2", "t"), ("g2", "t_d2u05"), ("t9", "t"), ("t9", "f"), ("t9", "fw"), ("two_sets", "t"), ("two_sets", "f")]


def _same_text(got: str, want: str, name: str):
    """Byte equality, except that the sign printed for a NaN ("nan" / "-nan") is not part of the contract."""
    if got == want:
        return
    g, w = got.split(), want.split()
    assert len(g) == len(w), name
    for a, b in zip(g, w):
        assert a == b or ("nan" in a and "nan" in b), (name, a, b)


def _run_cli(exe, case, mode, tmp_path):
    d = tmp_path / (case + "_" + mode)
    d.mkdir()
    for f in ("in.lmp2sqt", "output.binary"):
        shutil.copy(os.path.join(GOLD, case, f), d)
    r = subprocess.run([exe, *MODES[mode], "output.binary"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    return {f: (d / f).read_text() for f in sorted(os.listdir(d)) if f.endswith(".info") or f.endswith(".out")}


@pytest.mark.parametrize("case,mode", CASES)
def test_cli_writes_the_reference_files(tmp_path, case, mode):
    want = json.load(open(os.path.join(GOLD, case, mode + ".json")))
    got = _run_cli(_lib.ANALYSIS_BIN, case, mode, tmp_path)
    assert sorted(got) == sorted(want)
    for f in want:
        _same_text(got[f], want[f], f)


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
def test_cli_matches_live_reference_on_a_fresh_run(tmp_path):
    """Not a committed fixture: reference lmp2sqt -> output.binary -> both post-processors."""
    t = synth.make_trajectory(n_cells=3, n_atoms=70, n_frames=80, seed=31)
    src = tmp_path / "src"
    src.mkdir()
    synth.write_text_dump(t, str(src / "dump.lammpstrj"))
    synth.write_deck(str(src / "in.lmp2sqt"), limitCorrAv=0, nCBuffer=17, nCSpatial=7, nCTime=17, nCSkip=1)
    assert pyoracle.run_reference_cli(str(src), "in.lmp2sqt", ["dump.lammpstrj"]).returncode == 0
    for args in ([], ["-f"], ["-f", "-w"], ["-d", "6"]):
        outs = {}
        for tag, exe in (("ref", pyoracle.REF_ANALYSIS_BIN), ("ours", _lib.ANALYSIS_BIN)):
            d = tmp_path / (tag + "".join(args))
            d.mkdir()
            shutil.copy(src / "in.lmp2sqt", d)
            shutil.copy(src / "output.binary", d)
            assert subprocess.run([exe, *args, "output.binary"], cwd=d, capture_output=True).returncode == 0
            outs[tag] = {f: (d / f).read_text() for f in sorted(os.listdir(d)) if f.endswith(".info") or f.endswith(".out")}
        assert sorted(outs["ref"]) == sorted(outs["ours"]) and len(outs["ours"]) >= 26
        for f in outs["ref"]:
            _same_text(outs["ours"][f], outs["ref"][f], f)


def test_tables_through_the_library(tmp_path):
    """The ctypes mirror: definitions of the derived tables checked against their formulas on the loaded F(q,t)."""
    case = os.path.join(GOLD, "t9")
    a = Analysis.load(os.path.join(case, "in.lmp2sqt"), os.path.join(case, "output.binary"))
    S, T = a.c.nCSpatial, a.c.nCTime
    assert (S, T, a.c.nData) == (6, 9, 1)
    raw = pyoracle.parse_output_binary(os.path.join(case, "output.binary"), S, T)
    F_in = {b: raw[b].T.copy() for b in BLOCKS}          # [k][t]
    a.time(outdir=str(tmp_path))
    dt, q = a.c.deltaT, (np.arange(S) + 1) * a.c.kVal
    for j, b in enumerate(BLOCKS[:3]):
        F = a.table("F", j)
        np.testing.assert_array_equal(F, F_in[b])         # F keeps the input, `mean` is scaled by F(q,0)
        np.testing.assert_allclose(a.table("mean", j), F_in[b] / F_in[b][:, :1], rtol=1e-15)
        F1, G, D, Dc, H = (a.table(n, j) for n in ("F1", "Gamma", "D", "Dc", "H"))
        live = ~np.isnan(F1)
        assert live[:, 0].all()
        c = slice(1, T - 1)
        want = (F[:, 2:] - F[:, :-2]) / (2 * dt)
        np.testing.assert_allclose(F1[:, c][live[:, c]], want[live[:, c]], rtol=1e-14)
        np.testing.assert_allclose(D[live], (-G / (q[:, None] ** 2))[live], rtol=1e-15)
        np.testing.assert_allclose(H[live], (D * F[:, :1])[live], rtol=1e-15)
        tt = np.arange(T)[None, :] * dt
        with np.errstate(all="ignore"):
            want_c = -np.log(F / F[:, :1]) / (q[:, None] ** 2 * tt)
        m = live.copy()
        m[:, 0] = False
        np.testing.assert_allclose(Dc[m], want_c[m], rtol=1e-12)
    assert sorted(f for f in os.listdir(tmp_path)) == sorted(b + "00.out" for b in BLOCKS)
    a.close()
    with pytest.raises(ValueError):                       # nCTime 10: 18 points, not a power of two
        g2 = os.path.join(GOLD, "g2")
        Analysis.load(os.path.join(g2, "in.lmp2sqt"), os.path.join(g2, "output.binary")).fourier()
    with pytest.raises(OSError):
        Analysis.load(os.path.join(case, "in.lmp2sqt"), str(tmp_path / "missing.binary"))


def test_from_results_equals_the_file_route(tmp_path):
    """Results in memory -> analysis (no output.binary in between) gives the files the file route gives.
    The raw sums come from the oracle here (CPU test); tests/test_gpu_cli.py does the same from the engine."""
    from test_host import _results_from
    S, T, B = 6, 9, 3
    traj = synth.make_trajectory(n_cells=2, n_frames=50, seed=17)
    printed, _, log = pyoracle.run_port(traj, S, T, B, limitCorrAv=10**6)
    raw = printed["_raw_at_print"]                       # the raw sums at the end-of-run print
    deck = Deck(limitCorrAv=10**6, nCBuffer=B, nCSpatial=S, nCTime=T)
    # route 1: writers -> output.binary -> load
    res = _results_from(raw, S, T)
    d1 = tmp_path / "files"
    d1.mkdir()
    res.print_process(str(d1), deck, "Ft", date="000000")
    synth.write_deck(str(d1 / "in.lmp2sqt"), limitCorrAv=10**6, nCBuffer=B, nCSpatial=S, nCTime=T)
    a1 = Analysis.load(str(d1 / "in.lmp2sqt"), str(d1 / "output.binary"))
    o1 = tmp_path / "o1"
    o1.mkdir()
    a1.time(outdir=str(o1)).print(str(o1))
    # route 2: normalised results -> from_results
    res2 = _results_from(raw, S, T)
    res2.normalise(deck, "Ft")
    a2 = Analysis.from_results(res2, deck, "Ft")
    o2 = tmp_path / "o2"
    o2.mkdir()
    a2.time(outdir=str(o2)).print(str(o2))
    assert sorted(os.listdir(o1)) == sorted(os.listdir(o2)) and len(os.listdir(o1)) == 30
    for f in os.listdir(o1):
        _same_text((o2 / f).read_text(), (o1 / f).read_text(), f)


# ---- the TEXT post-processor (reference: analysis.c = AverageFqt2FqOrFt.c), on output.txt: SURVEY.md section 8f.2, second half ----
GOLD_TEXT = os.path.join(HERE, "golden", "analysis_text")
TEXT_MODES = {"t": [], "t_d2": ["-d", "2"], "f": ["-f"], "fw": ["-f", "-w"]}
TEXT_CASES = [("g5", "t"), ("g5", "t_d2"), ("g2", "t"), ("g4_f_only", "t"), ("t9_two_sets", "t"), ("t9_two_sets", "f"), ("t9_two_sets", "fw")]


def _same_text_but_heap_garbage(got: str, want: str, name: str):
    """Token equality; NaN signs aside.  One more exemption: when the very first decay test of wave number 1 fails, the
    reference reads the word in FRONT of its table (analysis.c:421-425 with ppT = -1: malloc bookkeeping, a denormal) and
    derives one Gamma / D entry from its logarithm; here that word is 0.  Such tokens are a denormal in the reference's file,
    or stand where ours is +-inf / nan because log(0) went in."""
    g, w = got.split(), want.split()
    assert len(g) == len(w), name
    odd = 0
    for a, b in zip(g, w):
        if a == b or ("nan" in a and "nan" in b):
            continue
        try:
            fb = float(b)
        except ValueError:
            raise AssertionError((name, a, b))
        if (abs(fb) < 1e-300 and float(a) == 0.0) or a.lstrip("-") in ("inf", "nan"):
            odd += 1
            continue
        raise AssertionError((name, a, b))
    assert odd <= 4, (name, odd)          # one element per derived table of one block, never more


def _run_text_cli(exe, case, mode, tmp_path):
    d = tmp_path / (case + "_" + mode)
    d.mkdir()
    for f in ("in.lmp2sqt", "output.txt"):
        shutil.copy(os.path.join(GOLD_TEXT, case, f), d)
    r = subprocess.run([exe, *TEXT_MODES[mode], "output.txt"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-500:]
    out = {f: (d / f).read_text() for f in sorted(os.listdir(d)) if f.endswith(".info") or f.endswith(".out")}
    out["<stdout>"] = r.stdout
    return out


@pytest.mark.parametrize("case,mode", TEXT_CASES)
def test_text_cli_writes_the_reference_files(tmp_path, case, mode):
    """bin/analysis_b200.out against what the unmodified analysis.c wrote for the same output.txt (files AND stdout),
    including its stale row parse (column 0 of the density blocks is the time) and the `-f`-only file whose self / cross
    blocks carry no rows."""
    want = json.load(open(os.path.join(GOLD_TEXT, case, mode + ".json")))
    got = _run_text_cli(_lib.ANALYSIS_TEXT_BIN, case, mode, tmp_path)
    assert sorted(got) == sorted(want)
    for f in want:
        _same_text_but_heap_garbage(got[f], want[f], f)
    if case == "g5":                                   # a case without the out-of-table read: byte for byte
        for f in want:
            _same_text(got[f], want[f], f)


@pytest.mark.skipif(not os.path.exists(pyoracle.REF_ANALYSIS_TEXT_BIN), reason="oracle/_ref/analysis_ref.out not built")
def test_text_cli_matches_live_reference_on_a_fresh_run(tmp_path):
    t = synth.make_trajectory(n_cells=3, n_atoms=70, n_frames=80, seed=37)
    src = tmp_path / "src"
    src.mkdir()
    synth.write_text_dump(t, str(src / "dump.lammpstrj"))
    synth.write_deck(str(src / "in.lmp2sqt"), limitCorrAv=0, nCBuffer=17, nCSpatial=7, nCTime=17, nCSkip=1)
    assert pyoracle.run_reference_cli(str(src), "in.lmp2sqt", ["dump.lammpstrj"]).returncode == 0
    for args in ([], ["-f"], ["-f", "-w"], ["-d", "6"]):
        outs = {}
        for tag, exe in (("ref", pyoracle.REF_ANALYSIS_TEXT_BIN), ("ours", _lib.ANALYSIS_TEXT_BIN)):
            d = tmp_path / (tag + "".join(args))
            d.mkdir()
            shutil.copy(src / "in.lmp2sqt", d)
            shutil.copy(src / "output.txt", d)
            r = subprocess.run([exe, *args, "output.txt"], cwd=d, capture_output=True, text=True)
            assert r.returncode == 0
            outs[tag] = {f: (d / f).read_text() for f in sorted(os.listdir(d)) if f.endswith(".info") or f.endswith(".out")}
            outs[tag]["<stdout>"] = r.stdout
        assert sorted(outs["ref"]) == sorted(outs["ours"]) and len(outs["ours"]) >= 17
        for f in outs["ref"]:
            _same_text_but_heap_garbage(outs["ours"][f], outs["ref"][f], f)


def test_text_post_processor_reads_our_own_output_txt(tmp_path):
    """The chain a user of the reference runs (BASELINE config 5's tail): results -> output.txt (our writers, byte-identical to
    the reference's on the golden cases, tests/test_host.py) -> text post-processor -> Fqt.* tables that carry the parsed
    columns.  Checks the documented stale parse: column k of Fqt.full-density is output.txt's k-th number of each row."""
    case = os.path.join(GOLD_TEXT, "g5")
    d = tmp_path / "run"
    d.mkdir()
    for f in ("in.lmp2sqt", "output.txt"):
        shutil.copy(os.path.join(case, f), d)
    assert subprocess.run([_lib.ANALYSIS_TEXT_BIN, "output.txt"], cwd=d, capture_output=True).returncode == 0
    rows = [ln.split() for ln in (d / "output.txt").read_text().split("\n")[2:32]]           # 30 lags of the full-density block
    tab = [ln.split() for ln in (d / "Fqt.full-density.info").read_text().strip().split("\n")[1:]]
    S = 10
    for t, r in enumerate(tab):
        assert len(r) == S + 1
        for k in range(S):
            assert float(r[1 + k]) == pytest.approx(float(rows[t][k]), rel=1e-4, abs=1e-12)       # %9.4e of the k-th number of the row
