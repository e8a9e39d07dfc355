"""Drop-in check: the command-line driver (lmp2sqt_b200/bin/lmp2sqt_b200.out, C host code over the C-ABI) run
exactly like the reference binary on the golden cases -- same flags, same deck, same text dumps -- must leave
the same files behind: the rewritten deck byte for byte, output.binary / output.txt / SSF / Dt equal to the
reference's files within the floating-point tolerance (values are printed at %8.4e / %8.4g, so text is
compared numerically at the printed precision)."""
import glob
import os
import shutil
import subprocess

import numpy as np
import pytest

import pyoracle
from golden_utils import CASES, case_dir, load_case, unpack_dumps
from lmp2sqt_b200 import _lib

pytestmark = pytest.mark.gpu


def _numbers(path):
    out = []
    for ln in open(path):
        if ln.startswith("#"):
            continue
        out.append([float(x) for x in ln.split()])
    return out


@pytest.mark.parametrize("name", CASES)
def test_cli_is_a_drop_in_for_the_reference_binary(name, tmp_path):
    c = load_case(name)
    g = case_dir(name)
    shutil.copy(os.path.join(g, "in.lmp2sqt"), tmp_path / "in.lmp2sqt")
    dumps = unpack_dumps(name, str(tmp_path))
    _run_and_compare(c, g, dumps, tmp_path, [])


FLAGS = {"Ft": ["-F", "-t"], "ft": ["-f", "-t"], "F": ["-F"], "t": ["-t"], "FtV": ["-F", "-t", "-V"]}


def _run_and_compare(c, g, dumps, tmp_path, extra_args, expect_in_stderr=None):
    flags = FLAGS[c["flags"]]
    r = subprocess.run([_lib.CLI_BIN, *flags, *extra_args, "-i", "in.lmp2sqt", *dumps], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    if expect_in_stderr:
        assert expect_in_stderr in r.stderr, r.stderr[-600:]
    # deck rewrite (UpdateNameList): byte-identical
    assert (tmp_path / "in.lmp2sqt").read_bytes() == open(os.path.join(g, "in.lmp2sqt.after"), "rb").read()
    # stdout banner lines of the flags
    assert r.stdout == open(os.path.join(g, "stdout.txt")).read()
    assert (tmp_path / "output.log").read_text() == open(os.path.join(g, "output.log")).read()
    S, T = c["S"], c["T"]
    has_f, has_t = "f" in c["flags"].lower(), "t" in c["flags"]
    # which files exist follows the flags (lmp2sqt.c:347-365, 927, 1033): output.* and SSF under -f / -F, Dt under -t
    assert (tmp_path / "output.binary").exists() == has_f == os.path.exists(os.path.join(g, "output.binary"))
    assert bool(glob.glob(str(tmp_path / "Dt000.info.*"))) == has_t == os.path.exists(os.path.join(g, "Dt000.info"))
    if has_f:
        _compare_f_files(c, g, tmp_path, S, T)
    text_files = (("SSF000.info.*", "SSF000.info"),) * has_f + (("Dt000.info.*", "Dt000.info"),) * has_t
    if "V" in c["flags"]:
        text_files += (("C2_v_v000.info.*", "C2_v_v000.info"), ("C2_v_rho000.info.*", "C2_v_rho000.info"), ("C2_rho_v000.info.*", "C2_rho_v000.info"))
    for pat, gold in text_files:
        f = glob.glob(str(tmp_path / pat))
        assert len(f) == 1, pat
        x, y = np.array(_numbers(f[0])), np.array(_numbers(os.path.join(g, gold)))
        assert x.shape == y.shape
        if gold.startswith("Dt"):                                # msdx/msdy/msdz: uninitialised memory in the reference
            x, y = np.delete(x, [2, 3, 4], axis=1), np.delete(y, [2, 3, 4], axis=1)
            assert open(f[0]).readline() == open(os.path.join(g, gold)).readline()      # column header (Cvv Cvcmvcm with -V)
        if gold.startswith("C2_"):                               # %8.4e of sums that cancel: compare on the scale of the column
            np.testing.assert_allclose(x, y, rtol=2e-4, atol=2e-4 * np.abs(y).max(), equal_nan=True, err_msg=gold)
        else:
            np.testing.assert_allclose(x, y, rtol=2e-3, atol=1e-4, equal_nan=True, err_msg=gold)


def _compare_f_files(c, g, tmp_path, S, T):
    got = pyoracle.parse_output_binary(str(tmp_path / "output.binary"), S, T)
    ref = pyoracle.parse_output_binary(os.path.join(g, "output.binary"), S, T)
    assert got["size"] == ref["size"] and got["magic"] == ref["magic"]
    assert (got["kVal"], got["deltaT"], got["rVal"]) == (ref["kVal"], ref["deltaT"], ref["rVal"])
    for key in ("full-density", "self-density", "cross-density"):
        if ref[key] is None:
            assert got[key] is None
            continue
        tol = 1e-6 * np.maximum(np.abs(ref[key]), 1e-3 * np.abs(ref[key][:1, :]))   # SURVEY.md 8a tolerance
        assert np.all(np.abs(got[key] - ref[key]) <= tol), key
    np.testing.assert_allclose(got["self-vanHove"], ref["self-vanHove"], rtol=1e-12)    # integer counts x factorDr
    if "t" not in c["flags"]:
        assert not ref["self-vanHove"].any() and not got["self-vanHove"].any()          # only -t fills the table (lmp2sqt.c:1370)
    # text files, numerically at the printed precision
    a, b = _numbers(tmp_path / "output.txt"), _numbers(os.path.join(g, "output.txt"))
    assert [len(x) for x in a] == [len(x) for x in b]
    for x, y in zip(a, b):
        # sigma columns can be sqrt(tiny negative) = nan in both
        np.testing.assert_allclose(x, y, rtol=2e-4, atol=1e-7, equal_nan=True)


def test_cli_velocity_case_matches_the_reference_binary(tmp_path):
    """-V: C2_v_v / C2_v_rho / C2_rho_v files, the Cvv / Cvcmvcm columns of the Dt file and everything
    else, against the files the reference wrote with `-F -t -V` on the same 8-column dump (tests/golden/velocity)."""
    import json
    g = os.path.join(os.path.dirname(case_dir("g1_midrun_print")), "velocity", "v1_velocity")
    with open(os.path.join(g, "case.json")) as f:
        c = json.load(f)
    shutil.copy(os.path.join(g, "in.lmp2sqt"), tmp_path / "in.lmp2sqt")
    import gzip
    with gzip.open(os.path.join(g, "dump0.lammpstrj.gz"), "rb") as fi, open(tmp_path / "dump0.lammpstrj", "wb") as fo:
        fo.write(fi.read())
    _run_and_compare(c, g, ["dump0.lammpstrj"], tmp_path, [])


def test_cli_velocity_needs_the_velocity_dump_style(tmp_path):
    """-V on a dump without vx vy vz: the reference correlates uninitialised memory (snapshot.c:426-437); here it is an error."""
    from lmp2sqt_b200 import synth
    t = synth.make_trajectory(n_cells=2, n_frames=12, seed=1)
    synth.write_deck(str(tmp_path / "in.lmp2sqt"), limitCorrAv=14825, nCBuffer=2, nCSpatial=4, nCTime=4)
    synth.write_text_dump(t, str(tmp_path / "d.lammpstrj"))
    r = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-V", "-i", "in.lmp2sqt", "d.lammpstrj"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "id type xu yu zu vx vy vz" in r.stderr


def test_cli_refuses_nCSpatial_beyond_the_kernels(tmp_path):
    """The reference takes any nCSpatial; the kernels cover up to 1024 (256 with -V): refused before the count pass and
    before the deck would be rewritten."""
    from lmp2sqt_b200 import synth
    t = synth.make_trajectory(n_cells=2, n_frames=8, seed=1)
    synth.write_deck(str(tmp_path / "in.lmp2sqt"), limitCorrAv=14825, nCBuffer=2, nCSpatial=1025, nCTime=4)
    before = (tmp_path / "in.lmp2sqt").read_bytes()
    synth.write_text_dump(t, str(tmp_path / "d.lammpstrj"))
    r = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-i", "in.lmp2sqt", "d.lammpstrj"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "beyond what the B200 kernels cover" in r.stderr
    assert (tmp_path / "in.lmp2sqt").read_bytes() == before


@pytest.mark.skipif(not os.path.exists(pyoracle.REF_ANALYSIS_BIN), reason="oracle/_ref/analysis_binary_ref.out not built")
@pytest.mark.parametrize("name", ["g2_final_print", "g5_ring30"])
def test_reference_analysis_binary_consumes_our_output(name, tmp_path):
    """the reference's own post-processor (analysis_binary.c, unmodified, built into
    oracle/_ref) run on OUR output.binary gives the same Fqt / Fqt_scaled / GammaQt / Dqt / D0Hqt files as when it
    is run on the reference's output.binary."""
    c = load_case(name)
    g = case_dir(name)
    ours, theirs = tmp_path / "ours", tmp_path / "theirs"
    ours.mkdir(); theirs.mkdir()
    shutil.copy(os.path.join(g, "in.lmp2sqt"), ours / "in.lmp2sqt")
    dumps = unpack_dumps(name, str(ours))
    r = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-i", "in.lmp2sqt", *dumps], cwd=ours, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    shutil.copy(os.path.join(g, "output.binary"), theirs / "output.binary")
    shutil.copy(os.path.join(g, "in.lmp2sqt.after"), theirs / "in.lmp2sqt")   # analysis_binary reads ./in.lmp2sqt (:184)
    for d in (ours, theirs):
        rr = subprocess.run([pyoracle.REF_ANALYSIS_BIN, "output.binary"], cwd=d, capture_output=True, text=True)
        assert rr.returncode == 0
    files = sorted(f for f in os.listdir(theirs) if f.endswith(".info") or f.endswith(".out"))
    assert len(files) >= 20
    for f in files:
        a, b = _numbers(ours / f), _numbers(theirs / f)
        assert [len(x) for x in a] == [len(x) for x in b], f
        derived = f.startswith(("GammaQt", "Dqt", "D0Hqt", "Fqt1"))     # finite differences of F and ln F amplify rounding
        for x, y in zip(a, b):
            np.testing.assert_allclose(x, y, rtol=1e-4 if derived else 2e-6, atol=1e-6 if derived else 1e-9, equal_nan=True, err_msg=f)


def test_cli_raw_binary_dump_equals_text_dump(tmp_path):
    """the binary frame source behind the reader interface. Same coordinates as text => same files."""
    from lmp2sqt_b200 import synth
    t = synth.make_trajectory(n_cells=3, n_frames=45, seed=9)
    outs = {}
    for kind in ("txt", "raw", "lammpsbin"):
        d = tmp_path / kind
        d.mkdir()
        synth.write_deck(str(d / "in.lmp2sqt"), limitCorrAv=14825, nCBuffer=5, nCSpatial=7, nCTime=10)
        dump = {"txt": "dump.lammpstrj", "raw": "dump.sqtraw", "lammpsbin": "dump.bin"}[kind]
        {"txt": synth.write_text_dump, "raw": synth.write_raw_dump, "lammpsbin": synth.write_lammps_binary}[kind](t, str(d / dump))
        r = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-i", "in.lmp2sqt", dump], cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1000:]
        outs[kind] = pyoracle.parse_output_binary(str(d / "output.binary"), 7, 10)
    for kind in ("raw", "lammpsbin"):
        for key in ("full-density", "self-density", "cross-density", "self-vanHove"):
            np.testing.assert_allclose(outs[kind][key], outs["txt"][key], rtol=1e-11, atol=1e-12)
        assert (tmp_path / kind / "in.lmp2sqt").read_bytes() == (tmp_path / "txt" / "in.lmp2sqt").read_bytes()


def test_analysis_on_device_results_equals_the_file_chain(tmp_path, monkeypatch):
    """the post-processing pipeline applied directly to results fetched from the engine gives the same
    files as CLI -> output.binary -> analysis CLI (and those equal the reference's, tests/test_analysis.py)."""
    from lmp2sqt_b200 import Deck, run_trajectories, synth
    from lmp2sqt_b200.analysis import Analysis
    S, T, B = 70, 9, 9                                     # nCSpatial 70: the tensor-core kernels produce the correlators
    monkeypatch.setenv("SQT_DETERMINISTIC", "1")           # both routes sum in the same order: the files must agree to the last digit
    t = synth.make_trajectory(n_cells=3, n_frames=40, seed=23)
    d = tmp_path / "chain"
    d.mkdir()
    synth.write_deck(str(d / "in.lmp2sqt"), limitCorrAv=14825, nCBuffer=B, nCSpatial=S, nCTime=T)
    synth.write_text_dump(t, str(d / "dump.lammpstrj"))
    r = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-i", "in.lmp2sqt", "dump.lammpstrj"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    r = subprocess.run([_lib.ANALYSIS_BIN, "output.binary"], cwd=d, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-1000:]
    deck = Deck(limitCorrAv=14825, nCBuffer=B, nCSpatial=S, nCTime=T)  # run_trajectories adds the computed part like the CLI does
    res, eng = run_trajectories([t], deck, "Ft")
    res.normalise(deck, "Ft")
    o = tmp_path / "direct"
    o.mkdir()
    Analysis.from_results(res, deck, "Ft").time(outdir=str(o)).print(str(o))
    names = sorted(f for f in os.listdir(o))
    assert len(names) == 30
    for f in names:
        a, b = (o / f).read_text().split(), (d / f).read_text().split()
        assert len(a) == len(b), f
        for x, y in zip(a, b):
            assert x == y or ("nan" in x and "nan" in y), (f, x, y)
