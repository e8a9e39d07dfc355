"""-M (flag_magnet) and -s (flag_s) on the CUDA path -- the rest of SURVEY.md section 8f.4 -- against the oracle restatement
(pinned bit for bit to the unmodified reference for these flags too, tests/test_oracle.py) on the same frames, and the
command-line driver against the files the reference binary wrote (tests/golden/velocity/m1_magnet, s1_stress, s2_stress_alone).

Reference: kmu_q1 lmp2sqt.c:1209-1214, C2_mu_mu / C2_rho_mu / C2_mu_rho :1565-1576, C2_v_mu / C2_mu_v :1556-1564 (rrCmm is
never accumulated: :420-432); sumVR_ct :1122-1142 (zeroed under -f only, :1120), orgSumVR :1273-1275, rrMSR1_R :1434-1436,
rrMSR2_VR :1461-1476; accumulation :399-407, 427-432; scaling :614-622, 660-669; files :933-1026, 1060-1072.
"""
import glob
import gzip
import json
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

import pyoracle
from golden_utils import GOLDEN
from lmp2sqt_b200 import Deck, SpacetimeCorr, SqtError, _lib, run_trajectories, synth
from parity import check_raw_sums

pytestmark = pytest.mark.gpu
BIG = 10**9


def _check_fields(got, raw, flags, label):
    n2, n1 = pyoracle.names_for(flags)
    for _, k in n2[7:]:                                       # the C2_* tables: sums over axes and particles that largely cancel
        x, y = got[k], raw[k]
        tol = 1e-6 * np.abs(y) + 1e-9 * np.abs(y).max()
        assert (np.abs(x - y) <= tol).all(), (label, k, float(np.abs(x - y).max()), float(np.abs(y).max()))
        assert np.abs(y).max() > 0
    for _, k in n1[3:]:
        np.testing.assert_allclose(got[k], raw[k], rtol=1e-9, atol=1e-9 * max(1e-300, np.abs(raw[k]).max()), err_msg=label + k)
    if "M" in flags:
        assert not got["rrCmmAv"].any()                       # the reference never adds to rrCmmAv
    if "s" in flags:
        np.testing.assert_allclose(got["rrMSR1_R_Av"], raw["rrMSR1_R_Av"], rtol=1e-9, atol=1e-12, err_msg=label + "MSR1")
        # squares of differences of two sums over all particles: absolute tolerance on the scale of the table
        y = raw["rrMSR2_VR_Av"]
        np.testing.assert_allclose(got["rrMSR2_VR_Av"], y, rtol=1e-7, atol=1e-10 * np.abs(y).max(), err_msg=label + "MSR2")
        assert np.abs(y).max() > 0
        np.testing.assert_allclose(got["rrMSR1_R_Av"].sum(axis=1), got["rrMSDAv"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("flags", ["FtM", "FtVM", "Fts", "s", "ftMs", "FtVMs"])
@pytest.mark.parametrize("S,T,B,ncell,n_atoms,frames,want", [
    (5, 8, 4, 2, 32, 40, "self_ring_vec_kernel"),
    (30, 10, 10, 5, 500, 35, "self_ring_vec_kernel"),      # in.lmp2sqt.150's nCSpatial
    (43, 6, 6, 4, 200, 20, "self_ring_kernel<C=43"),       # recurrence kernel: the per-axis pass reads the slot-minor ring
    (70, 8, 8, 5, 300, 30, "self_ring_mma_kernel"),
    (171, 6, 6, 8, 2000, 16, "self_ring_mma_kernel"),      # config 5's nCSpatial, several particle tiles
])
def test_magnet_and_stress_correlators_match_oracle(flags, S, T, B, ncell, n_atoms, frames, want):
    traj = synth.make_trajectory(n_cells=ncell, n_atoms=n_atoms, n_frames=frames, seed=90 + S, box_jitter=0.004)
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, flags)
    if "F" in flags:
        assert eng.kernel_config.startswith(want), eng.kernel_config
    if "M" in flags:
        assert "dipole-weighted" in eng.kernel_config, eng.kernel_config
    got = res.to_dict()
    _, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=BIG, flags=flags, finish=False)
    assert got["countCorrAv"] == raw["countCorrAv"] > 0
    if "F" in flags:
        check_raw_sums(got, raw, n_atoms, label="%s S%d " % (flags, S))   # everything the new flags do not touch is unchanged
    else:
        for k in ("rrMSDAv", "rrMQDAv"):
            np.testing.assert_allclose(got[k], raw[k], rtol=1e-9, atol=1e-12, err_msg=k)
    _check_fields(got, raw, flags, "%s S%d " % (flags, S))
    eng.close()


def test_magnet_stress_two_files_nCSkip_and_normalisation():
    """Two input files with nCSkip 1 (the cumulative sum_n v (x) r of a run without -f carries over from file to file, like
    the reference's never-zeroed sumVR_ct), then prePrintProcess on the fetched sums."""
    t1 = synth.make_trajectory(n_cells=3, n_atoms=60, n_frames=37, seed=3, box_jitter=0.01)
    t2 = synth.make_trajectory(n_cells=3, n_atoms=60, n_frames=29, seed=4, box_jitter=0.01)
    S, T, B = 9, 8, 8
    for flags in ("FtVMs", "s"):
        deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T, nCSkip=1)
        res, eng = run_trajectories([t1, t2], deck, flags)
        _, raw, _ = pyoracle.run_port(None, S, T, B, nCSkip=1, limitCorrAv=BIG, files=[t1, t2], flags=flags, finish=False)
        got = res.to_dict()
        _check_fields(got, raw, flags, "two files " + flags)
        eng.close()
    t = synth.make_trajectory(n_cells=2, n_frames=40, seed=1)
    deck = Deck(limitCorrAv=0, nCBuffer=5, nCSpatial=4, nCTime=10)
    flushed = []
    res, eng = run_trajectories([t], deck, "FtMs", on_flush=lambda r: flushed.append(r.to_dict()))
    printed, raw, _ = pyoracle.run_port(t, 4, 10, 5, limitCorrAv=0, flags="FtMs")
    assert len(flushed) == 1 and flushed[0]["countCorrAv"] == printed["countCorrAv"] == 15
    _check_fields(flushed[0], printed["_raw_at_print"], "FtMs", "flush ")
    res.normalise(deck, "FtMs")
    g = res.to_dict()
    for k in ("avC2_mu_mu", "avC2_rho_mu", "avC2_mu_rho", "rrMSR1_R_Av", "rrMSR2_VR_Av"):
        np.testing.assert_allclose(g[k], printed[k], rtol=1e-6, atol=1e-9 * np.abs(printed[k]).max(), err_msg=k)
    eng.close()


def test_magnet_and_stress_need_their_arrays():
    eng = SpacetimeCorr(Deck(limitCorrAv=BIG, nCBuffer=2, nCSpatial=4, nCTime=4), "FtM")
    eng.InitSpacetimeCorr()
    with pytest.raises(SqtError, match="dipoles"):
        eng.EvalSpacetimeCorr(np.random.rand(10, 3), 0.0, 5.0)
    eng.close()
    eng = SpacetimeCorr(Deck(limitCorrAv=BIG, nCBuffer=2, nCSpatial=4, nCTime=4), "s")
    eng.InitSpacetimeCorr()
    with pytest.raises(SqtError, match="velocities"):
        eng.EvalSpacetimeCorr(np.random.rand(10, 3), 0.0, 5.0)
    eng.close()
    with pytest.raises(SqtError, match="-V / -M need nCSpatial"):
        SpacetimeCorr(Deck(limitCorrAv=BIG, nCBuffer=2, nCSpatial=300, nCTime=4), "FtM")


def _numbers(path):
    return [[float(x) for x in ln.split()] for ln in open(path) if not ln.startswith("#")]


@pytest.mark.parametrize("name", ["m1_magnet", "s1_stress", "s2_stress_alone"])
@pytest.mark.parametrize("gpus", [1, 2])
def test_cli_magnet_and_stress_cases_match_the_reference_binary(name, gpus, tmp_path):
    """`lmp2sqt_b200.out -F -t -M`, `-F -t -s`, `-s` like the reference binary: same banner lines, same deck rewrite, the
    same set of files and the same numbers at the printed precision."""
    if gpus > torch.cuda.device_count():
        pytest.skip("needs %d GPUs" % gpus)
    g = os.path.join(GOLDEN, "velocity", name)
    c = json.load(open(os.path.join(g, "case.json")))
    shutil.copy(os.path.join(g, "in.lmp2sqt"), tmp_path / "in.lmp2sqt")
    with gzip.open(os.path.join(g, "dump0.lammpstrj.gz"), "rb") as fi, open(tmp_path / "dump0.lammpstrj", "wb") as fo:
        fo.write(fi.read())
    args = ["-" + ch for ch in c["flags"]] + (["-g", "2"] if gpus == 2 else [])
    r = subprocess.run([_lib.CLI_BIN, *args, "-i", "in.lmp2sqt", "dump0.lammpstrj"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert (tmp_path / "in.lmp2sqt").read_bytes() == open(os.path.join(g, "in.lmp2sqt.after"), "rb").read()
    assert r.stdout == open(os.path.join(g, "stdout.txt")).read()
    assert (tmp_path / "output.log").read_text() == open(os.path.join(g, "output.log")).read()
    want = sorted(f for f in os.listdir(g) if f.endswith(".info") or f.startswith("output."))
    made = sorted(f.split(".info")[0] + ".info" if ".info." in f else f for f in os.listdir(tmp_path) if ".info." in f or f.startswith("output."))
    assert made == want
    for fn in want:
        if fn in ("output.log", "output.binary"):
            continue
        ours = glob.glob(str(tmp_path / (fn + ".*")))[0] if fn.endswith(".info") else str(tmp_path / fn)
        assert open(ours).readline() == open(os.path.join(g, fn)).readline() or not fn.endswith(".info") or fn.startswith(("C2_", "SSF"))
        x, y = _numbers(ours), _numbers(os.path.join(g, fn))
        assert [len(a) for a in x] == [len(b) for b in y], fn
        x, y = np.array([v for row in x for v in row]), np.array([v for row in y for v in row])
        if fn.startswith(("C2_", "Stress")):                     # %8.4e / %8.4g of sums that cancel: on the scale of the table
            np.testing.assert_allclose(x, y, rtol=2e-4, atol=2e-4 * np.abs(y).max(), equal_nan=True, err_msg=fn)
        else:
            np.testing.assert_allclose(x, y, rtol=2e-3, atol=1e-4, equal_nan=True, err_msg=fn)
    if "F" in c["flags"]:
        a = pyoracle.parse_output_binary(str(tmp_path / "output.binary"), c["S"], c["T"])
        b = pyoracle.parse_output_binary(os.path.join(g, "output.binary"), c["S"], c["T"])
        for key in ("full-density", "self-density", "cross-density"):
            tol = 1e-6 * np.maximum(np.abs(b[key]), 1e-3 * np.abs(b[key][:1, :]))
            assert np.all(np.abs(a[key] - b[key]) <= tol), key
        np.testing.assert_allclose(a["self-vanHove"], b["self-vanHove"], rtol=1e-12)


def test_cli_magnet_needs_the_dipole_dump_style(tmp_path):
    t = synth.make_trajectory(n_cells=2, n_frames=12, seed=1)
    synth.write_deck(str(tmp_path / "in.lmp2sqt"), limitCorrAv=14825, nCBuffer=2, nCSpatial=4, nCTime=4)
    synth.write_text_dump(t, str(tmp_path / "d.lammpstrj"), velocities=True)
    r = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-M", "-i", "in.lmp2sqt", "d.lammpstrj"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode != 0 and "mux muy muz" in r.stderr
    r = subprocess.run([_lib.CLI_BIN, "-s", "-i", "in.lmp2sqt", "d.lammpstrj"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    assert glob.glob(str(tmp_path / "Stress000.info.*"))
    # -V -M on the 11-column style: both column groups are read (an extension: the reference's reader never reaches them)
    synth.write_deck(str(tmp_path / "in.lmp2sqt"), limitCorrAv=14825, nCBuffer=2, nCSpatial=4, nCTime=4)
    synth.write_text_dump(t, str(tmp_path / "all.lammpstrj"), velocities=True, dipoles=True)
    r = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-V", "-M", "-i", "in.lmp2sqt", "all.lammpstrj"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-800:]
    assert glob.glob(str(tmp_path / "C2_v_mu00*.info.*")) and glob.glob(str(tmp_path / "C2_mu_v00*.info.*"))


def test_integration_snippet_runs_like_the_cli(tmp_path):
    """INTEGRATION.md sections 3 / 3b as a program (tests/data/integration_snippet.c, run_all): compiled here, run with every
    flag on an 11-column dump, it must leave the same files as `lmp2sqt_b200.out -F -t -V -M -s` apart from the deck rewrite."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "lmp2sqt_b200")
    exe = tmp_path / "snippet.out"
    env = {k: v for k, v in os.environ.items() if k != "CC"}
    r = subprocess.run(["gcc", "-std=gnu99", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "data", "integration_snippet.c"),
                        "-L", pkg, "-lsqt_b200", "-lsqt_host", "-lm", "-Wl,-rpath," + pkg, "-o", str(exe)], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    t = synth.make_trajectory(n_cells=2, n_frames=30, seed=4)
    outs = {}
    for kind in ("cli", "snippet"):
        d = tmp_path / kind
        d.mkdir()
        # limitCorrAv as the CLI's count pass would leave it (the snippet, like the patch of INTEGRATION.md, starts after that pass)
        synth.write_deck(str(d / "in.lmp2sqt"), limitCorrAv=14825 if kind == "cli" else 14825 + (30 - 6) // (6 // 3), nCBuffer=3, nCSpatial=5, nCTime=6)
        synth.write_text_dump(t, str(d / "all.lammpstrj"), velocities=True, dipoles=True)
        cmd = [_lib.CLI_BIN, "-F", "-t", "-V", "-M", "-s", "-i", "in.lmp2sqt", "all.lammpstrj"] if kind == "cli" else [str(exe), "in.lmp2sqt", "all.lammpstrj", "1"]
        r = subprocess.run(cmd, cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1500:]
        outs[kind] = {f.split(".info")[0] if ".info." in f else f: (d / f).read_bytes() for f in os.listdir(d) if ".info." in f or f.startswith("output.")}
    assert sorted(outs["cli"]) == sorted(outs["snippet"]) and len(outs["cli"]) == 14
    for k in outs["cli"]:
        if k == "output.binary":
            a, b = (pyoracle.parse_output_binary(str(tmp_path / kind / k), 5, 6) for kind in ("cli", "snippet"))
            for key in ("full-density", "self-density", "cross-density", "self-vanHove"):
                np.testing.assert_allclose(a[key], b[key], rtol=1e-9, atol=1e-12)
        elif k != "output.log":
            x, y = (np.array([float(v) for ln in outs[kind][k].decode().split("\n") if not ln.startswith("#") for v in ln.split()]) for kind in ("cli", "snippet"))
            np.testing.assert_allclose(x, y, rtol=1e-3, atol=1e-3 * max(1e-300, np.abs(y).max()), err_msg=k)
