"""Per-type partials F_ab(q,t), F_s,a(q,t), MSD_a, MQD_a, van Hove_a on the CUDA path (BASELINE config 4
as written).  PARITY UNPINNED UPSTREAM: the reference has no per-type code (`type` is parsed at snapshot.c:362,368 and never
used).  The checker is the numpy statement oracle/pytypes.py, which tests/test_oracle.py anchors to the pinned restatement
(one species = the reference's own numbers), and the sum rules the survey names: sum_ab F_ab = F, sum_a N_a F_s,a = N F_s,
one species = the plain path."""
import glob
import os
import subprocess

import numpy as np
import pytest
import torch

import pyoracle
import pytypes
from lmp2sqt_b200 import Deck, SpacetimeCorr, SqtError, _lib, run_trajectories, synth
from parity import check_raw_sums

pytestmark = pytest.mark.gpu
BIG = 10**9


def _species(n, n_types, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(1, n_types + 1, size=n).astype(np.int32)


def _check_types(got, ref, label):
    assert got["countCorrAv"] == ref["countCorrAv"] > 0, label
    assert list(got["nPtls_a"]) == list(ref["nPtls_a"]), label
    nt = len(ref["nPtls_a"])
    for a in range(nt):
        for b in range(nt):
            x, y = got["avF_ab"][a, b], ref["avF_ab"][a, b]
            scale = np.sqrt(np.abs(ref["avF_ab"][a, a][:, :1] * ref["avF_ab"][b, b][:, :1]))      # sqrt(S_aa(q) S_bb(q)) per q
            tol = 1e-6 * np.maximum(np.abs(y), 1e-3 * scale) + 1e-300
            assert (np.abs(x - y) <= tol).all(), (label, "F_ab", a, b, float(np.abs(x - y).max()))
        x, y = got["avF_s_a"][a], ref["avF_s_a"][a]
        tol = 1e-6 * np.maximum(np.abs(y), 1e-3 * np.abs(y[:, :1])) + 1e-300
        assert (np.abs(x - y) <= tol).all(), (label, "F_s_a", a)
        np.testing.assert_allclose(got["rrMSD_a"][a], ref["rrMSD_a"][a], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(got["rrMQD_a"][a], ref["rrMQD_a"][a], rtol=1e-9, atol=1e-12)
        assert np.array_equal(got["avDrTable_a"][a], ref["avDrTable_a"][a]), (label, "van Hove", a)   # integer counts: exact


def _check_sum_rules(got, label):
    F = got["avF_qq2"]
    np.testing.assert_allclose(got["avF_ab"].sum(axis=(0, 1)), F, rtol=1e-9, atol=1e-9 * np.abs(F).max(), err_msg=label + "sum_ab F_ab = F")
    np.testing.assert_allclose(got["avF_s_a"].sum(axis=0), got["avF_s_qq2"], rtol=1e-12, atol=1e-9, err_msg=label + "sum_a F_s,a = F_s")
    np.testing.assert_allclose(got["rrMSD_a"].sum(axis=0), got["rrMSDAv"], rtol=1e-12, atol=1e-12)
    assert np.array_equal(got["avDrTable_a"].sum(axis=0), got["avDrTable"])
    # lag 0: F_ab(q,0) = Re[rho_a rho_b*] is symmetric in a, b
    np.testing.assert_allclose(got["avF_ab"][..., 0], np.swapaxes(got["avF_ab"], 0, 1)[..., 0], rtol=1e-12, atol=1e-9)


@pytest.mark.parametrize("S,T,B,ncell,n_atoms,frames,n_types,want", [
    (10, 8, 4, 3, 100, 30, 2, "self_ring_vec_kernel"),       # in.lmp2sqt.150.smallQ's nCSpatial
    (30, 10, 10, 5, 500, 35, 3, "self_ring_vec_kernel"),     # in.lmp2sqt.150's nCSpatial
    (43, 6, 6, 4, 200, 20, 2, "self_ring_kernel<C=43"),      # recurrence kernel
    (70, 8, 8, 5, 300, 30, 2, "self_ring_mma_kernel"),
    (128, 6, 6, 8, 2000, 16, 2, "self_ring_mma_kernel"),     # config 4's nCSpatial: two species, several particle tiles each
    (171, 6, 6, 6, 800, 16, 4, "self_ring_mma_kernel"),
])
def test_per_type_partials_match_the_numpy_statement_and_the_plain_path(S, T, B, ncell, n_atoms, frames, n_types, want):
    traj = synth.make_trajectory(n_cells=ncell, n_atoms=n_atoms, n_frames=frames, seed=300 + S, box_jitter=0.004)
    traj.types = _species(n_atoms, n_types, S) if n_types != 2 else np.where(np.arange(n_atoms) % 5 == 0, 2, 1).astype(np.int32)  # 80:20 by id
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, "Ft", n_types=n_types)
    assert eng.kernel_config.startswith(want) and "per-type partials: %d species" % n_types in eng.kernel_config, eng.kernel_config
    got = res.to_dict()
    ref = pytypes.run_types([traj], traj.types, S, T, B)
    _check_types(got, ref, "S%d " % S)
    _check_sum_rules(got, "S%d " % S)
    # the totals are what the reference computes, whatever the species: against the pinned restatement ...
    _, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=BIG, finish=False)
    check_raw_sums(got, raw, n_atoms, label="typed totals S%d " % S)
    # ... and against the plain engine path on the same frames
    plain, e0 = run_trajectories([traj], deck, "Ft")
    p = plain.to_dict()
    for k in ("avF_qq2", "avF_s_qq2", "avF_d_qq2", "rrMSDAv", "rrMQDAv"):
        np.testing.assert_allclose(got[k], p[k], rtol=1e-9, atol=1e-9 * np.abs(p[k]).max(), err_msg=k)
    assert np.array_equal(got["avDrTable"], p["avDrTable"])
    eng.close(); e0.close()


def test_one_species_is_the_plain_path_and_an_empty_species_is_harmless():
    traj = synth.make_trajectory(n_cells=3, n_atoms=90, n_frames=30, seed=17)
    S, T, B = 12, 6, 3
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T)
    plain, e0 = run_trajectories([traj], deck, "Ft")
    one, e1 = run_trajectories([traj], deck, "Ft", n_types=1)          # n_types 1 = off: the very same launches
    assert e0.kernel_config == e1.kernel_config and e0.kernel_launches == e1.kernel_launches
    p, o = plain.to_dict(), one.to_dict()
    for k in ("avF_qq2", "avF_s_qq2", "rrMSDAv"):
        np.testing.assert_allclose(p[k], o[k], rtol=1e-12, atol=1e-9, err_msg=k)   # (dynamic tiles: the summation order varies from run to run)
    # three species declared, none of the second kind
    traj.types = np.where(np.arange(90) % 2 == 0, 1, 3).astype(np.int32)
    res, e3 = run_trajectories([traj], deck, "Ft", n_types=3)
    got = res.to_dict()
    assert list(got["nPtls_a"]) == [45, 0, 45]
    assert not got["avF_ab"][1].any() and not got["avF_ab"][:, 1].any() and not got["avF_s_a"][1].any()
    ref = pytypes.run_types([traj], np.where(traj.types == 3, 2, 1), S, T, B)
    np.testing.assert_allclose(got["avF_ab"][np.ix_([0, 2], [0, 2])], ref["avF_ab"], rtol=1e-9, atol=1e-9 * np.abs(ref["avF_ab"]).max())
    _check_sum_rules(got, "empty species ")
    e0.close(); e1.close(); e3.close()


def test_per_type_two_files_midrun_flush_and_normalisation():
    t1 = synth.make_trajectory(n_cells=3, n_atoms=60, n_frames=37, seed=3, box_jitter=0.01)
    t2 = synth.make_trajectory(n_cells=3, n_atoms=60, n_frames=29, seed=4, box_jitter=0.01)
    t1.types = t2.types = _species(60, 2, 5)
    S, T, B = 9, 8, 8
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T, nCSkip=1)
    res, eng = run_trajectories([t1, t2], deck, "Ft", n_types=2)
    got = res.to_dict()
    _check_types(got, pytypes.run_types([t1, t2], t1.types, S, T, B, nCSkip=1), "two files ")
    eng.close()
    # mid-run flush: the latched block carries the species' sums too, the fresh ones restart from zero
    t = synth.make_trajectory(n_cells=2, n_frames=40, seed=1)
    t.types = _species(32, 2, 6)
    deck = Deck(limitCorrAv=0, nCBuffer=5, nCSpatial=4, nCTime=10)
    flushed = []
    res, eng = run_trajectories([t], deck, "Ft", n_types=2, on_flush=lambda r: flushed.append(r.to_dict()))
    assert len(flushed) == 1 and flushed[0]["countCorrAv"] == 15
    _check_sum_rules(flushed[0], "flush ")
    full = pytypes.run_types([t], t.types, 4, 10, 5)                    # all 17 completed buffers of the run
    now = eng.fetch().to_dict()                                         # what accumulated after the flush
    assert now["countCorrAv"] == full["countCorrAv"] - 15
    np.testing.assert_allclose(flushed[0]["avF_ab"] + now["avF_ab"], full["avF_ab"], rtol=1e-9, atol=1e-9 * np.abs(full["avF_ab"]).max())
    # normalisation: F_s,a(q,0) = 1, sum_ab F_ab = F and sum_a N_a F_s,a = N F_s of the normalised plain arrays
    res.normalise(deck, "Ft")
    res.normalise_types(deck)
    g = res.to_dict()
    np.testing.assert_allclose(g["avF_s_a"][:, :, 0], 1.0, rtol=1e-12)
    np.testing.assert_allclose(g["avF_ab"].sum(axis=(0, 1)), g["avF_qq2"], rtol=1e-9, atol=1e-12)
    Na = g["nPtls_a"].astype(float)
    np.testing.assert_allclose((Na[:, None, None] * g["avF_s_a"]).sum(axis=0), Na.sum() * g["avF_s_qq2"], rtol=1e-9)
    np.testing.assert_allclose((Na[:, None] * g["rrMSD_a"]).sum(axis=0), Na.sum() * g["rrMSDAv"], rtol=1e-9, atol=1e-15)
    eng.close()


def test_per_type_argument_errors():
    deck = Deck(limitCorrAv=BIG, nCBuffer=2, nCSpatial=4, nCTime=4)
    with pytest.raises(SqtError, match="do not combine"):
        SpacetimeCorr(deck, "FtV", n_types=2)
    with pytest.raises(SqtError, match="n_types must be"):
        SpacetimeCorr(deck, "Ft", n_types=9)
    eng = SpacetimeCorr(deck, "Ft", n_types=2)
    eng.InitSpacetimeCorr()
    with pytest.raises(SqtError, match="sqt_set_types"):
        eng.EvalSpacetimeCorr(np.random.rand(10, 3), 0.0, 5.0)          # species not given
    eng.close()
    eng = SpacetimeCorr(deck, "Ft", n_types=2)
    with pytest.raises(SqtError, match="outside 1 .. 2"):
        eng.set_types(np.array([1, 2, 3], dtype=np.int32))
    eng.close()
    eng = SpacetimeCorr(deck, "Ft")
    with pytest.raises(SqtError, match="without n_types"):
        eng.set_types(np.array([1, 1], dtype=np.int32))
    eng.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("S", [10, 100])
def test_per_type_partials_on_two_gpus(S):
    """Particle-sharded: each shard groups ITS atoms by species; rho_a(k) and the species' self blocks cross the allreduces."""
    T, B = 6, 6
    traj = synth.make_trajectory(n_cells=3, n_atoms=101, n_frames=30, seed=78, box_jitter=0.005)
    traj.types = _species(101, 3, 9)
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, "Ft", n_gpus=2, n_types=3)
    assert eng.n_gpus == 2
    got = res.to_dict()
    _check_types(got, pytypes.run_types([traj], traj.types, S, T, B), "2 GPUs S%d " % S)
    _check_sum_rules(got, "2 GPUs ")
    eng.close()


def test_cli_per_type_option(tmp_path):
    """`-T 2`: the species come from the dump's type column; the reference's own files are unchanged, types000.info.<date>
    carries the normalised partials."""
    t = synth.make_trajectory(n_cells=3, n_frames=40, seed=9)
    t.types = np.where(np.arange(t.n_atoms) % 5 == 0, 2, 1).astype(np.int32)
    outs = {}
    two = torch.cuda.device_count() >= 2      # with a second GPU the typed run is also sharded (`-g 2`): one reader, species per shard
    for kind, extra in (("plain", []), ("typed", ["-T", "2"] + (["-g", "2"] if two else []))):
        d = tmp_path / kind
        d.mkdir()
        synth.write_deck(str(d / "in.lmp2sqt"), limitCorrAv=14825, nCBuffer=5, nCSpatial=7, nCTime=10)
        synth.write_text_dump(t, str(d / "dump.lammpstrj"))
        r = subprocess.run([_lib.CLI_BIN, "-F", "-t", *extra, "-i", "in.lmp2sqt", "dump.lammpstrj"], cwd=d, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-1000:]
        outs[kind] = pyoracle.parse_output_binary(str(d / "output.binary"), 7, 10)
    for key in ("full-density", "self-density", "cross-density", "self-vanHove"):
        np.testing.assert_allclose(outs["typed"][key], outs["plain"][key], rtol=1e-9, atol=1e-12)
    f = glob.glob(str(tmp_path / "typed" / "types000.info.*"))
    assert len(f) == 1 and not glob.glob(str(tmp_path / "plain" / "types*"))
    txt = open(f[0]).read()
    assert txt.startswith("# per-type partials: n_types 2, atoms %d %d," % ((t.types == 1).sum(), (t.types == 2).sum()))
    blocks = txt.split("# F_ab ")[1:]
    assert len(blocks) == 4
    Fab = np.array([[[float(v) for v in ln.split()[1:]] for ln in b.split("\n")[1:11]] for b in blocks])   # [ab][lag][nk]
    np.testing.assert_allclose(Fab.sum(axis=0), outs["plain"]["full-density"], rtol=2e-3, atol=2e-4)         # printed at %8.4e
