"""-V (flag_velocity) and the getopt flag combinations on the CUDA path, against the oracle
restatement (pinned bit for bit to the unmodified reference, tests/test_oracle.py incl. the -V fixture) on the same frames.

Reference: kvel_q1 lmp2sqt.c:1203-1208, orgV :1283-1287, Cvv / Cvcmvcm :1397-1401, 1441, 1452-1457, C2_v_v / C2_rho_v /
C2_v_rho :1542-1555, accumulation :394-398, 423-426; flags: EvalOneTimeCorr :1246-1249, EvalTwoTimeEach :1370, :1428,
EvalTwoTimeKSpace :1479, :1527.
"""
import numpy as np
import pytest

import pyoracle
from lmp2sqt_b200 import Deck, SpacetimeCorr, run_trajectories, synth
from parity import check_raw_sums

pytestmark = pytest.mark.gpu
BIG = 10**9
V2D = ["avC2_v_rho", "avC2_rho_v", "avC2_v_v"]
V1D = ["rrCvvAv", "rrCvcmvcmAv"]


def _check_velocity(got, raw, label):
    for k in V2D:
        x, y = got[k], raw[k]
        tol = 1e-6 * np.abs(y) + 1e-9 * np.abs(y).max()       # sums over axes and particles that largely cancel
        assert (np.abs(x - y) <= tol).all(), (label, k, float(np.abs(x - y).max()), float(np.abs(y).max()))
        assert np.abs(y).max() > 0
    for k in V1D:
        np.testing.assert_allclose(got[k], raw[k], rtol=1e-9, atol=1e-9 * np.abs(raw[k]).max(), err_msg=label + k)


@pytest.mark.parametrize("S,T,B,ncell,n_atoms,frames,want", [
    (5, 8, 4, 2, 32, 40, "self_ring_vec_kernel"),          # the committed -V fixture's shape
    (30, 10, 10, 5, 500, 35, "self_ring_vec_kernel"),      # in.lmp2sqt.150's nCSpatial
    (43, 6, 6, 4, 200, 20, "self_ring_kernel<C=43"),       # recurrence kernel: positions ring slot-minor, velocity ring particle-minor
    (70, 8, 8, 5, 300, 30, "self_ring_mma_kernel"),
    (171, 6, 6, 8, 2000, 16, "self_ring_mma_kernel"),      # config 5's nCSpatial, several particle tiles
    (256, 4, 4, 3, 100, 12, "self_ring_mma_kernel"),       # the largest nCSpatial -V takes
])
def test_velocity_correlators_match_oracle(S, T, B, ncell, n_atoms, frames, want):
    traj = synth.make_trajectory(n_cells=ncell, n_atoms=n_atoms, n_frames=frames, seed=40 + S, box_jitter=0.004)
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, "FtV")
    assert eng.kernel_config.startswith(want) and "velocity-weighted" in eng.kernel_config, eng.kernel_config
    got = res.to_dict()
    _, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=BIG, flags="FtV", finish=False)
    assert traj.n_atoms == n_atoms
    check_raw_sums(got, raw, n_atoms, label="V S%d " % S)     # everything -V does not touch is unchanged
    _check_velocity(got, raw, "S%d " % S)
    eng.close()


def test_velocity_nCSkip_two_files_and_midrun_flush():
    t1 = synth.make_trajectory(n_cells=3, n_atoms=60, n_frames=37, seed=3, box_jitter=0.01)
    t2 = synth.make_trajectory(n_cells=3, n_atoms=60, n_frames=29, seed=4, box_jitter=0.01)
    S, T, B = 9, 8, 8
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T, nCSkip=1)
    res, eng = run_trajectories([t1, t2], deck, "FtV")
    _, raw, _ = pyoracle.run_port(None, S, T, B, nCSkip=1, limitCorrAv=BIG, files=[t1, t2], flags="FtV", finish=False)
    got = res.to_dict()
    check_raw_sums(got, raw, 60, label="V two files ")
    _check_velocity(got, raw, "two files ")
    eng.close()
    # mid-run flush: the latched block carries the velocity sums too, and the fresh sums restart from zero
    t = synth.make_trajectory(n_cells=2, n_frames=40, seed=1)
    deck = Deck(limitCorrAv=0, nCBuffer=5, nCSpatial=4, nCTime=10)
    flushed = []
    res, eng = run_trajectories([t], deck, "FtV", on_flush=lambda r: flushed.append(r.to_dict()))
    printed, raw, _ = pyoracle.run_port(t, 4, 10, 5, limitCorrAv=0, flags="FtV")
    assert len(flushed) == 1 and flushed[0]["countCorrAv"] == printed["countCorrAv"] == 15
    at_print = printed["_raw_at_print"]
    _check_velocity(flushed[0], at_print, "flush ")
    res.normalise(deck, "FtV")
    g = res.to_dict()
    for k in V2D + V1D:                                        # prePrintProcess' scaling of the -V arrays (lmp2sqt.c:607-611, 650-653)
        np.testing.assert_allclose(g[k], printed[k], rtol=1e-6, atol=1e-9 * np.abs(printed[k]).max(), err_msg=k)
    eng.close()


def test_velocity_requires_velocities_and_supported_nCSpatial():
    from lmp2sqt_b200 import SqtError
    with pytest.raises(SqtError, match="-V / -M need nCSpatial"):
        SpacetimeCorr(Deck(limitCorrAv=BIG, nCBuffer=2, nCSpatial=300, nCTime=4), "FtV")
    eng = SpacetimeCorr(Deck(limitCorrAv=BIG, nCBuffer=2, nCSpatial=4, nCTime=4), "FtV")
    eng.InitSpacetimeCorr()
    with pytest.raises(SqtError, match="velocities"):
        eng.EvalSpacetimeCorr(np.random.rand(10, 3), 0.0, 5.0)
    eng.close()
    with pytest.raises(SqtError, match="beyond what the kernels cover"):
        SpacetimeCorr(Deck(limitCorrAv=BIG, nCBuffer=2, nCSpatial=1025, nCTime=4), "Ft")


ALL = ["avF", "StdDevF", "avF_s", "StdDevF_s", "avF_d", "StdDevF_d", "avDrTable", "rrMSDAv", "rrMQDAv", "rrMSDCMAv"]
ENG = {"avF": "avF_qq2", "StdDevF": "StdDevF_qq2", "avF_s": "avF_s_qq2", "StdDevF_s": "StdDevF_s_qq2", "avF_d": "avF_d_qq2",
       "StdDevF_d": "StdDevF_d_qq2"}


@pytest.mark.parametrize("S,ncell,n_atoms", [(6, 2, 32), (43, 3, 90), (70, 3, 100), (171, 5, 400)])
@pytest.mark.parametrize("flags", ["F", "f", "t", "ft"])
def test_flag_combinations_match_oracle(flags, S, ncell, n_atoms):
    """Without -t nothing of EvalTwoTimeEach exists (MSD, MQD, MSDCM, van Hove table all zero); without -F no F_s / F_d;
    without -f no k-space at all.  The engine also only *runs* what is asked for."""
    T, B, frames = 6, 3, 26
    traj = synth.make_trajectory(n_cells=ncell, n_atoms=n_atoms, n_frames=frames, seed=70 + S)
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, flags)
    got = res.to_dict()
    _, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=BIG, flags=flags, finish=False)
    assert got["countCorrAv"] == raw["countCorrAv"] > 0
    for k in ALL:
        x, y = got[ENG.get(k, k)], raw[k]
        if not y.any():
            assert not x.any(), (flags, k)
            continue
        if k == "avDrTable":
            assert np.array_equal(x, y)
        elif k.startswith("rr"):
            np.testing.assert_allclose(x, y, rtol=1e-7 if k == "rrMSDCMAv" else 1e-9, atol=1e-18, err_msg=k)
        else:
            tol = 1e-6 * np.maximum(np.abs(y), 1e-3 * np.abs(y[:, :1])) + 1e-300
            assert (np.abs(x - y) <= tol).all(), (flags, k)
    cfg = eng.kernel_config
    if "F" not in flags and "t" not in flags:
        assert cfg.startswith("no per-particle kernel"), cfg
    elif "F" not in flags:
        assert cfg.startswith("self_ring_vec_kernel<C=4") and "MSD / van Hove only" in cfg, cfg   # one chain, whatever nCSpatial is
    eng.close()
