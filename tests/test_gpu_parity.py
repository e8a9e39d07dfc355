"""GPU parity tests proper: the CUDA path, called through the C-ABI (ctypes -> libsqt_b200.so), against
the CPU oracle (oracle/sqt_oracle.c, itself pinned bit-for-bit to the unmodified reference by
tests/test_oracle.py) on the same seeded frames."""
import numpy as np
import pytest

import pyoracle
from lmp2sqt_b200 import Deck, run_trajectories, synth
from parity import check_raw_sums, fd_identity

pytestmark = pytest.mark.gpu

BIG = 10**6  # limitCorrAv that is never reached: compare the raw sums of the whole run


def _run_both(trajs, S, T, B, skip=0, flags="Ft", rVal=0.2):
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T, nCSkip=skip, rVal=rVal)
    eng_res, eng = run_trajectories(trajs, deck, flags)
    got = eng_res.to_dict()
    _, raw, log = pyoracle.run_port(None, S, T, B, nCSkip=skip, rVal=rVal, limitCorrAv=BIG, files=trajs, flags=flags, finish=False)
    return got, raw, eng, log


@pytest.mark.parametrize("fuse", ["1", "0"])   # nCSpatial <= 32 on one GPU: fused two-launch frame, or the separate kernels
@pytest.mark.parametrize("S,T,B,ncell,frames", [
    (4, 10, 5, 2, 40),      # the smoke shape
    (10, 30, 30, 3, 100),   # smallQ-like, B == T
    (30, 12, 4, 2, 50),     # in.lmp2sqt.150's nCSpatial (C=30 kernel)
    (7, 9, 9, 2, 31),       # odd sizes, padded chains
    (13, 7, 7, 3, 30),      # one slot per warp (C = 16) with an odd slot count
])
def test_ring_matches_oracle(monkeypatch, S, T, B, ncell, frames, fuse):
    monkeypatch.setenv("SQT_FUSE", fuse)
    traj = synth.make_trajectory(n_cells=ncell, n_frames=frames, seed=11)
    got, raw, eng, _ = _run_both([traj], S, T, B)
    assert ("vec_finish_kernel" in eng.kernel_config) == (fuse == "1"), eng.kernel_config
    check_raw_sums(got, raw, traj.n_atoms, label="S%d T%d B%d " % (S, T, B))
    fd_identity(got)


@pytest.mark.parametrize("S", [43, 64, 86, 100, 128, 171])
def test_large_nCSpatial_lane_groups(S):
    """nCSpatial 86 / 128 / 171 are the BASELINE configs 3-5 (K = 258 / 384 / 513); above 86 the m-range is
    split over lane groups with the recurrence state handed through shuffles."""
    traj = synth.make_trajectory(n_cells=2, n_atoms=29, n_frames=30, seed=5)
    got, raw, eng, _ = _run_both([traj], S, 8, 8)
    check_raw_sums(got, raw, traj.n_atoms, label="S%d " % S)
    fd_identity(got)


def test_ragged_skip_two_files_varbox():
    """Ragged atom count, nCSkip, two input files (ring reset, averages persist), fluctuating box."""
    t1 = synth.make_trajectory(n_cells=3, n_atoms=50, n_frames=37, seed=3, box_jitter=0.01)
    t2 = synth.make_trajectory(n_cells=3, n_atoms=50, n_frames=29, seed=4, box_jitter=0.01)
    got, raw, eng, log = _run_both([t1, t2], 6, 8, 8, skip=1)
    check_raw_sums(got, raw, 50, label="two files ")


def test_bookkeeping_is_exact():
    """countCorrAv and every TBuf.count after every frame equal the reference state machine's."""
    from lmp2sqt_b200 import SpacetimeCorr
    traj = synth.make_trajectory(n_cells=2, n_frames=45, seed=2)
    S, T, B, skip = 3, 10, 4, 1
    eng = SpacetimeCorr(Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T, nCSkip=skip))
    ora = pyoracle.OraclePort(S, T, B, skip, limitCorrAv=BIG)
    eng.InitSpacetimeCorr(); ora.begin_file()
    for f in range(traj.n_frames):
        eng.EvalSpacetimeCorr(traj.positions[f], 0.0, traj.box[f, 0, 1])
        ora.push_frame(traj.positions[f], 0.0, traj.box[f, 0, 1])
        assert eng.countCorrAv == ora.countCorrAv
        assert eng.buffer_counts() == ora.buffer_counts()
    eng.close()


@pytest.mark.parametrize("S,mma", [(12, "0"), (12, "1"), (9, "1"), (64, "1"), (128, "1"), (171, "1"), (171, "0"), (239, "1"), (256, "1")])
def test_rho_matches_oracle_each_frame(monkeypatch, S, mma):
    """rho(k) after every frame, through the classic kernel's virtual slots and through the tensor-core rho kernel."""
    from lmp2sqt_b200 import SpacetimeCorr
    monkeypatch.setenv("SQT_MMA", mma)
    traj = synth.make_trajectory(n_cells=3, n_atoms=101, n_frames=6, seed=9, drift=(0.3, -0.2, 0.1))
    eng = SpacetimeCorr(Deck(limitCorrAv=BIG, nCBuffer=2, nCSpatial=S, nCTime=4))
    ora = pyoracle.OraclePort(S, 4, 2, limitCorrAv=BIG)
    eng.InitSpacetimeCorr(); ora.begin_file()
    for f in range(traj.n_frames):
        eng.EvalSpacetimeCorr(traj.positions[f], 0.0, traj.box[f, 0, 1])
        ora.push_frame(traj.positions[f], 0.0, traj.box[f, 0, 1])
        np.testing.assert_allclose(eng.rho(), ora.rho(), rtol=0, atol=1e-10 * traj.n_atoms)
    assert ("rho_mma_kernel" in eng.kernel_config) == (mma == "1")
    eng.close()


def test_midrun_flush_matches_oracle():
    """limitCorrAv reached mid-run: results latched at that frame, averages zeroed, later buffers accumulate
    into fresh sums, end-of-run print suppressed (lmp2sqt.c:438-440, 366, 327)."""
    traj = synth.make_trajectory(n_cells=2, n_frames=40, seed=1)
    S, T, B = 4, 10, 5
    deck = Deck(limitCorrAv=0, nCBuffer=B, nCSpatial=S, nCTime=T)   # + computed 15
    flushed = []
    res, eng = run_trajectories([traj], deck, "Ft", on_flush=lambda r: flushed.append(r.to_dict()))
    assert len(flushed) == 1 and flushed[0]["countCorrAv"] == 15
    assert eng.number_call_print == 1
    printed, raw, log = pyoracle.run_port(traj, S, T, B, limitCorrAv=0)
    assert printed["countCorrAv"] == 15
    # normalise the engine's latched sums on the host and compare with what the oracle "printed"
    res.normalise(deck, "Ft")
    got = res.to_dict()
    for g, r in [("avF_qq2", "avF"), ("avF_s_qq2", "avF_s"), ("avF_d_qq2", "avF_d"), ("rrMSDAv", "rrMSDAv"), ("rrMQDAv", "rrMQDAv")]:
        np.testing.assert_allclose(got[g], printed[r], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(got["avDrTable"], printed["avDrTable"], rtol=1e-12)
    # one more buffer completed after the flush and sits in the fresh sums
    assert eng.countCorrAv == raw["countCorrAv"] == 1


def test_known_answers_static_and_translation():
    """KATs of SURVEY.md section 4: static trajectory (F_s = 1, MSD = 0, all displacements in bin 0) and rigid
    translation (F_s(q_m, tau) = 1/3 sum_a cos(m k0 d_a tau), MSD = |d|^2 tau^2)."""
    S, T, B = 5, 6, 6
    t = synth.static_trajectory(2, 20)
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([t], deck)
    g = res.to_dict(); cnt = g["countCorrAv"]; N = t.n_atoms
    np.testing.assert_allclose(g["avF_s_qq2"], 3.0 * N * cnt, rtol=1e-13)
    assert np.all(g["rrMSDAv"] == 0) and np.all(g["rrMQDAv"] == 0)
    assert np.all(g["avDrTable"][0] == N * cnt) and np.all(g["avDrTable"][1:] == 0)
    np.testing.assert_allclose(g["avF_qq2"], g["avF_qq2"][:, :1] * np.ones((1, T)), rtol=1e-12)
    d = (0.1, 0.02, 0.0); Lb = 10.0
    t = synth.translating_trajectory(50, 25, d=d, box_len=Lb)
    res, eng = run_trajectories([t], deck)
    g = res.to_dict(); cnt = g["countCorrAv"]; N = 50
    k0 = 2 * np.pi / Lb
    m = np.arange(1, S + 1)[:, None]; tau = np.arange(T)[None, :]
    want = sum(np.cos(m * k0 * da * tau) for da in d) * N * cnt
    np.testing.assert_allclose(g["avF_s_qq2"], want, rtol=1e-9, atol=1e-9 * N * cnt)
    np.testing.assert_allclose(g["rrMSDAv"], (d[0]**2 + d[1]**2) * tau[0]**2 * N * cnt, rtol=1e-9, atol=1e-12)


def test_cubic_lattice_structure_factor():
    """Simple-cubic lattice n^3, L = n a: S(q_m) = N if n | m else 0."""
    n = 4
    t = synth.cubic_lattice_trajectory(n, 12)
    S, T, B = 9, 4, 2
    res, eng = run_trajectories([t], Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T))
    g = res.to_dict(); cnt = g["countCorrAv"]; N = n**3
    Sq = g["avF_qq2"][:, 0] / (3.0 * N * cnt)
    want = np.array([N if (m % n == 0) else 0.0 for m in range(1, S + 1)])
    np.testing.assert_allclose(Sq, want, atol=1e-9 * N)


@pytest.mark.parametrize("S", [200, 700])
def test_very_large_nCSpatial(S):
    """Beyond the BASELINE shapes: 8 and 16 lane groups per warp."""
    traj = synth.make_trajectory(n_cells=2, n_atoms=21, n_frames=14, seed=6)
    got, raw, eng, _ = _run_both([traj], S, 4, 4)
    check_raw_sums(got, raw, traj.n_atoms, label="S%d " % S)


@pytest.mark.parametrize("T,B", [(5, 1), (7, 7), (6, 99), (450 // 30, 450 // 30)])
def test_ring_shapes(T, B):
    """One buffer, B == T, nCBuffer > nCTime (clipped like GetNameList does, lmp2sqt.c:2105)."""
    traj = synth.make_trajectory(n_cells=2, n_frames=40, seed=12)
    got, raw, eng, _ = _run_both([traj], 3, T, B)
    check_raw_sums(got, raw, traj.n_atoms, label="T%d B%d " % (T, B))


def test_device_and_pinned_host_frames_equal_staged_frames():
    """The three ingest paths (engine staging, caller's pinned memory, device-resident frame) give identical sums."""
    import ctypes as C
    import torch
    from lmp2sqt_b200 import SpacetimeCorr
    traj = synth.make_trajectory(n_cells=3, n_frames=30, seed=21)
    n = traj.n_atoms
    deck = Deck(limitCorrAv=BIG, nCBuffer=4, nCSpatial=9, nCTime=8)
    outs = []
    for mode in ("staged", "pinned", "device", "pageable"):
        eng = SpacetimeCorr(deck)
        eng.InitSpacetimeCorr()
        keep = []
        for f in range(traj.n_frames):
            soa = np.ascontiguousarray(traj.positions[f].T)
            L = traj.box[f, 0, 1]
            if mode == "staged":
                eng.EvalSpacetimeCorr(traj.positions[f], 0.0, L)
            elif mode == "pageable":
                eng._check(eng.lib.sqt_push_frame(eng.h, soa.ctypes.data, n, n, 0.0, float(L)), "push")
            elif mode == "pinned":
                t = torch.from_numpy(soa).pin_memory()
                keep.append(t)
                eng._check(eng.lib.sqt_push_frame(eng.h, C.c_void_p(t.data_ptr()), n, n, 0.0, float(L)), "push")
            else:
                t = torch.from_numpy(soa).cuda()
                torch.cuda.synchronize()
                keep.append(t)
                eng.push_device(t.data_ptr(), n, n, 0.0, L)
        outs.append(eng.fetch().to_dict())
        eng.close()
    for o in outs[1:]:
        for k in ("avF_qq2", "avF_s_qq2", "avF_d_qq2", "rrMSDAv", "rrMQDAv", "rrMSDCMAv", "avDrTable"):
            np.testing.assert_allclose(o[k], outs[0][k], rtol=1e-12, atol=1e-12)
        assert o["countCorrAv"] == outs[0]["countCorrAv"]


def test_synth_source_is_shard_independent():
    """The on-device frame source generates the same atoms whatever slice a rank owns (counter-based RNG)."""
    from lmp2sqt_b200 import SynthSource
    full = SynthSource(3, 0, 108)
    a, b = SynthSource(3, 0, 40), SynthSource(3, 40, 68)
    for _ in range(3):
        full.next(); a.next(); b.next()
        f = full.to_host()
        assert np.array_equal(f[:, :40], a.to_host()) and np.array_equal(f[:, 40:], b.to_host())
    assert abs(full.L - (108 / 0.8442) ** (1 / 3)) < 1e-12
    step = full.to_host().copy()
    full.next()
    d = full.to_host() - step
    assert 0.05 < d.std() < 0.11          # sigma_step = 0.077 per axis per frame


def test_atom_count_change_is_an_error():
    from lmp2sqt_b200 import SpacetimeCorr, SqtError
    eng = SpacetimeCorr(Deck(limitCorrAv=BIG, nCBuffer=2, nCSpatial=3, nCTime=4))
    eng.InitSpacetimeCorr()
    eng.EvalSpacetimeCorr(np.random.rand(10, 3), 0.0, 5.0)
    with pytest.raises(SqtError, match="atom count changed"):
        eng.EvalSpacetimeCorr(np.random.rand(11, 3), 0.0, 5.0)
    with pytest.raises(SqtError, match="xhi must exceed xlo"):
        eng.EvalSpacetimeCorr(np.random.rand(10, 3), 1.0, 1.0)
    eng.close()


def test_deterministic_mode_is_bit_reproducible(monkeypatch):
    monkeypatch.setenv("SQT_DETERMINISTIC", "1")
    traj = synth.make_trajectory(n_cells=4, n_frames=20, seed=33)
    deck = Deck(limitCorrAv=BIG, nCBuffer=3, nCSpatial=12, nCTime=6)
    a, _ = run_trajectories([traj], deck)
    b, _ = run_trajectories([traj], deck)
    a, b = a.to_dict(), b.to_dict()
    for k in ("avF_qq2", "avF_s_qq2", "StdDevF_s_qq2", "rrMSDAv", "rrMQDAv"):
        assert np.array_equal(a[k], b[k]), k


# ---- FP64 tensor-core self path (sqt_mma.cu): nCSpatial >= 64 by default, SQT_MMA=1/0 forces it on/off ----------

@pytest.mark.parametrize("S,spw", [(9, 4), (64, 4), (64, 2), (86, 2), (120, 2), (128, 2), (128, 4), (171, 4), (171, 2), (192, 2), (230, 4), (240, 4), (256, 2)])
def test_tensor_core_self_path_matches_oracle(monkeypatch, S, spw):
    """Blocked-Chebyshev self terms on DMMA: every column-tile count (S 9 -> 1 tile ... S 239 -> 4 tiles), every
    slots-per-warp variant, ragged particle count (the last 32-particle tile is partial)."""
    monkeypatch.setenv("SQT_MMA", "1")
    monkeypatch.setenv("SQT_MMA_SPW", str(spw))
    traj = synth.make_trajectory(n_cells=3, n_atoms=77, n_frames=30, seed=15)
    got, raw, eng, _ = _run_both([traj], S, 9, 5)
    assert "mma" in eng.kernel_config, eng.kernel_config
    check_raw_sums(got, raw, traj.n_atoms, label="mma S%d spw%d " % (S, spw))
    fd_identity(got)


def test_tensor_core_path_fluctuating_box_two_files(monkeypatch):
    monkeypatch.setenv("SQT_MMA", "1")
    t1 = synth.make_trajectory(n_cells=3, n_atoms=50, n_frames=37, seed=3, box_jitter=0.01)
    t2 = synth.make_trajectory(n_cells=3, n_atoms=50, n_frames=29, seed=4, box_jitter=0.01)
    got, raw, eng, log = _run_both([t1, t2], 70, 8, 8, skip=1)
    assert "mma" in eng.kernel_config
    check_raw_sums(got, raw, 50, label="mma two files ")


def test_tensor_core_path_agrees_with_classic_kernel(monkeypatch):
    """Same frames through both self-term kernels: integer outputs identical, floating sums within round-off."""
    traj = synth.make_trajectory(n_cells=4, n_frames=24, seed=8)
    deck = Deck(limitCorrAv=BIG, nCBuffer=6, nCSpatial=100, nCTime=6)
    monkeypatch.setenv("SQT_MMA", "0")
    a, ea = run_trajectories([traj], deck)
    assert "mma" not in ea.kernel_config
    monkeypatch.setenv("SQT_MMA", "1")
    b, eb = run_trajectories([traj], deck)
    assert "mma" in eb.kernel_config
    a, b = a.to_dict(), b.to_dict()
    assert np.array_equal(a["avDrTable"], b["avDrTable"])
    assert a["countCorrAv"] == b["countCorrAv"]
    for k in ("avF_qq2", "avF_s_qq2", "rrMSDAv", "rrMQDAv"):
        scale = np.abs(a[k]).max()
        assert np.abs(a[k] - b[k]).max() <= 1e-11 * scale, k


def test_tensor_core_path_deterministic_mode(monkeypatch):
    monkeypatch.setenv("SQT_DETERMINISTIC", "1")
    monkeypatch.setenv("SQT_MMA", "1")
    traj = synth.make_trajectory(n_cells=4, n_frames=20, seed=33)
    deck = Deck(limitCorrAv=BIG, nCBuffer=3, nCSpatial=70, nCTime=6)
    a, _ = run_trajectories([traj], deck)
    b, _ = run_trajectories([traj], deck)
    a, b = a.to_dict(), b.to_dict()
    for k in ("avF_qq2", "avF_s_qq2", "StdDevF_s_qq2", "rrMSDAv", "rrMQDAv"):
        assert np.array_equal(a[k], b[k]), k
