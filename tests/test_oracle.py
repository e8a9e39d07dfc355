"""The oracle is only trustworthy if it is pinned: (1) against the committed golden vectors the UNMODIFIED
reference produced (tests/golden/, runs anywhere), (2) against the reference itself, bit for bit, where
oracle/_ref is available (this container; on the GPU box the prebuilt oracle/_ref travels with the repo)."""
import os

import numpy as np
import pytest

import pyoracle
from golden_utils import CASES, case_dir, case_trajectories, load_case
from lmp2sqt_b200 import synth

N2 = pyoracle.NAMES2D
N1 = pyoracle.NAMES1D


@pytest.mark.parametrize("name", CASES)
def test_port_reproduces_reference_golden_arrays(name):
    """Raw fp64 sums and normalised arrays read from the reference's globals: bit-identical."""
    c = load_case(name)
    trajs = case_trajectories(name)
    printed, raw, log = pyoracle.run_port(None, c["S"], c["T"], c["B"], nCSkip=c["skip"] or 0,
                                          limitCorrAv=10**9 - sum((t.n_frames - c["T"]) // (c["T"] // min(c["B"], c["T"])) for t in trajs),
                                          files=trajs, flags=c["flags"], finish=False)
    a = c["arrays"]
    assert raw["countCorrAv"] == int(a["raw_countCorrAv"])
    assert [l[0] for l in log] == list(a["count_log"])
    for k in N2 + N1:
        assert np.array_equal(raw[k], a["raw_" + k]), k


@pytest.mark.parametrize("name", CASES)
def test_port_reproduces_reference_output_binary(name):
    """output.binary written by the reference CLI (exact fp64) vs the restatement driven like main()."""
    c = load_case(name)
    if "f" not in c["flags"].lower():
        pytest.skip("-t alone writes no output.binary (lmp2sqt.c:354)")
    trajs = case_trajectories(name)
    printed, raw, log = pyoracle.run_port(None, c["S"], c["T"], c["B"], nCSkip=c["skip"] or 0, limitCorrAv=c["limit"],
                                          files=trajs, flags=c["flags"])
    ob = pyoracle.parse_output_binary(os.path.join(case_dir(name), "output.binary"), c["S"], c["T"])
    assert ob["magic"] == "##space-time corr" and ob["consumed"] == ob["size"]
    assert np.array_equal(ob["full-density"], printed["avF"].T)
    if "F" in c["flags"]:
        assert np.array_equal(ob["self-density"], printed["avF_s"].T)
        assert np.array_equal(ob["cross-density"], printed["avF_d"].T)
    else:
        assert ob["self-density"] is None and ob["cross-density"] is None
    assert np.array_equal(ob["self-vanHove"], printed["avDrTable"].T)
    assert ob["kVal"] == printed["kVal"]


@pytest.mark.parametrize("name", CASES)
def test_port_matches_printed_msd_digits(name):
    """MSD / MQD exist in the reference's files only at %8.4g (Dt*.info, lmp2sqt.c:1060): the restatement's
    fp64 values must round to exactly those digits."""
    c = load_case(name)
    if "t" not in c["flags"]:
        pytest.skip("no Dt file without -t (lmp2sqt.c:1033)")
    trajs = case_trajectories(name)
    printed, _, _ = pyoracle.run_port(None, c["S"], c["T"], c["B"], nCSkip=c["skip"] or 0, limitCorrAv=c["limit"],
                                      files=trajs, flags=c["flags"])
    rows = [ln.split() for ln in open(os.path.join(case_dir(name), "Dt000.info")).read().strip().split("\n")[1:]]
    for nt, r in enumerate(rows):
        assert r[1] == ("%8.4g" % printed["rrMSDAv"][nt]).strip()
        assert r[7] == ("%8.4g" % printed["rrMQDAv"][nt]).strip()


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("skip,flags", [(0, "Ft"), (2, "Ft"), (0, "ft"), (1, "t")])
def test_port_bit_identical_to_live_reference(skip, flags):
    """Fresh random trajectory (not a committed fixture), two input files, fluctuating box."""
    t1 = synth.make_trajectory(n_cells=2, n_frames=47, box_jitter=0.01, seed=101 + skip)
    t2 = synth.make_trajectory(n_cells=2, n_frames=31, box_jitter=0.01, seed=202 + skip)
    S, T, B = 5, 10, 4
    ref = pyoracle.run_reference([t1, t2], S, T, B, nCSkip=skip, limit=1000, flags=flags)
    extra = sum((t.n_frames - T) // (T // B) for t in (t1, t2))
    pr, raw, log = pyoracle.run_port(None, S, T, B, nCSkip=skip, limitCorrAv=1000 - extra, files=[t1, t2], flags=flags)
    assert ref["raw"]["countCorrAv"] == raw["countCorrAv"] and ref["log"] == [l[0] for l in log]
    for k in N2 + N1:
        assert np.array_equal(ref["raw"][k], raw[k]), k
        assert np.array_equal(ref["printed"][k], pr[k], equal_nan=True), k


def test_known_answer_identities_on_the_oracle():
    """KATs of SURVEY.md section 4 on the checker itself."""
    S, T, B = 5, 6, 6
    t = synth.static_trajectory(2, 20)
    pr, raw, _ = pyoracle.run_port(t, S, T, B, limitCorrAv=10**6)
    N, cnt = t.n_atoms, raw["countCorrAv"]
    assert cnt == sum(max(0, (20 - (nb * T) // B) // T) for nb in range(B))
    np.testing.assert_allclose(pr["avF_s"], 1.0, rtol=1e-12)
    assert np.all(pr["rrMSDAv"] == 0)
    np.testing.assert_allclose(pr["avF_d"], 2 * (pr["avF"] - pr["avF_s"]), atol=1e-12)
    tr = synth.translating_trajectory(50, 40, d=(0.1, 0.0, 0.0), box_len=10.0)
    pr, raw, _ = pyoracle.run_port(tr, 4, 5, 1, limitCorrAv=10**6)
    k0 = 2 * np.pi / 10.0
    m = np.arange(1, 5)[:, None]; tau = np.arange(5)[None, :]
    np.testing.assert_allclose(pr["avF_s"], (np.cos(m * k0 * 0.1 * tau) + 2) / 3, rtol=1e-10)
    np.testing.assert_allclose(pr["rrMSDAv"], 0.01 * tau[0] ** 2, rtol=1e-9)


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("flags", ["FtV", "ftV"])
def test_port_velocity_correlators_bit_identical_to_live_reference(flags):
    """SURVEY.md section 8f.4, oracle first: the restatement of the -V path (current density k v e^{ikr}, its
    correlators C2_v_v / C2_rho_v / C2_v_rho, the velocity autocorrelation Cvv and Cvcmvcm; lmp2sqt.c:1203-1208,
    1397-1457, 1542-1555, 609-613, 640-652) against the unmodified reference run with flag_velocity on the same frames."""
    t1 = synth.make_trajectory(n_cells=2, n_frames=43, box_jitter=0.01, seed=303)
    t2 = synth.make_trajectory(n_cells=2, n_frames=29, box_jitter=0.01, seed=404)
    S, T, B = 6, 8, 4
    ref = pyoracle.run_reference([t1, t2], S, T, B, limit=1000, flags=flags)
    extra = sum((t.n_frames - T) // (T // B) for t in (t1, t2))
    pr, raw, log = pyoracle.run_port(None, S, T, B, limitCorrAv=1000 - extra, files=[t1, t2], flags=flags)
    assert ref["raw"]["countCorrAv"] == raw["countCorrAv"] and ref["log"] == [l[0] for l in log]
    names = N2 + N1 + pyoracle.NAMES2D_V + pyoracle.NAMES1D_V
    for k in names:
        assert np.array_equal(ref["raw"][k], raw[k]), k
        assert np.array_equal(ref["printed"][k], pr[k], equal_nan=True), k
    assert np.abs(raw["avC2_v_v"]).max() > 0 and np.abs(raw["rrCvvAv"]).max() > 0   # the -V path really ran
    # without -V nothing of it leaks into the standard arrays
    pr0, raw0, _ = pyoracle.run_port(None, S, T, B, limitCorrAv=1000 - extra, files=[t1, t2], flags=flags.replace("V", ""))
    for k in N2 + N1:
        assert np.array_equal(raw0[k], raw[k]), k


def test_port_velocity_correlators_reproduce_committed_reference_fixture():
    """The same -V restatement against a COMMITTED fixture (tests/golden/velocity/v1_velocity: written by the unmodified
    reference CLI with -F -t -V on a dump that carries vx vy vz, plus the raw fp64 sums read from its globals), so that
    the check also runs where /root/reference does not exist."""
    import gzip
    import json
    from golden_utils import GOLDEN, read_dump_text
    d = os.path.join(GOLDEN, "velocity", "v1_velocity")
    c = json.load(open(os.path.join(d, "case.json")))
    a = dict(np.load(os.path.join(d, "arrays.npz")))
    with gzip.open(os.path.join(d, "dump0.lammpstrj.gz"), "rt") as f:
        traj = read_dump_text(f.read())
    assert traj.velocities is not None and traj.velocities.shape == traj.positions.shape
    S, T, B = c["S"], c["T"], c["B"]
    extra = (traj.n_frames - T) // (T // B)
    printed, raw, log = pyoracle.run_port(None, S, T, B, limitCorrAv=10**9 - extra, files=[traj], flags=c["flags"], finish=False)
    assert raw["countCorrAv"] == int(a["raw_countCorrAv"]) and [l[0] for l in log] == list(a["count_log"])
    for k in N2 + N1 + pyoracle.NAMES2D_V + pyoracle.NAMES1D_V:
        assert np.array_equal(raw[k], a["raw_" + k]), k
    # what the reference printed: the Dt file's Cvv / Cvcmvcm columns (%8.4f / %8.4g) and the C2 tables (%8.4e)
    printed, _, _ = pyoracle.run_port(None, S, T, B, limitCorrAv=c["limit"], files=[traj], flags=c["flags"])
    rows = [ln.split() for ln in open(os.path.join(d, "Dt000.info")).read().strip().split("\n")[1:]]
    for nt, r in enumerate(rows):
        assert r[8] == ("%8.4f" % printed["rrCvvAv"][nt]).strip() and r[9] == ("%8.4g" % printed["rrCvcmvcmAv"][nt]).strip()
    for fn, key in (("C2_v_v000.info", "avC2_v_v"), ("C2_rho_v000.info", "avC2_rho_v"), ("C2_v_rho000.info", "avC2_v_rho")):
        tab = [ln.split() for ln in open(os.path.join(d, fn)).read().strip().split("\n")]
        assert len(tab) == T
        for nt, r in enumerate(tab):
            assert r[1:] == [("%8.4e" % printed[key][nk, nt]).strip() for nk in range(S)], (fn, nt)


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("seed", range(16))
def test_port_bit_identical_to_live_reference_random_sweep(seed):
    """Random decks (nCSpatial, ring shape incl. nCBuffer > nCTime, nCSkip), flag sets (incl. -V) and file sets (short files
    are skipped like main() does): raw sums, ring bookkeeping per frame and the normalised arrays, bit for bit."""
    rng = np.random.default_rng(7000 + seed)
    S, T = int(rng.integers(1, 9)), int(rng.integers(2, 10))
    B = int(rng.integers(1, T + 3))
    skip = int(rng.integers(0, 3))
    flags = str(rng.choice(["Ft", "ft", "FtV", "ftV"]))
    Bc = max(1, min(B, T))
    trajs = []
    for i in range(int(rng.integers(1, 4))):
        frames = int(rng.choice([T + 1 + T // Bc, 2 * T + 3, 3 * T + 1, 40]))
        trajs.append(synth.make_trajectory(n_cells=int(rng.integers(1, 3)) if i == 0 else 1, n_frames=frames, seed=int(rng.integers(0, 10**6)),
                                           box_jitter=float(rng.choice([0.0, 0.01]))))
    n0 = trajs[0].n_atoms
    trajs = [t for t in trajs if t.n_atoms == n0]                      # one atom count per run
    # only files main() accepts (lmp2sqt.c:221-236): run_port applies that rule, the in-process reference driver does not
    trajs = [t for t in trajs if t.n_frames >= 5 and int((t.n_frames - T) / (T // Bc)) >= 1]
    if not trajs:
        pytest.skip("no acceptable file in this draw")
    extra = sum((t.n_frames - T) // (T // Bc) for t in trajs if t.n_frames >= 5 and (t.n_frames - T) // (T // Bc) >= 1)
    ref = pyoracle.run_reference(trajs, S, T, B, nCSkip=skip, limit=1000, flags=flags)
    pr, raw, log = pyoracle.run_port(None, S, T, B, nCSkip=skip, limitCorrAv=1000 - extra, files=trajs, flags=flags)
    assert ref["raw"]["countCorrAv"] == raw["countCorrAv"] and ref["log"] == [l[0] for l in log]
    names = N2 + N1 + (pyoracle.NAMES2D_V + pyoracle.NAMES1D_V if "V" in flags else [])
    for k in names:
        assert np.array_equal(ref["raw"][k], raw[k]), k
        if raw["countCorrAv"] > 0:
            assert np.array_equal(ref["printed"][k], pr[k], equal_nan=True), k


# ---- -M (flag_magnet) and -s (flag_s): SURVEY.md section 8f.4, the rest of the row ------------------------------------------
FIELD_CASES = ["m1_magnet", "s1_stress", "s2_stress_alone"]


def _extra_names(flags):
    n2, n1 = pyoracle.names_for(flags)
    names = [n for _, n in n2] + [n for _, n in n1]
    if "s" in flags:
        names += ["rrMSR1_R_Av", "rrMSR2_VR_Av"]
    return names


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
@pytest.mark.parametrize("flags", ["FtM", "FtVM", "Fts", "s", "ftMs", "FtVMs", "fM"])
def test_port_magnet_and_stress_bit_identical_to_live_reference(flags):
    """The restatement of -M (kmu_q1 lmp2sqt.c:1209-1214, C2_mu_mu / C2_rho_mu / C2_mu_rho :1565-1576, C2_v_mu / C2_mu_v
    :1556-1564, rrCmm never accumulated :420-432) and of -s (sumVR_ct :1122-1142 -- zeroed under -f only, :1120 --, orgSumVR
    :1273-1275, rrMSR1_R :1434-1436, rrMSR2_VR :1461-1476, scaling :660-669) against the unmodified reference on the same
    frames, two input files."""
    t1 = synth.make_trajectory(n_cells=2, n_frames=43, box_jitter=0.01, seed=505)
    t2 = synth.make_trajectory(n_cells=2, n_frames=29, box_jitter=0.01, seed=606)
    S, T, B = 6, 8, 4
    ref = pyoracle.run_reference([t1, t2], S, T, B, limit=1000, flags=flags)
    extra = sum((t.n_frames - T) // (T // B) for t in (t1, t2))
    pr, raw, log = pyoracle.run_port(None, S, T, B, limitCorrAv=1000 - extra, files=[t1, t2], flags=flags)
    assert ref["raw"]["countCorrAv"] == raw["countCorrAv"] > 0 and ref["log"] == [l[0] for l in log]
    names = _extra_names(flags)
    for k in names:
        assert np.array_equal(ref["raw"][k], raw[k]), k
        assert np.array_equal(ref["printed"][k], pr[k], equal_nan=True), k
    if "s" in flags:
        for k in ("rrMSR2_VR_Av_dig", "rrMSR2_VR_Av_offdig"):
            assert np.array_equal(ref["printed"][k], pr[k]), k
        assert np.abs(raw["rrMSR2_VR_Av"]).max() > 0 and np.abs(raw["rrMSR1_R_Av"]).max() > 0
        np.testing.assert_allclose(raw["rrMSR1_R_Av"].sum(axis=1), raw["rrMSDAv"], rtol=1e-12)   # msdx + msdy + msdz = MSD
    if "M" in flags:
        assert np.abs(raw["avC2_mu_mu"]).max() > 0
        assert not raw["rrCmmAv"].any()                      # the reference's quirk: printed, never summed


@pytest.mark.parametrize("name", FIELD_CASES)
def test_port_magnet_and_stress_reproduce_committed_reference_fixtures(name):
    """The same against COMMITTED fixtures (tests/golden/velocity/: written by the unmodified reference CLI with -F -t -M on a
    dump that carries mux muy muz, and with -F -t -s / -s on one that carries vx vy vz, plus the raw fp64 sums read from its
    globals), so that the check also runs where /root/reference does not exist."""
    import gzip
    import json
    from golden_utils import GOLDEN, read_dump_text
    d = os.path.join(GOLDEN, "velocity", name)
    c = json.load(open(os.path.join(d, "case.json")))
    a = dict(np.load(os.path.join(d, "arrays.npz")))
    with gzip.open(os.path.join(d, "dump0.lammpstrj.gz"), "rt") as f:
        traj = read_dump_text(f.read())
    S, T, B, skip = c["S"], c["T"], c["B"], c["skip"] or 0
    _, raw, log = pyoracle.run_port(None, S, T, B, nCSkip=skip, limitCorrAv=10**8, files=[traj], flags=c["flags"], finish=False)
    assert raw["countCorrAv"] == int(a["raw_countCorrAv"]) and [l[0] for l in log] == list(a["count_log"])
    for k in _extra_names(c["flags"]):
        assert np.array_equal(raw[k], a["raw_" + k]), k
    # what the reference binary printed
    printed, _, _ = pyoracle.run_port(None, S, T, B, nCSkip=skip, limitCorrAv=c["limit"], files=[traj], flags=c["flags"])
    rows = [ln.split() for ln in open(os.path.join(d, "Dt000.info")).read().strip().split("\n")[1:]]
    if "s" in c["flags"]:
        for nt, r in enumerate(rows):
            assert r[2:5] == [("%8.4g" % printed["rrMSR1_R_Av"][nt, j]).strip() for j in range(3)], nt
        tab = [ln.split() for ln in open(os.path.join(d, "Stress000.info")).read().strip().split("\n")[1:]]
        assert len(tab) == T
        for nt, r in enumerate(tab):
            m = printed["rrMSR2_VR_Av"][nt]
            want = [printed["rrMSR2_VR_Av_dig"][nt], printed["rrMSR2_VR_Av_offdig"][nt], m[1], m[3], m[7], m[5], m[2], m[6]]
            assert r[1:] == [("%8.4g" % v).strip() for v in want], nt
    if "M" in c["flags"]:
        assert all(r[8] == "0.0000" for r in rows)               # the Cmm column
        for fn, key in (("C2_mu_mu000.info", "avC2_mu_mu"), ("C2_rho_mu000.info", "avC2_rho_mu"), ("C2_mu_rho000.info", "avC2_mu_rho")):
            tab = [ln.split() for ln in open(os.path.join(d, fn)).read().strip().split("\n")]
            for nt, r in enumerate(tab):
                assert r[1:] == [("%8.4e" % printed[key][nk, nt]).strip() for nk in range(S)], (fn, nt)


def test_per_type_statement_is_anchored_to_the_pinned_restatement():
    """the reference has no per-type code, so the checker of the per-type kernels (oracle/pytypes.py,
    numpy) is anchored instead: with every atom in one species it gives the pinned restatement's numbers (= the reference's)
    to rounding, and for any species assignment the partials sum to them."""
    import pytypes
    t1 = synth.make_trajectory(n_cells=2, n_frames=40, seed=5, box_jitter=0.004)
    t2 = synth.make_trajectory(n_cells=2, n_frames=30, seed=6, box_jitter=0.004)
    S, T, B = 7, 8, 4
    for skip in (0, 1):
        _, raw, _ = pyoracle.run_port(None, S, T, B, nCSkip=skip, limitCorrAv=10**8, files=[t1, t2], flags="Ft", finish=False)
        one = pytypes.run_types([t1, t2], np.ones(32, int), S, T, B, nCSkip=skip)
        assert one["countCorrAv"] == raw["countCorrAv"] > 0
        for a, b in (("avF_ab", "avF"), ("avF_s_a", "avF_s"), ("rrMSD_a", "rrMSDAv"), ("rrMQD_a", "rrMQDAv")):
            np.testing.assert_allclose(one[a].reshape(raw[b].shape), raw[b], rtol=1e-12, atol=1e-12 * np.abs(raw[b]).max(), err_msg=a)
        assert np.array_equal(one["avDrTable_a"][0], raw["avDrTable"])
        ty = (np.arange(32) % 3 == 0).astype(int) + 1
        two = pytypes.run_types([t1, t2], ty, S, T, B, nCSkip=skip)
        np.testing.assert_allclose(two["avF_ab"].sum(axis=(0, 1)), raw["avF"], rtol=1e-12, atol=1e-12 * np.abs(raw["avF"]).max())
        np.testing.assert_allclose(two["avF_s_a"].sum(axis=0), raw["avF_s"], rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(two["rrMSD_a"].sum(axis=0), raw["rrMSDAv"], rtol=1e-12)
        assert np.array_equal(two["avDrTable_a"].sum(axis=0), raw["avDrTable"])
