"""Shared parity helpers: engine (CUDA, through the C-ABI) vs the CPU oracle, with the tolerances of
SURVEY.md section 8a written out."""
import numpy as np

PAIRS2D = [("avF_qq2", "avF"), ("StdDevF_qq2", "StdDevF"), ("avF_s_qq2", "avF_s"), ("StdDevF_s_qq2", "StdDevF_s"),
           ("avF_d_qq2", "avF_d"), ("StdDevF_d_qq2", "StdDevF_d")]

REL_TOL_F = 1e-6      # F, F_s, F_d: |x - ref| <= 1e-6 * max(|ref|, 1e-3 * |ref(q, tau=0)|)
REL_TOL_MSD = 1e-9    # MSD, MQD vs the validated restatement
TOL_LAG0 = 1e-10      # S(q) = F(q,0) and F_s(q,0) = 1 (relative)


def check_raw_sums(got: dict, ref: dict, n_atoms: int, hist_exact: bool = True, label: str = ""):
    """``got``: engine Results.to_dict() (raw sums); ``ref``: oracle raw dict.  Both are sums over
    countCorrAv buffers of per-buffer values; compare element-wise with the survey's tolerance."""
    assert got["countCorrAv"] == ref["countCorrAv"], (label, got["countCorrAv"], ref["countCorrAv"])
    cnt = max(1, ref["countCorrAv"])
    for g, r in PAIRS2D:
        x, y = got[g], ref[r]
        if g.startswith("StdDev"):
            # sums of squares of O(N) quantities: same relative tolerance on the square scale
            floor_ = 1e-3 * np.abs(y[:, :1])
        else:
            floor_ = 1e-3 * np.abs(y[:, :1])
        tol = REL_TOL_F * np.maximum(np.abs(y), floor_) + 1e-300
        bad = np.abs(x - y) > tol
        assert not bad.any(), "%s %s: max rel err %.3e at %s" % (label, g, float(np.max(np.abs(x - y) / tol) * REL_TOL_F), np.argwhere(bad)[:3])
    # lag 0: S(q) and F_s(q,0) = N per axis -> 3 N cnt
    np.testing.assert_allclose(got["avF_qq2"][:, 0], ref["avF"][:, 0], rtol=TOL_LAG0, atol=1e-9 * 3 * n_atoms * cnt)
    np.testing.assert_allclose(got["avF_s_qq2"][:, 0], 3.0 * n_atoms * cnt, rtol=TOL_LAG0)
    for n in ("rrMSDAv", "rrMQDAv"):
        np.testing.assert_allclose(got[n], ref[n], rtol=REL_TOL_MSD, atol=1e-12, err_msg=label + n)
    np.testing.assert_allclose(got["rrMSDCMAv"], ref["rrMSDCMAv"], rtol=1e-7, atol=1e-18, err_msg=label + "MSDCM")
    if hist_exact:
        assert np.array_equal(got["avDrTable"], ref["avDrTable"]), label + " van Hove counts differ"


def fd_identity(got: dict):
    """F_d = 2 (F - F_s)."""
    lhs = got["avF_d_qq2"]
    rhs = 2.0 * (got["avF_qq2"] - got["avF_s_qq2"])
    np.testing.assert_allclose(lhs, rhs, rtol=1e-12, atol=1e-9 * np.max(np.abs(got["avF_qq2"])))
