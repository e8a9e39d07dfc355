"""TEST INFRASTRUCTURE ONLY: numpy statement of the per-type partial correlators.

PARITY UNPINNED UPSTREAM: the reference has no per-type code at all (``type`` is parsed at snapshot.c:362,368 and never
used).  What is stated here is the textbook definition on the reference's own k-set, ring and accumulation rules:

    rho_a(k, t)   = sum over atoms n of species a of exp(i k . r_n(t)),  k = m k0 e_axis, m = 1 .. nCSpatial, k0 = 2 pi / L(t)
    F_ab(m, lag)  = sum over axes of Re[rho_a(t0 + lag) conj(rho_b(t0))]        (lmp2sqt.c:1584-1586 with rho -> rho_a, rho_b)
    F_s,a(m, lag) = sum over axes and atoms of species a of cos(m (k0 x(t) - k0' x(t0)))   (:1531-1535 restricted to a)
    MSD_a, MQD_a, van Hove_a: lmp2sqt.c:1368-1460 restricted to a

with the ring of time origins of lmp2sqt.c:446-475, 1638-1647 and raw sums over completed buffers like AccumSpacetimeCorr
(:376-443).  The statement is anchored to the reference through identities the tests check: with every atom in one species it
reproduces the pinned restatement (oracle/sqt_oracle.c, bit-identical to the reference) to rounding, and
sum_ab F_ab = F, sum_a F_s,a = F_s, sum_a MSD_a = MSD, sum_a hist_a = hist hold for any species assignment.
Trigonometric functions are evaluated directly (the reference's SLOW_FULL_MATH branch, :1198-1202), which the survey
measured to agree with the recurrence to 5e-16.
"""
from __future__ import annotations

import numpy as np


def run_types(trajs, types, S, T, B, nCSkip=0, rVal=0.2, n_types=None):
    """``trajs``: list of Trajectory (input files); ``types``: [n] ints 1..n_types (``n_types`` defaults to the largest species
    present).  Returns raw sums (dict)."""
    types = np.asarray(types)
    nt = int(n_types) if n_types else int(types.max())
    B = max(1, min(B, T))
    masks = [types == a + 1 for a in range(nt)]
    m = np.arange(1, S + 1)
    out = {"avF_ab": np.zeros((nt, nt, S, T)), "avF_s_a": np.zeros((nt, S, T)), "avDrTable_a": np.zeros((nt, S, T)),
           "rrMSD_a": np.zeros((nt, T)), "rrMQD_a": np.zeros((nt, T)), "countCorrAv": 0,
           "nPtls_a": np.array([int(k.sum()) for k in masks])}
    bufs = [dict(Fab=np.zeros((nt, nt, S, T)), Fs=np.zeros((nt, S, T)), hist=np.zeros((nt, S, T)), msd=np.zeros((nt, T)),
                 mqd=np.zeros((nt, T))) for _ in range(B)]
    for t in trajs:
        count = [-((nb * T) // B) for nb in range(B)]   # C: (-nb * T) / B truncates toward zero (lmp2sqt.c:470)
        skip = 0
        for f in range(t.n_frames):
            if skip < nCSkip:
                skip += 1
                continue
            skip = 0
            pos = t.positions[f]
            L = t.box[f, 0, 1] - t.box[f, 0, 0]
            k0 = 2.0 * np.pi / L
            ang = (k0 * pos)[:, :, None] * m[None, None, :]              # [n][axis][m]
            c, s = np.cos(ang), np.sin(ang)
            rho = np.array([[c[k].sum(axis=0), s[k].sum(axis=0)] for k in masks])   # [a][re|im][axis][m]
            for nb in range(B):
                b = bufs[nb]
                if count[nb] == 0:
                    b["c0"], b["s0"], b["pos0"], b["rho0"] = c, s, pos, rho
                if count[nb] >= 0:
                    lag = count[nb]
                    self_nm = (c * b["c0"] + s * b["s0"]).sum(axis=1)    # [n][m]
                    d = pos - b["pos0"]
                    dr2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
                    bins = np.floor(np.sqrt(dr2) / rVal).astype(np.int64)
                    for a, k in enumerate(masks):
                        b["Fs"][a, :, lag] = self_nm[k].sum(axis=0)
                        b["msd"][a, lag] = dr2[k].sum()
                        b["mqd"][a, lag] = (dr2[k] * dr2[k]).sum()
                        bb = bins[k]
                        b["hist"][a, :, lag] = np.bincount(bb[bb < S], minlength=S)[:S]
                        for a2 in range(nt):
                            b["Fab"][a, a2, :, lag] = (rho[a, 0] * b["rho0"][a2, 0] + rho[a, 1] * b["rho0"][a2, 1]).sum(axis=0)
                count[nb] += 1
            for nb in range(B):
                if count[nb] == T:
                    b = bufs[nb]
                    out["avF_ab"] += b["Fab"]; out["avF_s_a"] += b["Fs"]; out["avDrTable_a"] += b["hist"]
                    out["rrMSD_a"] += b["msd"]; out["rrMQD_a"] += b["mqd"]
                    count[nb] = 0
                    out["countCorrAv"] += 1
    return out
