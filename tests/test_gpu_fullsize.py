"""BASELINE config 5 at its full size on one B200 (16 078 716 atoms, nCSpatial 171 => K = 513, ring 150/150), checked
through size-independent properties -- no CPU oracle can run this shape (the reference's tables need > 1 TB):

  * bookkeeping: countCorrAv and every buffer phase equal the closed form of SURVEY.md section 8a;
  * F_s(q, 0) = 1 and MSD(0) = MQD(0) = 0 exactly, van Hove lag 0 = all particles in bin 0;
  * the on-device source is an unwrapped Gaussian random walk (sigma = 0.077 per axis per frame), for which
    F_s(q, tau) = exp(-q^2 sigma^2 tau / 2), MSD = 3 sigma^2 tau, MQD = 15 sigma^4 tau^2 in expectation
    (statistical error ~ 1/sqrt(3 N) = 1.4e-4), and the van Hove counts of every lag sum to N;
  * F_d = 2 (F - F_s) (the reference's cross-term identity).
Needs ~62 GB of device memory and ~35 s.
"""
import numpy as np
import pytest
import torch

from lmp2sqt_b200 import Deck, SpacetimeCorr, SynthSource

pytestmark = pytest.mark.gpu

N_CELLS, S, T, B = 159, 171, 150, 150
SIGMA = 0.077


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.get_device_properties(0).total_memory < 100e9,
                    reason="needs a GPU with > 100 GB")
def test_config5_full_size_invariants():
    n = 4 * N_CELLS ** 3
    assert n == 16078716
    deck = Deck(limitCorrAv=10**6, nCBuffer=B, nCSpatial=S, nCTime=T)
    eng = SpacetimeCorr(deck, "Ft")
    src = SynthSource(N_CELLS, 0, n)
    eng.InitSpacetimeCorr()
    frames = T + 3                                       # buffers 0..3 complete (buffer nb starts at frame nb, B == T)
    for p in range(frames):
        eng.push_device(src.next(), n, n, 0.0, src.L)
        if p in (0, 1, T - 1, T, frames - 1):            # closed form of the ring phase, every buffer
            done = sum(max(0, (p + 1 - nb) // T) for nb in range(B))
            assert eng.countCorrAv == done, p
            assert eng.buffer_counts() == [((p + 1 - nb) % T if p + 1 > nb else p + 1 - nb) for nb in range(B)], p
    assert "self_ring_mma_kernel" in eng.kernel_config and "rho_mma_kernel" in eng.kernel_config
    res = eng.fetch()
    r = res.to_dict()
    cnt = r["countCorrAv"]
    assert cnt == 4
    L = src.L
    q = (np.arange(S) + 1) * 2 * np.pi / L
    tau = np.arange(T)
    Fs = r["avF_s_qq2"] / (3.0 * n * cnt)                # raw sums: sum over axes, particles, buffers
    F = r["avF_qq2"] / (3.0 * n * cnt)
    msd = r["rrMSDAv"] / (n * cnt)
    mqd = r["rrMQDAv"] / (n * cnt)
    hist = r["avDrTable"]
    # lag 0: exact
    np.testing.assert_allclose(Fs[:, 0], 1.0, rtol=0, atol=1e-12)
    assert msd[0] == 0.0 and mqd[0] == 0.0
    assert hist[0, 0] == n * cnt and hist[1:, 0].sum() == 0
    # Gaussian random walk: expectation values, statistical tolerance
    want = np.exp(-0.5 * (q[:, None] * SIGMA) ** 2 * tau[None, :])
    np.testing.assert_allclose(Fs, want, rtol=0, atol=6e-4)
    np.testing.assert_allclose(msd[1:], 3 * SIGMA ** 2 * tau[1:], rtol=2e-3)
    np.testing.assert_allclose(mqd[1:], 15 * SIGMA ** 4 * tau[1:] ** 2, rtol=5e-3)
    # every particle of every completed buffer lands in exactly one bin (r < nCS * rVal = 34.2 always holds here)
    np.testing.assert_array_equal(hist.sum(axis=0), np.full(T, n * cnt))
    # collective part: S(q) = F(q, 0) is a structure factor (positive), F_d = 2 (F - F_s)
    assert (F[:, 0] > 0).all()
    Fd = r["avF_d_qq2"] / (3.0 * n * cnt)
    np.testing.assert_allclose(Fd, 2 * (F - Fs), rtol=0, atol=1e-12 * np.abs(F).max())
    eng.close()
    src.close()
