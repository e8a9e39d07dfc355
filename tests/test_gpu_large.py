"""GPU parity at scale: the shapes of the BASELINE configs and particle counts large enough that the kernels run with
many particle tiles, several warps per slot group (KW > 1 partial sets), dynamic tile scheduling and the cross-block van Hove
histogram -- everything the 21..108-atom cases of test_gpu_parity.py cannot reach.

* config 1 exactly and a config-2-shaped prefix against raw fp64 sums written once by the UNMODIFIED reference
  (tests/golden/large/*.npz, tests/golden/make_golden_large.py); the seeded trajectory is regenerated here and pinned by
  its sha256;
* >= 8,192 atoms through each self-term kernel family (tensor core nCSpatial 171 / 128 / 64, recurrence kernel nCSpatial 43,
  particle-per-lane vector kernel nCSpatial 30) against the oracle restatement run here on the same frames;
* 1,000,188 atoms (config 3's count): tensor-core kernels == recurrence kernel, two independent implementations.

Tolerances are the contract's (tests/parity.py): 1e-6 relative on F, F_s, F_d with the section-8a floor, lag 0 to 1e-10,
MSD / MQD 1e-9, van Hove counts exact, bookkeeping exact.
"""
import hashlib
import os

import numpy as np
import pytest

import pyoracle
from lmp2sqt_b200 import Deck, SpacetimeCorr, SynthSource, run_trajectories, synth
from parity import check_raw_sums, fd_identity

pytestmark = pytest.mark.gpu
BIG = 10**9
LARGE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "large")


def _digest(t) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(t.positions).tobytes())
    h.update(np.ascontiguousarray(t.box).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("fuse", ["1", "0"])   # two launches per frame (rho(k) inside the vector kernel + vec_finish_kernel) / separate kernels
@pytest.mark.parametrize("name,want_kernel", [("c1_shape", "self_ring_vec_kernel<C=32"), ("c2_shape", "self_ring_vec_kernel<C=12")])
def test_baseline_shapes_match_the_reference(monkeypatch, name, want_kernel, fuse):
    monkeypatch.setenv("SQT_FUSE", fuse)
    """BASELINE config 1 (4,000 atoms, nCSpatial 30, 150/150, 320 frames) and config 2's shape (32,000 atoms, nCSpatial 10,
    150/150, 200-frame prefix): raw sums of the unmodified reference, single thread (lmp2sqt.c:1477-1591, 1368-1460)."""
    g = dict(np.load(os.path.join(LARGE, name + ".npz")))
    n_cells, n_frames, seed, S, T, B = (int(v) for v in g["params"])
    traj = synth.make_trajectory(n_cells=n_cells, n_frames=n_frames, seed=seed)
    assert _digest(traj) == str(g["pos_sha256"]), "the seeded trajectory differs from the one the reference was run on"
    ref = {k[4:]: (v.item() if v.shape == () else v) for k, v in g.items() if k.startswith("raw_")}
    res, eng = run_trajectories([traj], Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T), "Ft")
    assert eng.kernel_config.startswith(want_kernel), eng.kernel_config
    assert ("vec_finish_kernel" in eng.kernel_config) == (fuse == "1"), eng.kernel_config
    got = res.to_dict()
    # the closed form of the ring bookkeeping: 320 frames at 150/150 -> 171 completed buffers
    assert got["countCorrAv"] == ref["countCorrAv"] == sum(max(0, (n_frames - (nb * T) // B) // T) for nb in range(B))
    check_raw_sums(got, ref, traj.n_atoms, label=name + " ")
    fd_identity(got)
    eng.close()


@pytest.mark.parametrize("S,n_atoms,T,B,frames,env,want", [
    (171, 8500, 8, 8, 20, {}, "self_ring_mma_kernel<NT=3"),                       # config 5's kernel variant, ragged last tile
    (171, 8500, 8, 8, 20, {"SQT_TILE": "256"}, "self_ring_mma_kernel<NT=3"),      # the headline's 256-particle tiles
    (128, 8500, 8, 8, 20, {}, "self_ring_mma_kernel<NT=2"),                       # config 4's variant (second column layout)
    (86, 8500, 8, 8, 20, {}, "self_ring_mma_kernel<NT=2"),                        # config 3's variant
    (64, 32768, 8, 8, 20, {}, "self_ring_mma_kernel<NT=1"),                       # more tiles than warps: several tiles per warp
    (43, 8500, 8, 8, 20, {}, "self_ring_kernel<C=43"),                            # recurrence kernel
    (30, 16384, 12, 12, 30, {}, "self_ring_vec_kernel<C=32"),                     # particle-per-lane vector kernel
])
def test_many_tiles_match_oracle(monkeypatch, S, n_atoms, T, B, frames, env, want):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n_cells = int(np.ceil((n_atoms / 4) ** (1 / 3)))
    traj = synth.make_trajectory(n_cells=n_cells, n_atoms=n_atoms, n_frames=frames, seed=101 + S)
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, "Ft")
    cfg = eng.kernel_config
    assert cfg.startswith(want), cfg
    # the point of this test: more than one particle tile and more than one warp per slot group
    part = cfg.split("slot groups x ")[1]
    kw = int(part.split(" warps")[0])
    assert kw > 1, cfg
    got = res.to_dict()
    _, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=BIG, finish=False)
    check_raw_sums(got, raw, n_atoms, label="S%d N%d " % (S, n_atoms))
    fd_identity(got)
    eng.close()


def test_million_atoms_tensor_core_equals_recurrence_kernel(monkeypatch):
    """Config 3's particle count (1,000,188) and nCSpatial (86), ring 6/6, frames from the on-device source: the tensor-core
    kernels and the recurrence kernel are independent implementations (blocked Chebyshev on DMMA vs the reference's own
    recurrence); integer outputs must be identical and the floating sums equal to round-off, far inside the 1e-6 contract."""
    n_cells, S, T, B, frames = 63, 86, 6, 6, 14
    n = 4 * n_cells ** 3
    outs = []
    for mma in ("1", "0"):
        monkeypatch.setenv("SQT_MMA", mma)
        eng = SpacetimeCorr(Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T))
        src = SynthSource(n_cells, 0, n)
        eng.InitSpacetimeCorr()
        for _ in range(frames):
            eng.push_device(src.next(), n, n, 0.0, src.L)
        outs.append(eng.fetch().to_dict())
        assert ("self_ring_mma_kernel" in eng.kernel_config) == (mma == "1")
        eng.close(); src.close()
    a, b = outs
    assert a["countCorrAv"] == b["countCorrAv"] == sum(max(0, (frames - (nb * T) // B) // T) for nb in range(B))
    assert np.array_equal(a["avDrTable"], b["avDrTable"])
    for k in ("avF_qq2", "avF_s_qq2", "avF_d_qq2", "rrMSDAv", "rrMQDAv"):
        ref = np.abs(b[k])
        tol = 1e-9 * np.maximum(ref, 1e-3 * ref[..., :1] if ref.ndim == 2 else 0.0) + 1e-300
        assert (np.abs(a[k] - b[k]) <= tol).all(), (k, float(np.max(np.abs(a[k] - b[k]) / tol)) * 1e-9)


@pytest.mark.parametrize("S,n_atoms,want", [(128, 8500, "self_ring_mma_kernel<NT=2"), (43, 8500, "self_ring_kernel<C=43"), (30, 16384, "self_ring_vec_kernel<C=32")])
def test_many_tiles_velocity_magnet_stress_match_oracle(S, n_atoms, want):
    """-V -M -s at many particle tiles per kernel: the weighted passes of the tensor-core rho(k) kernel (static tiles over
    hundreds of blocks), the velocity ring, the per-axis pass over both ring layouts and sum v (x) r against the restatement."""
    from test_gpu_fields import _check_fields
    T = B = 6
    n_cells = int(np.ceil((n_atoms / 4) ** (1 / 3)))
    traj = synth.make_trajectory(n_cells=n_cells, n_atoms=n_atoms, n_frames=16, seed=211 + S, box_jitter=0.003)
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, "FtVMs")
    assert eng.kernel_config.startswith(want), eng.kernel_config
    got = res.to_dict()
    _, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=BIG, flags="FtVMs", finish=False)
    check_raw_sums(got, raw, n_atoms, label="VMs S%d N%d " % (S, n_atoms))
    _check_fields(got, raw, "FtVMs", "S%d N%d " % (S, n_atoms))
    eng.close()


@pytest.mark.parametrize("S,n_atoms,want", [(128, 8500, "self_ring_mma_kernel<NT=2"), (10, 32000, "self_ring_vec_kernel<C=12")])
def test_many_tiles_per_type_partials_match_the_numpy_statement(S, n_atoms, want):
    """Per-type partials with thousands of atoms per species: species ranges that start in the middle of the dump order,
    several tiles per species pass; config 4's nCSpatial with its 80:20 mixture and config 2's shape with three species."""
    import pytypes
    from test_gpu_types import _check_sum_rules, _check_types
    T = B = 6
    n_cells = int(np.ceil((n_atoms / 4) ** (1 / 3)))
    traj = synth.make_trajectory(n_cells=n_cells, n_atoms=n_atoms, n_frames=16, seed=311 + S)
    nt = 2 if S == 128 else 3
    traj.types = (np.where(np.arange(n_atoms) % 5 == 0, 2, 1) if nt == 2 else 1 + (np.arange(n_atoms) * 7919 % 13) % 3).astype(np.int32)
    deck = Deck(limitCorrAv=BIG, nCBuffer=B, nCSpatial=S, nCTime=T)
    res, eng = run_trajectories([traj], deck, "Ft", n_types=nt)
    assert eng.kernel_config.startswith(want), eng.kernel_config
    got = res.to_dict()
    _check_types(got, pytypes.run_types([traj], traj.types, S, T, B), "S%d N%d " % (S, n_atoms))
    _check_sum_rules(got, "S%d N%d " % (S, n_atoms))
    _, raw, _ = pyoracle.run_port(traj, S, T, B, limitCorrAv=BIG, finish=False)
    check_raw_sums(got, raw, n_atoms, label="typed totals S%d N%d " % (S, n_atoms))
    eng.close()
