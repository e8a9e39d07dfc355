"""The N>1 path on CPU: world_size-2 gloo processes exercising the host-side logic of the particle-sharded
decomposition: shard bounds, unique-id plumbing through torch.distributed, and the
algebra the GPU ranks rely on -- per-shard self terms / MSD / van Hove counts ADD, rho(k) adds before it is
multiplied (F from the allreduced rho), StdDev squares the all-particle value.  The per-shard numbers come
from the CPU oracle run on each rank's slice of atoms; the reduction goes through torch.distributed (gloo)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    from lmp2sqt_b200 import shard_bounds, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # plumbing the bench uses for the NCCL unique id: rank 0 makes 128 opaque bytes, everyone receives them
    ids = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    assert ids[0] == bytes(range(128))

    S, T, B = 5, 6, 3
    traj = synth.make_trajectory(n_cells=2, n_frames=24, seed=42)
    N = traj.n_atoms
    lo, hi = shard_bounds(N, rank, world)
    ora = pyoracle.OraclePort(S, T, B, limitCorrAv=10**6)
    ora.begin_file()
    # every rank walks the same ring state machine over its own atoms
    rho_hist = []
    for f in range(traj.n_frames):
        ora.push_frame(traj.positions[f, lo:hi], 0.0, traj.box[f, 0, 1])
        r = torch.from_numpy(ora.rho().copy())
        dist.all_reduce(r)                      # per-frame allreduce of the partial rho(k)
        rho_hist.append(r.numpy())
    raw = ora.raw()
    add = {k: torch.from_numpy(raw[k].copy()) for k in ("avF_s", "rrMSDAv", "rrMQDAv", "avDrTable")}
    for v in add.values():
        dist.all_reduce(v)                      # linear quantities: shard sums add
    counts = torch.tensor([raw["countCorrAv"], hi - lo])
    gathered = [torch.zeros_like(counts) for _ in range(world)]
    dist.all_gather(gathered, counts)
    if rank == 0:
        np.savez(out_path, rho=np.array(rho_hist), avF_s=add["avF_s"].numpy(), msd=add["rrMSDAv"].numpy(),
                 mqd=add["rrMQDAv"].numpy(), hist=add["avDrTable"].numpy(), counts=torch.stack(gathered).numpy())
    dist.destroy_process_group()


def test_particle_sharding_algebra_world2(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    from lmp2sqt_b200 import shard_bounds, synth
    world = 2
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    d = np.load(out)
    S, T, B = 5, 6, 3
    traj = synth.make_trajectory(n_cells=2, n_frames=24, seed=42)
    full = pyoracle.OraclePort(S, T, B, limitCorrAv=10**6)
    full.begin_file()
    for f in range(traj.n_frames):
        full.push_frame(traj.positions[f], 0.0, traj.box[f, 0, 1])
        np.testing.assert_allclose(d["rho"][f], full.rho(), atol=1e-11)     # allreduced partial rho == rho of all atoms
    raw = full.raw()
    assert list(d["counts"][:, 0]) == [raw["countCorrAv"]] * world            # identical ring bookkeeping on every rank
    assert int(d["counts"][:, 1].sum()) == traj.n_atoms                       # shards partition the atoms
    np.testing.assert_allclose(d["avF_s"], raw["avF_s"], rtol=1e-12)
    np.testing.assert_allclose(d["msd"], raw["rrMSDAv"], rtol=1e-12)
    np.testing.assert_allclose(d["mqd"], raw["rrMQDAv"], rtol=1e-12)
    assert np.array_equal(d["hist"], raw["avDrTable"])


def test_shard_bounds_cover_and_balance():
    from lmp2sqt_b200 import shard_bounds
    for n in (1, 7, 4000, 16078716):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker_types(rank, world, port, out_path):
    """Per-type partials and the -V / -M / -s fields under particle sharding: what is linear in the particles adds over the
    shards (F_s,a, MSD_a, van Hove_a, per-axis MSD; species counts), what is multiplied is reduced first (rho_a(k), the
    weighted densities kvel / kmu, sum v (x) r)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    import pytypes
    from lmp2sqt_b200 import shard_bounds, synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S, T, B = 4, 6, 3
    traj = synth.make_trajectory(n_cells=2, n_frames=20, seed=43)
    N = traj.n_atoms
    types = (np.arange(N) % 3 == 0).astype(np.int64) + 1
    lo, hi = shard_bounds(N, rank, world)
    sub = synth.Trajectory(traj.positions[:, lo:hi], traj.box, traj.timesteps, traj.types[lo:hi], traj.velocities[:, lo:hi])
    tp = pytypes.run_types([sub], types[lo:hi], S, T, B, n_types=2)
    lin = {k: torch.from_numpy(np.ascontiguousarray(tp[k], dtype=np.float64)) for k in ("avF_s_a", "rrMSD_a", "rrMQD_a", "avDrTable_a")}
    cnt = torch.from_numpy(tp["nPtls_a"].astype(np.float64))
    for v in list(lin.values()) + [cnt]:
        dist.all_reduce(v)
    # fields that are multiplied after the reduction: per frame, this shard's part
    m = np.arange(1, S + 1)
    fields = []
    for f in range(traj.n_frames):
        pos, vel, mu = traj.positions[f, lo:hi], traj.velocities[f, lo:hi], traj.dipoles[f, lo:hi]
        k0 = 2.0 * np.pi / traj.box[f, 0, 1]
        ang = (k0 * pos)[:, :, None] * m[None, None, :]
        c, s = np.cos(ang), np.sin(ang)
        rho_a = np.array([[c[types[lo:hi] == a + 1].sum(axis=0), s[types[lo:hi] == a + 1].sum(axis=0)] for a in range(2)])
        kw = (m * k0)[None, None, :]
        kvel = np.array([(-(kw * vel[:, :, None]) * s).sum(axis=0), ((kw * vel[:, :, None]) * c).sum(axis=0)])      # lmp2sqt.c:1203-1208
        kmu = np.array([(-(kw * mu[:, :, None]) * s).sum(axis=0), ((kw * mu[:, :, None]) * c).sum(axis=0)])         # :1209-1214
        vr = (vel[:, :, None] * pos[:, None, :]).sum(axis=0)                                                          # :1122-1142
        t = torch.from_numpy(np.concatenate([rho_a.ravel(), kvel.ravel(), kmu.ravel(), vr.ravel()]))
        dist.all_reduce(t)
        fields.append(t.numpy())
    if rank == 0:
        np.savez(out_path, fields=np.array(fields), cnt=cnt.numpy(), **{k: v.numpy() for k, v in lin.items()})
    dist.destroy_process_group()


def test_per_type_and_field_sharding_algebra_world2(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pytypes
    from lmp2sqt_b200 import synth
    out = str(tmp_path / "t0.npz")
    mp.spawn(_worker_types, args=(2, _free_port(), out), nprocs=2, join=True)
    d = np.load(out)
    S, T, B = 4, 6, 3
    traj = synth.make_trajectory(n_cells=2, n_frames=20, seed=43)
    N = traj.n_atoms
    types = (np.arange(N) % 3 == 0).astype(np.int64) + 1
    full = pytypes.run_types([traj], types, S, T, B)
    assert list(d["cnt"]) == list(full["nPtls_a"])
    for k in ("avF_s_a", "rrMSD_a", "rrMQD_a"):
        np.testing.assert_allclose(d[k], full[k], rtol=1e-12, atol=1e-12)
    assert np.array_equal(d["avDrTable_a"], full["avDrTable_a"])
    m = np.arange(1, S + 1)
    for f in range(traj.n_frames):                      # the reduced fields equal the all-atom fields
        pos, vel, mu = traj.positions[f], traj.velocities[f], traj.dipoles[f]
        k0 = 2.0 * np.pi / traj.box[f, 0, 1]
        ang = (k0 * pos)[:, :, None] * m[None, None, :]
        c, s = np.cos(ang), np.sin(ang)
        rho_a = np.array([[c[types == a + 1].sum(axis=0), s[types == a + 1].sum(axis=0)] for a in range(2)])
        kw = (m * k0)[None, None, :]
        kvel = np.array([(-(kw * vel[:, :, None]) * s).sum(axis=0), ((kw * vel[:, :, None]) * c).sum(axis=0)])
        kmu = np.array([(-(kw * mu[:, :, None]) * s).sum(axis=0), ((kw * mu[:, :, None]) * c).sum(axis=0)])
        vr = (vel[:, :, None] * pos[:, None, :]).sum(axis=0)
        want = np.concatenate([rho_a.ravel(), kvel.ravel(), kmu.ravel(), vr.ravel()])
        np.testing.assert_allclose(d["fields"][f], want, rtol=1e-11, atol=1e-10)
