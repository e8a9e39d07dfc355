"""Synthetic LJ-liquid-like trajectories in the LAMMPS text-dump format the reference reads.

The reference reader (reference snapshot.c:166-382) accepts exactly this layout::

    ITEM: TIMESTEP
    <step>
    ITEM: NUMBER OF ATOMS
    <N>
    ITEM: BOX BOUNDS pp pp pp
    <xlo> <xhi>
    <ylo> <yhi>
    <zlo> <zhi>
    ITEM: ATOMS id type xu yu zu
    <id> <type> <xu> <yu> <zu>          (N lines, single-space separated, snapshot.c:22)

Atoms are identified by *line order* (``id`` is parsed and ignored, snapshot.c:357-368), so
every frame lists the atoms in the same order.

Workload shape: fcc lattice ``N = 4 n^3`` at reduced density 0.8442 in a
cubic box ``L = (N / 0.8442)^(1/3)``, Gaussian thermal jitter ``sigma0`` per axis, then an
*unwrapped* Gaussian random walk with ``sigma_step`` per axis per frame.
"""
from __future__ import annotations

import io
import os
from dataclasses import dataclass

import numpy as np

RHO_STAR = 0.8442
SIGMA0 = 0.10
SIGMA_STEP = 0.077
DEFAULT_SEED = 20261019


@dataclass
class Trajectory:
    """positions[f, n, 3] (float64, unwrapped), box bounds per frame [f, 3, 2], timesteps[f]."""

    positions: np.ndarray
    box: np.ndarray
    timesteps: np.ndarray
    types: np.ndarray
    velocities: np.ndarray | None = None   # [f, n, 3]; only dumps written with ``velocities=True`` carry them
    _dipoles_raw: np.ndarray | None = None

    @property
    def dipoles_raw(self) -> np.ndarray:
        """[f, n, 3] dipole vectors as a dump carries them (``mux muy muz`` at ``%.6f``, length about 1.3): a slow rotational
        random walk per particle; made on first use from a generator of their own, so the positions and velocities of a
        seed never depend on whether the dipoles were asked for."""
        if self._dipoles_raw is None:
            rng = np.random.Generator(np.random.PCG64(977 + 31 * self.n_atoms + self.n_frames))
            f, n = self.n_frames, self.n_atoms
            mu = np.empty((f, n, 3))
            m = rng.normal(0.0, 1.0, size=(n, 3))
            for i in range(f):
                m = m / np.linalg.norm(m, axis=1, keepdims=True)
                mu[i] = 1.3 * m
                m = m + 0.25 * rng.normal(0.0, 1.0, size=(n, 3))
            self._dipoles_raw = np.round(mu, 6)
        return self._dipoles_raw

    @property
    def dipoles(self) -> np.ndarray:
        """[f, n, 3] what the reference's reader keeps of them: unit vectors, normalised with its own operation order
        (make_atom_dipole, snapshot.c:440-459)."""
        r = self.dipoles_raw
        mu1 = np.sqrt(r[..., 0] * r[..., 0] + r[..., 1] * r[..., 1] + r[..., 2] * r[..., 2])
        return r / mu1[..., None]

    @property
    def n_frames(self) -> int:
        return self.positions.shape[0]

    @property
    def n_atoms(self) -> int:
        return self.positions.shape[1]


def fcc_lattice(n_cells: int, box_len: float) -> np.ndarray:
    """4 n^3 fcc sites in [0, L)^3, ordered cell-major then basis."""
    a = box_len / n_cells
    basis = np.array([[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]])
    idx = np.arange(n_cells)
    cells = np.stack(np.meshgrid(idx, idx, idx, indexing="ij"), axis=-1).reshape(-1, 3)
    sites = (cells[:, None, :] + basis[None, :, :]).reshape(-1, 3)
    return sites * a


def box_length(n_atoms: int, rho: float = RHO_STAR) -> float:
    return float((n_atoms / rho) ** (1.0 / 3.0))


def make_trajectory(
    n_cells: int = 2,
    n_frames: int = 40,
    seed: int = DEFAULT_SEED,
    sigma0: float = SIGMA0,
    sigma_step: float = SIGMA_STEP,
    n_atoms: int | None = None,
    drift: tuple[float, float, float] = (0.0, 0.0, 0.0),
    box_jitter: float = 0.0,
    round_decimals: int | None = 6,
) -> Trajectory:
    """Random-walk liquid.  ``n_atoms`` (optional) truncates the lattice to a ragged count.

    ``box_jitter`` > 0 makes the box length fluctuate per frame (NPT-like) to exercise the
    reference's per-frame ``kVal = 2 pi / L`` (lmp2sqt.c:1620,1632).  ``round_decimals`` rounds
    the coordinates to what the text dump will carry (``%.6f``) so that array and file agree.
    """
    rng = np.random.Generator(np.random.PCG64(seed))
    n_full = 4 * n_cells**3
    L = box_length(n_full)
    base = fcc_lattice(n_cells, L)
    if n_atoms is not None:
        base = base[:n_atoms]
    n = base.shape[0]
    pos0 = base + rng.normal(0.0, sigma0, size=(n, 3))
    steps = rng.normal(0.0, sigma_step, size=(n_frames, n, 3))
    steps[0] = 0.0
    steps += np.asarray(drift)[None, None, :] * (np.arange(n_frames) > 0)[:, None, None]
    pos = pos0[None] + np.cumsum(steps, axis=0)
    box = np.zeros((n_frames, 3, 2))
    if box_jitter > 0:
        Ls = L * (1.0 + box_jitter * rng.standard_normal(n_frames))
    else:
        Ls = np.full(n_frames, L)
    box[:, :, 1] = Ls[:, None]
    # velocities: an Ornstein-Uhlenbeck-like sequence per particle (unit variance, correlation 0.8 per frame), independent
    # of the positions' walk -- only their statistics matter to the -V correlators' tests
    vel = np.empty_like(pos)
    vel[0] = rng.normal(0.0, 1.0, size=(n, 3))
    for f in range(1, n_frames):
        vel[f] = 0.8 * vel[f - 1] + 0.6 * rng.normal(0.0, 1.0, size=(n, 3))
    if round_decimals is not None:
        pos = np.round(pos, round_decimals)
        box = np.round(box, round_decimals)
        vel = np.round(vel, round_decimals)
    return Trajectory(pos, box, np.arange(n_frames, dtype=np.int64) * 100, np.ones(n, dtype=np.int32), vel)


def static_trajectory(n_cells: int, n_frames: int, seed: int = DEFAULT_SEED) -> Trajectory:
    """All frames identical: F_s = 1, F(q,t) = S(q), MSD = 0 (SURVEY.md section 4 KAT)."""
    t = make_trajectory(n_cells, 1, seed)
    return Trajectory(
        np.repeat(t.positions, n_frames, axis=0),
        np.repeat(t.box, n_frames, axis=0),
        np.arange(n_frames, dtype=np.int64) * 100,
        t.types,
    )


def translating_trajectory(n_atoms: int, n_frames: int, d=(0.1, 0.0, 0.0), box_len: float = 10.0,
                           seed: int = DEFAULT_SEED) -> Trajectory:
    """Rigid translation by ``d`` per frame: F_s(q_m, tau) = 1/3 sum_a cos(m k0 d_a tau)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p0 = np.round(rng.uniform(0.0, box_len, size=(n_atoms, 3)), 6)
    pos = p0[None] + np.arange(n_frames)[:, None, None] * np.asarray(d)[None, None, :]
    box = np.zeros((n_frames, 3, 2))
    box[:, :, 1] = box_len
    return Trajectory(np.round(pos, 6), box, np.arange(n_frames, dtype=np.int64),
                      np.ones(n_atoms, dtype=np.int32))


def cubic_lattice_trajectory(n_side: int, n_frames: int, a: float = 1.0) -> Trajectory:
    """Simple-cubic lattice, static: S(q_m) = N if n_side | m else 0."""
    idx = np.arange(n_side)
    p0 = np.stack(np.meshgrid(idx, idx, idx, indexing="ij"), axis=-1).reshape(-1, 3) * a
    pos = np.repeat(p0[None].astype(np.float64), n_frames, axis=0)
    box = np.zeros((n_frames, 3, 2))
    box[:, :, 1] = n_side * a
    return Trajectory(pos, box, np.arange(n_frames, dtype=np.int64), np.ones(len(p0), dtype=np.int32))


def write_text_dump(traj: Trajectory, path_or_file, frames: slice | None = None, velocities: bool = False, dipoles: bool = False) -> None:
    """Write ``traj`` as a LAMMPS text dump (``%.6f`` coordinates, ids 1..N in fixed order); ``velocities`` adds the
    ``vx vy vz`` columns, ``dipoles`` the ``mux muy muz`` columns (before the velocities): the reference's other accepted
    headers (snapshot.c:263-331)."""
    own = isinstance(path_or_file, (str, os.PathLike))
    f = open(path_or_file, "w") if own else path_or_file
    try:
        n = traj.n_atoms
        ids = np.arange(1, n + 1)
        rng = range(traj.n_frames)[frames] if frames is not None else range(traj.n_frames)
        for fi in rng:
            b = traj.box[fi]
            f.write("ITEM: TIMESTEP\n%d\nITEM: NUMBER OF ATOMS\n%d\nITEM: BOX BOUNDS pp pp pp\n" % (traj.timesteps[fi], n))
            for a in range(3):
                f.write("%.6f %.6f\n" % (b[a, 0], b[a, 1]))
            f.write("ITEM: ATOMS id type xu yu zu" + (" mux muy muz" if dipoles else "") + (" vx vy vz" if velocities else "") + "\n")
            buf = io.StringIO()
            cols = [ids, traj.types, traj.positions[fi]]
            if dipoles:
                cols.append(traj.dipoles_raw[fi])
            if velocities:
                cols.append(traj.velocities[fi])
            np.savetxt(buf, np.column_stack(cols), fmt="%d %d" + " %.6f" * (3 * (len(cols) - 2)))
            f.write(buf.getvalue())
    finally:
        if own:
            f.close()


def write_deck(path, kVal=0.0981748, deltaT=0.06, mass=1.0, rVal=0.2, limitCorrAv=0,
               nCBuffer=5, nCSpatial=4, nCTime=10, nCSkip=None) -> None:
    """Write an ``in.lmp2sqt`` deck in the reference's NameList syntax (lmp2sqt.c:2126-2169)."""
    with open(path, "w") as f:
        f.write("NameList -- data\n")
        f.write("kVal\t\t%#g \n" % kVal)
        f.write("deltaT\t\t%#g \n" % deltaT)
        f.write("mass\t\t%#g \n" % mass)
        f.write("rVal\t\t%#g \n" % rVal)
        f.write("limitCorrAv\t%d \n" % limitCorrAv)
        f.write("nCBuffer\t%d \n" % nCBuffer)
        f.write("nCSpatial\t%d \n" % nCSpatial)
        f.write("nCTime\t\t%d \n" % nCTime)
        if nCSkip is not None:
            f.write("nCSkip\t\t%d \n" % nCSkip)
        f.write("----\n")


def write_raw_dump(traj: Trajectory, path: str, frames: slice | None = None) -> None:
    """Write ``traj`` as a raw binary frame file (SQTRAW01, include/sqt_host.h) -- the high-rate alternative to text."""
    import ctypes as C
    from . import _lib
    h = _lib.host_lib()
    rng = range(traj.n_frames)[frames] if frames is not None else range(traj.n_frames)
    first = True
    for fi in rng:
        soa = np.ascontiguousarray(traj.positions[fi].T, dtype=np.float64)
        box = np.ascontiguousarray(traj.box[fi].reshape(6), dtype=np.float64)
        rc = h.sqt_raw_append_frame(os.fsencode(path), int(first), int(traj.timesteps[fi]), traj.n_atoms, box.ctypes.data,
                                    soa[0].ctypes.data, soa[1].ctypes.data, soa[2].ctypes.data)
        if rc:
            raise OSError("sqt_raw_append_frame failed for %s" % path)
        first = False


def write_lammps_binary(traj: Trajectory, path: str, revision: int = 2, chunks: int = 2, frames: slice | None = None,
                        columns: str = "id type xu yu zu", extra_cols: int = 0, triclinic: bool = False) -> None:
    """Write ``traj`` in LAMMPS' native binary dump layout (what ``dump custom file.bin id type xu yu zu`` produces and
    LAMMPS' tools/binary2txt.cpp reads; layout restated in csrc/host/sqt_reader.c).  ``revision`` 0 writes the pre-2020
    layout without magic string and column names; ``chunks`` splits the atoms as a multi-rank run would."""
    import struct
    rng = range(traj.n_frames)[frames] if frames is not None else range(traj.n_frames)
    n = traj.n_atoms
    size_one = 5 + extra_cols
    with open(path, "wb") as f:
        for fi in rng:
            if revision > 0:
                magic = b"DUMPCUSTOM"
                f.write(struct.pack("<q", -len(magic)) + magic + struct.pack("<ii", 1, revision))
            f.write(struct.pack("<qq", int(traj.timesteps[fi]), n))
            f.write(struct.pack("<i6i", int(triclinic), 0, 0, 0, 0, 0, 0))
            b = traj.box[fi]
            f.write(struct.pack("<6d", b[0, 0], b[0, 1], b[1, 0], b[1, 1], b[2, 0], b[2, 1]))
            if triclinic:
                f.write(struct.pack("<3d", 0.0, 0.0, 0.0))
            f.write(struct.pack("<i", size_one))
            if revision > 1:
                units = b"lj"
                cols = (columns + "".join(" c%d" % k for k in range(extra_cols))).encode()
                f.write(struct.pack("<i", len(units)) + units + struct.pack("<b", 1) + struct.pack("<d", 0.005 * traj.timesteps[fi]))
                f.write(struct.pack("<i", len(cols)) + cols)
            rows = np.zeros((n, size_one))
            rows[:, 0] = np.arange(1, n + 1)
            rows[:, 1] = traj.types
            rows[:, 2:5] = traj.positions[fi]
            bounds = np.linspace(0, n, chunks + 1).astype(int)
            f.write(struct.pack("<i", chunks))
            for k in range(chunks):
                blk = np.ascontiguousarray(rows[bounds[k]:bounds[k + 1]], dtype="<f8")
                f.write(struct.pack("<i", blk.size))
                f.write(blk.tobytes())
