"""Loading the committed golden cases (tests/golden/<case>/, produced by the unmodified reference:
tests/golden/make_golden.py)."""
import gzip
import io
import json
import os

import numpy as np

from lmp2sqt_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(d for d in os.listdir(GOLDEN) if os.path.isfile(os.path.join(GOLDEN, d, "case.json")))


def case_dir(name):
    return os.path.join(GOLDEN, name)


def load_case(name):
    with open(os.path.join(GOLDEN, name, "case.json")) as f:
        c = json.load(f)
    c["arrays"] = dict(np.load(os.path.join(GOLDEN, name, "arrays.npz")))
    return c


def read_dump_text(text: str) -> synth.Trajectory:
    """Minimal pure-Python reader of the LAMMPS text dump (test helper; the product reader is C)."""
    lines = text.split("\n")
    i = 0
    pos, box, ts, vel, mu = [], [], [], [], []
    while i < len(lines) and lines[i].startswith("ITEM: TIMESTEP"):
        ts.append(int(lines[i + 1]))
        n = int(lines[i + 3])
        b = [[float(v) for v in lines[i + 5 + a].split()[:2]] for a in range(3)]
        cols = lines[i + 8].split()[2:]             # "ITEM: ATOMS id type xu yu zu [mux muy muz] [vx vy vz]"
        arr = np.loadtxt(io.StringIO("\n".join(lines[i + 9:i + 9 + n])), ndmin=2)
        pos.append(arr[:, 2:5])
        if "mux" in cols:
            mu.append(arr[:, cols.index("mux"):cols.index("mux") + 3])
        if "vx" in cols:
            vel.append(arr[:, cols.index("vx"):cols.index("vx") + 3])
        box.append(b)
        i += 9 + n
    t = synth.Trajectory(np.array(pos), np.array(box), np.array(ts), np.ones(pos[0].shape[0], dtype=np.int32),
                         np.array(vel) if len(vel) == len(pos) else None)
    if len(mu) == len(pos):
        t._dipoles_raw = np.array(mu)              # Trajectory.dipoles normalises them like the reference's reader
    return t


def case_trajectories(name):
    d = case_dir(name)
    out = []
    for fn in sorted(os.listdir(d)):
        if fn.startswith("dump") and fn.endswith(".gz"):
            with gzip.open(os.path.join(d, fn), "rt") as f:
                out.append(read_dump_text(f.read()))
    return out


def unpack_dumps(name, dst):
    d = case_dir(name)
    names = []
    for fn in sorted(os.listdir(d)):
        if fn.startswith("dump") and fn.endswith(".gz"):
            with gzip.open(os.path.join(d, fn), "rb") as f, open(os.path.join(dst, fn[:-3]), "wb") as g:
                g.write(f.read())
            names.append(fn[:-3])
    return names


def dt_rows_without_uninitialised_columns(text: str):
    """Dt*.info rows with the msdx/msdy/msdz columns (2..4) dropped: without `-s` the reference prints uninitialised
    memory there (lmp2sqt.c:1060-1072; zeros in most runs, garbage in e.g. golden case g7), this implementation prints 0."""
    lines = text.strip().split("\n")
    out = [lines[0]]
    for ln in lines[1:]:
        f = ln.split()
        out.append(" ".join(f[:2] + f[5:]))
    return out
