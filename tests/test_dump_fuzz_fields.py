"""The velocity and dipole columns of the text-dump reader (sqt_reader_read_atoms_vm, what -V / -s / -M feed the engine with)
against the reference's own ReadDump (snapshot.c:166-382) on random, partly damaged dumps of the three wider styles:
`... vx vy vz` (make_atom_vel), `... mux muy muz` and `... mux muy muz vx vy vz` (both read through make_atom_dipole: the
reference's 11-column branch is shadowed, snapshot.c:262 vs :308), including the unit-vector normalisation of the dipoles
(snapshot.c:440-459) bit for bit.  Needs oracle/_ref; a frozen sample with the reference's recorded answers runs everywhere."""
import ctypes as C
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest

import pyoracle
from lmp2sqt_b200 import _lib
from lmp2sqt_b200._lib import SqtFrameHeader
from test_dump_fuzz import BOXES, HERE, _write, libc

FROZEN = os.path.join(HERE, "golden", "dump_fuzz_fields.pkl")
STYLES = {1: "ITEM: ATOMS id type xu yu zu vx vy vz", 2: "ITEM: ATOMS id type xu yu zu mux muy muz",
          3: "ITEM: ATOMS id type xu yu zu mux muy muz vx vy vz"}


def make_dumps(n, seed=29):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        style = int(rng.integers(1, 4))
        frames, natoms = int(rng.integers(1, 5)), int(rng.integers(1, 9))
        ncol = {1: 8, 2: 8, 3: 11}[style]
        lines = []
        for f in range(frames):
            box = BOXES[int(rng.integers(0, len(BOXES)))] if rng.random() < 0.2 else BOXES[0]
            lines += ["ITEM: TIMESTEP", str(f * 100), "ITEM: NUMBER OF ATOMS", str(natoms), box, "0.000000 16.796000", "-1.5 16.796 0.0",
                      "0 1.6796e1", STYLES[style]]
            for i in range(natoms):
                v = rng.normal(0, 2.0, ncol - 2)
                if style != 1 and rng.random() < 0.1:
                    v[3:6] = rng.normal(0, 3e-4, 3)                         # a dipole shorter than 1e-3: the reference leaves the field unset
                vals = ["%.6f" % x for x in v]
                r = rng.random()
                if r < 0.05:
                    vals = vals[:int(rng.integers(0, len(vals)))]          # short line
                elif r < 0.09:
                    vals[int(rng.integers(0, len(vals)))] = "%.3e" % float(vals[0])
                sep = " " if rng.random() > 0.06 else "  "
                lines.append(("%d%s1%s" % (i + 1, sep, sep)) + sep.join(vals))
            if rng.random() < 0.03:
                lines.pop()
        text = "\n".join(lines) + ("\n" if rng.random() < 0.9 else "")
        if rng.random() < 0.06:
            text = text[:int(rng.integers(0, len(text) + 1))]
        out.append((style, text))
    return out


def ours(path, style, threads):
    h = _lib.host_lib()
    r = h.sqt_reader_open(path.encode(), threads)
    if not r:
        return []
    hd, frames = SqtFrameHeader(), []
    while h.sqt_reader_next_header(r, C.byref(hd)):
        if hd.n_atoms <= 0:
            break
        a = np.zeros((6, hd.n_atoms))
        p = [a[i].ctypes.data for i in range(6)]
        ok = (h.sqt_reader_read_atoms_vm(r, C.byref(hd), p[0], p[1], p[2], p[3], p[4], p[5], None, None, None) if style == 1 else
              h.sqt_reader_read_atoms_vm(r, C.byref(hd), p[0], p[1], p[2], None, None, None, p[3], p[4], p[5]))
        if not ok:
            break
        frames.append((int(hd.timestep), int(hd.n_atoms), a.T.copy()))
    h.sqt_reader_close(r)
    return frames


def _reference_worker(job):
    with open(job, "rb") as f:
        items = pickle.load(f)
    L = C.CDLL(pyoracle.REF_SO)
    L.ReadDump.restype = C.POINTER(pyoracle._Snapshot)
    L.ReadDump.argtypes = [C.c_void_p]
    res = []
    for style, p in items:
        fp = libc.fopen(p.encode(), b"r")
        frames = []
        while True:
            s = L.ReadDump(fp)
            if not s:
                break
            s = s.contents
            if s.NumAtoms <= 0:
                break
            a = np.ctypeslib.as_array(C.cast(s.atoms, C.POINTER(C.c_byte)), shape=(s.NumAtoms * 96,)).view(pyoracle.ATOM_DTYPE)
            ext = np.column_stack([a["vx"], a["vy"], a["vz"]]) if style == 1 else np.column_stack([a["mux"], a["muy"], a["muz"]])
            # (style 2 / 3: a dipole shorter than 1e-3 leaves mux muy muz as malloc gave them; mu1 tells which rows those are)
            unset = np.zeros(s.NumAtoms, bool) if style == 1 else ~(a["mu1"] > 0.001)
            frames.append((int(s.timestep), int(s.NumAtoms), np.column_stack([a["x"], a["y"], a["z"], ext]), unset))
        libc.fclose(fp)
        res.append(frames)
    with open(job + ".out", "wb") as f:
        pickle.dump(res, f)


def reference(items, tmp):
    job = os.path.join(tmp, "job.pkl")
    with open(job, "wb") as f:
        pickle.dump(items, f)
    code = "import sys; sys.path[:0] = [%r, %r, %r]; import test_dump_fuzz_fields as t; t._reference_worker(%r)" % (
        HERE, os.path.join(os.path.dirname(HERE), "oracle"), os.path.dirname(HERE), job)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    with open(job + ".out", "rb") as f:
        return pickle.load(f)


def _check(dumps, paths, ref):
    n_vals = 0
    for (style, text), p, want in zip(dumps, paths, ref):
        for threads in (1, 3):
            got = ours(p, style, threads)
            assert [(g[0], g[1]) for g in got] == [(w[0], w[1]) for w in want], (style, text)
            for g, w in zip(got, want):
                keep = ~w[3]
                assert np.array_equal(g[2][keep], w[2][keep], equal_nan=True), (style, text)      # bit for bit, normalisation included
                assert not g[2][w[3], 3:].any()                                                   # unset in the reference: zero here
                n_vals += int(keep.sum())
    assert n_vals > 100


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
def test_velocity_and_dipole_columns_agree_with_reference_readdump(tmp_path, monkeypatch):
    monkeypatch.setenv("SQT_READER_LINES_PER_THREAD", "2")
    dumps = make_dumps(240)
    paths = _write([t for _, t in dumps], str(tmp_path))
    _check(dumps, paths, reference([(s, p) for (s, _), p in zip(dumps, paths)], str(tmp_path)))


def test_velocity_and_dipole_columns_on_frozen_sample(tmp_path, monkeypatch):
    monkeypatch.setenv("SQT_READER_LINES_PER_THREAD", "2")
    with open(FROZEN, "rb") as f:
        rec = pickle.load(f)
    dumps = make_dumps(240)[:len(rec)]
    _check(dumps, _write([t for _, t in dumps], str(tmp_path)), rec)


if __name__ == "__main__":                                           # regenerate the frozen sample (needs oracle/_ref)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        dumps = make_dumps(240)[:70]
        paths = _write([t for _, t in dumps], td)
        with open(FROZEN, "wb") as f:
            pickle.dump(reference([(s, p) for (s, _), p in zip(dumps, paths)], td), f)
