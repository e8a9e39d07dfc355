"""Text-dump readers against the reference's own ReadDump (snapshot.c:166-382), called in the unmodified reference
library on randomly damaged dumps: the same frames are delivered (and the file is given up at the same frame), with the
same timestep, atom count, box and coordinates -- by the stream reader (sqt_dump.c), by the mmap reader (sqt_reader.c,
one-pass and count-pass-then-parse) with one and several threads.  Needs oracle/_ref; a frozen sample with the
reference's recorded answers runs everywhere."""
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

HERE = os.path.dirname(os.path.abspath(__file__))
FROZEN = os.path.join(HERE, "golden", "dump_fuzz.pkl")
libc = C.CDLL(None)
libc.fopen.restype = C.c_void_p
libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
libc.fclose.argtypes = [C.c_void_p]

HEADERS = ["ITEM: ATOMS id type xu yu zu", "ITEM: ATOMS id type xu yu zu vx vy vz", "ITEM: ATOMS id type xu yu zu mux muy muz",
           "ITEM: ATOMS id type xu yu zu mux muy muz vx vy vz", "ITEM: ATOMS id type x y z", "ITEM: ATOMS id type xu yu zu "]
BOXES = ["ITEM: BOX BOUNDS pp pp pp", "ITEM: BOX BOUNDS pp pp ff", "ITEM: BOX BOUNDS ff ff pp", "ITEM: BOX BOUNDS ff ff ff",
         "ITEM: BOX BOUNDS pp pp pp extra", "ITEM: BOX BOUNDS xy xz yz pp pp pp"]


def make_dumps(n, seed=11):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        frames = int(rng.integers(1, 5))
        natoms = int(rng.integers(1, 9))
        lines = []
        for f in range(frames):
            hdr = HEADERS[int(rng.integers(0, len(HEADERS)))] if rng.random() < 0.35 else HEADERS[0]
            ncol = {0: 5, 1: 8, 2: 8, 3: 11, 4: 5, 5: 5}[HEADERS.index(hdr)]
            box = BOXES[int(rng.integers(0, len(BOXES)))] if rng.random() < 0.25 else BOXES[0]
            lines += ["ITEM: TIMESTEP" if rng.random() > 0.03 else "ITEM: TIMESTAMP", str(f * 100) if rng.random() > 0.05 else "12x",
                      "ITEM: NUMBER OF ATOMS" if rng.random() > 0.03 else "ITEM: NUMBER", str(natoms),
                      box, "0.000000 16.796000", "-1.5 16.796 0.0", "0 1.6796e1", hdr]
            for i in range(natoms):
                vals = ["%.6f" % v for v in rng.normal(0, 20, ncol - 2)]
                r = rng.random()
                if r < 0.06:
                    vals = vals[:int(rng.integers(0, len(vals)))]          # short line
                elif r < 0.10:
                    vals[0] = "%.3e" % float(vals[0])
                elif r < 0.13:
                    vals[1] = "nan" if rng.random() < 0.5 else "1.5abc"
                sep = " " if rng.random() > 0.08 else rng.choice(["  ", "\t", " \t"])
                ln = ("%d%s1%s" % (i + 1, sep, sep)) + sep.join(vals)
                if rng.random() < 0.03:
                    ln = ""                                                 # blank line inside the atom block
                if rng.random() < 0.03:
                    ln += "\r"
                lines.append(ln)
            if rng.random() < 0.04:
                lines.pop()                                                 # one atom line too few
            if rng.random() < 0.04:
                lines.append("stray line")
        text = "\n".join(lines)
        if rng.random() < 0.85:
            text += "\n"
        if rng.random() < 0.08:
            text = text[:int(rng.integers(0, len(text) + 1))]               # truncated file
        out.append(text)
    return out


def _pack(ts, n, box, pbc, xyz):
    return (int(ts), int(n), tuple(float(b) for b in box), int(pbc), np.array(xyz, dtype=np.float64))


def ours_stream(path):
    h = _lib.host_lib()
    fp = libc.fopen(path.encode(), b"r")
    frames, hd = [], SqtFrameHeader()
    while h.sqt_dump_read_header(fp, C.byref(hd)):
        if hd.n_atoms <= 0:
            break
        xyz = np.zeros((3, hd.n_atoms))
        if not h.sqt_dump_read_atoms(fp, C.byref(hd), xyz[0].ctypes.data, xyz[1].ctypes.data, xyz[2].ctypes.data):
            break
        frames.append(_pack(hd.timestep, hd.n_atoms, (hd.xlo, hd.xhi, hd.ylo, hd.yhi, hd.zlo, hd.zhi), hd.pbcType, xyz.T))
    libc.fclose(fp)
    return frames


def ours_mmap(path, threads, two_pass):
    h = _lib.host_lib()
    r = h.sqt_reader_open(path.encode(), threads)
    if not r:
        return []
    hd = SqtFrameHeader()
    if two_pass:
        while h.sqt_reader_next_header(r, C.byref(hd)) and hd.n_atoms > 0 and h.sqt_reader_skip_atoms(r, C.byref(hd)):
            pass
        h.sqt_reader_rewind(r)
    frames = []
    while h.sqt_reader_next_header(r, C.byref(hd)):
        if hd.n_atoms <= 0:
            break
        xyz = np.zeros((3, hd.n_atoms))
        if not h.sqt_reader_read_atoms(r, C.byref(hd), xyz[0].ctypes.data, xyz[1].ctypes.data, xyz[2].ctypes.data):
            break
        frames.append(_pack(hd.timestep, hd.n_atoms, (hd.xlo, hd.xhi, hd.ylo, hd.yhi, hd.zlo, hd.zhi), hd.pbcType, xyz.T))
    h.sqt_reader_close(r)
    return frames


def _reference_worker(job):
    with open(job, "rb") as f:
        paths = pickle.load(f)
    L = C.CDLL(pyoracle.REF_SO)
    L.ReadDump.restype = C.POINTER(pyoracle._Snapshot)
    L.ReadDump.argtypes = [C.c_void_p]
    res = []
    for p in paths:
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
            b = s.Box
            frames.append(_pack(s.timestep, s.NumAtoms, (b.xlow, b.xhigh, b.ylow, b.yhigh, b.zlow, b.zhigh), b.pbcTYPE,
                                np.column_stack([a["x"], a["y"], a["z"]])))
        libc.fclose(fp)
        res.append(frames)
    with open(job + ".out", "wb") as f:
        pickle.dump(res, f)


def reference(paths, tmp):
    job = os.path.join(tmp, "job.pkl")
    with open(job, "wb") as f:
        pickle.dump(paths, f)
    code = "import sys; sys.path[:0] = [%r, %r, %r]; import test_dump_fuzz as t; t._reference_worker(%r)" % (
        HERE, os.path.join(os.path.dirname(HERE), "oracle"), os.path.dirname(HERE), job)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    with open(job + ".out", "rb") as f:
        return pickle.load(f)


def _same(a, b):
    if len(a) != len(b):
        return False
    for x, y in zip(a, b):
        if x[:4] != y[:4] or not np.array_equal(x[4], y[4], equal_nan=True):
            return False
    return True


def _write(decks, tmp):
    paths = []
    for i, t in enumerate(decks):
        p = os.path.join(tmp, "d%d.lammpstrj" % i)
        with open(p, "w", newline="") as f:
            f.write(t)
        paths.append(p)
    return paths


def _check_all(paths, ref, texts):
    for p, want, t in zip(paths, ref, texts):
        for name, got in (("stream", ours_stream(p)), ("mmap", ours_mmap(p, 1, False)), ("mmap 3 threads", ours_mmap(p, 3, False)),
                          ("mmap two-pass", ours_mmap(p, 2, True))):
            assert _same(got, want), (name, t, [(g[0], g[1]) for g in got], [(w[0], w[1]) for w in want])


@pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")
def test_readers_agree_with_reference_readdump_on_damaged_dumps(tmp_path, monkeypatch):
    monkeypatch.setenv("SQT_READER_LINES_PER_THREAD", "2")
    texts = make_dumps(300)
    paths = _write(texts, str(tmp_path))
    _check_all(paths, reference(paths, str(tmp_path)), texts)


def test_readers_on_frozen_sample(tmp_path, monkeypatch):
    monkeypatch.setenv("SQT_READER_LINES_PER_THREAD", "2")
    with open(FROZEN, "rb") as f:
        rec = pickle.load(f)
    texts = make_dumps(300)[:len(rec)]
    _check_all(_write(texts, str(tmp_path)), rec, texts)


if __name__ == "__main__":                                           # regenerate the frozen sample (needs oracle/_ref)
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        texts = make_dumps(300)[:80]
        with open(FROZEN, "wb") as f:
            pickle.dump(reference(_write(texts, td), td), f)
