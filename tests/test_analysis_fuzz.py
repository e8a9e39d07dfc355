"""Post-processor against the unmodified reference binary (oracle/_ref/analysis_binary_ref.out) on random, well-formed
output.binary files: random sizes, one to three sets per file, blocks with and without data, decaying / noisy / partly
negative correlators (the logarithmic derivatives then produce inf and nan), random options.  Every file the reference
writes must come out byte for byte (the sign printed for a nan aside)."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

import pyoracle
from lmp2sqt_b200 import _lib, synth

pytestmark = pytest.mark.skipif(not os.path.exists(pyoracle.REF_ANALYSIS_BIN), reason="oracle/_ref/analysis_binary_ref.out not built")
BLOCKS = ("full-density", "self-density", "cross-density", "self-vanHove")


def _output_binary(rng, S, T, sets, with_self):
    out = b""
    for _ in range(sets):
        tag = b"##space-time corr"
        out += struct.pack("<i", len(tag)) + tag + struct.pack("<i", 4)
        for j, name in enumerate(BLOCKS):
            nb = name.encode()
            out += struct.pack("<i", len(nb)) + nb + struct.pack("<3d", 0.374, 0.06 * int(rng.integers(1, 3)), 0.2)
            has = 1 if (j in (0, 3) or with_self) else 0
            out += struct.pack("<i", has)
            if not has:
                continue
            t = np.arange(T)[:, None]
            k = np.arange(1, S + 1)[None, :]
            kind = rng.integers(0, 4)
            if j == 3:
                a = rng.integers(0, 50, (T, S)).astype(float) * rng.random()
            elif kind == 0:
                a = (1.0 + rng.random()) * np.exp(-0.02 * k * k * t) * (1 + 0.01 * rng.standard_normal((T, S)))
            elif kind == 1:
                a = np.exp(-0.3 * k * t) + 0.05 * rng.standard_normal((T, S))          # dips below zero
            elif kind == 2:
                a = np.cos(0.4 * k * t) * np.exp(-0.05 * t)                            # oscillating
            else:
                a = rng.random((T, S)) * rng.choice([1e-6, 1.0, 1e6])
            out += np.ascontiguousarray(a, dtype="<f8").tobytes()
    return out


def _same_text(got, want, name):
    if got == want:
        return
    g, w = got.split(), want.split()
    assert len(g) == len(w), name
    for a, b in zip(g, w):
        assert a == b or ("nan" in a and "nan" in b), (name, a, b)


@pytest.mark.parametrize("seed", range(30))
def test_analysis_cli_equals_reference_on_random_files(tmp_path, seed):
    rng = np.random.default_rng(500 + seed)
    fourier = rng.random() < 0.4
    # Fourier mode: 2 (T - 1) must be a power of two, and T >= 9 (below, the reference's frequency axis divides by nv = 0: SIGFPE)
    T = int(rng.choice([9, 17, 33])) if fourier else int(rng.integers(3, 20))
    S = int(rng.integers(1, 9))
    args = (["-f"] + (["-w"] if rng.random() < 0.5 else [])) if fourier else \
           (["-d", str(int(rng.integers(0, T + 3)))] if rng.random() < 0.5 else []) + (["-u", "%.3f" % rng.random()] if rng.random() < 0.5 else [])
    src = tmp_path / "src"
    src.mkdir()
    synth.write_deck(str(src / "in.lmp2sqt"), deltaT=0.06, nCBuffer=2, nCSpatial=S, nCTime=T, kVal=0.374)
    (src / "output.binary").write_bytes(_output_binary(rng, S, T, int(rng.integers(1, 4)), rng.random() < 0.7))
    outs = {}
    for tag, exe in (("ref", pyoracle.REF_ANALYSIS_BIN), ("ours", _lib.ANALYSIS_BIN)):
        d = tmp_path / tag
        shutil.copytree(src, d)
        r = subprocess.run([exe, *args, "output.binary"], cwd=d, capture_output=True)
        outs[tag] = (r.returncode, {f: (d / f).read_text() for f in sorted(os.listdir(d)) if f.endswith(".info") or f.endswith(".out")})
    assert outs["ref"][0] == 0, "reference failed on a well-formed file"
    assert outs["ours"][0] == 0
    assert sorted(outs["ours"][1]) == sorted(outs["ref"][1])
    for f in outs["ref"][1]:
        _same_text(outs["ours"][1][f], outs["ref"][1][f], (seed, args, f))
