"""The command-line driver's bookkeeping before the first frame is processed -- which input files are accepted
(lmp2sqt.c:221-247), limitCorrAv = deck + sum over files, the rewritten deck (UpdateNameList, :257) -- against the
unmodified reference binary on random decks and file sets.  Runs without a GPU: our driver does all of this before it
creates the engine (where it then stops with the engine's error), the reference runs the whole (tiny) job."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import pyoracle
from lmp2sqt_b200 import _lib, synth

pytestmark = pytest.mark.skipif(not pyoracle.have_reference(), reason="oracle/_ref not built")


def _case(rng, d):
    T = int(rng.integers(2, 9))
    B = int(rng.integers(1, T + 3))                       # may exceed nCTime: clipped by GetNameList (lmp2sqt.c:2105)
    skip = int(rng.integers(0, 3)) if rng.random() < 0.4 else None
    limit = int(rng.choice([0, 3, 14825]))
    synth.write_deck(os.path.join(d, "in.lmp2sqt"), limitCorrAv=limit, nCBuffer=B, nCSpatial=int(rng.integers(2, 6)), nCTime=T, nCSkip=skip)
    names = []
    for i in range(int(rng.integers(1, 4))):
        frames = int(rng.choice([2, 4, 5, 6, T, T + 1, T + 2, 2 * T + 3, 25]))
        t = synth.make_trajectory(n_cells=1, n_frames=frames, seed=int(rng.integers(0, 10**6)))
        name = "dump%d.lammpstrj" % i
        synth.write_text_dump(t, os.path.join(d, name))
        names.append(name)
    if rng.random() < 0.2:
        names.append("missing.lammpstrj")
    return names


@pytest.mark.parametrize("seed", range(24))
def test_file_acceptance_and_deck_rewrite_match_reference(tmp_path, seed):
    rng = np.random.default_rng(1000 + seed)
    ref_dir, our_dir = tmp_path / "ref", tmp_path / "ours"
    ref_dir.mkdir()
    names = _case(rng, str(ref_dir))
    shutil.copytree(ref_dir, our_dir)
    r = pyoracle.run_reference_cli(str(ref_dir), "in.lmp2sqt", names)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")       # bookkeeping only: never touch a GPU even where one exists
    o = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-i", "in.lmp2sqt", *names], cwd=our_dir, capture_output=True, text=True, env=env)
    ref_deck, our_deck = (ref_dir / "in.lmp2sqt").read_bytes(), (our_dir / "in.lmp2sqt").read_bytes()
    assert our_deck == ref_deck, (names, r.stderr[-600:], o.stderr[-600:])
    # the same files are turned away with the same words
    def rejected(err):
        return sorted(ln for ln in err.split("\n") if "this file(" in ln or "Can`t open" in ln)
    assert rejected(o.stderr) == rejected(r.stderr)
    usable = "Check your input options" not in r.stdout
    assert ("Check your input options" not in o.stdout) == usable
    if usable:
        assert "sqt_multi_create failed" in o.stderr             # got as far as the engine, which needs a GPU
