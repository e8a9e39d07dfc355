#!/usr/bin/env python
"""Regenerates tests/golden/analysis/* : what the UNMODIFIED reference post-processor (`analysis_binary.out`, built
into oracle/_ref by `make -C oracle ref`) writes for committed output.binary files.

  analysis/<case>/in.lmp2sqt, output.binary     inputs (from the golden cases, or made here by the reference lmp2sqt)
  analysis/<case>/<mode>.json                   {file name: file text} of every *.info / *.out the reference wrote,
                                                mode = t (default), t_d2u05 (-d 2 -u 0.5), f (-f), fw (-f -w)

Fourier modes need 2 (nCTime - 1) to be a power of two (the reference corrupts its heap otherwise), hence the extra
`t9` case with nCTime 9; `two_sets` concatenates two prints so that the mean over sets is exercised.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402
from lmp2sqt_b200 import synth  # noqa: E402

MODES = {"t": [], "t_d2u05": ["-d", "2", "-u", "0.5"], "f": ["-f"], "fw": ["-f", "-w"]}


def run_ref(indir, args):
    with tempfile.TemporaryDirectory() as td:
        shutil.copy(os.path.join(indir, "in.lmp2sqt"), td)
        shutil.copy(os.path.join(indir, "output.binary"), td)
        r = subprocess.run([pyoracle.REF_ANALYSIS_BIN, *args, "output.binary"], cwd=td, capture_output=True)
        assert r.returncode == 0, (indir, args, r.returncode)
        return {f: open(os.path.join(td, f)).read() for f in sorted(os.listdir(td)) if f.endswith(".info") or f.endswith(".out")}


def main():
    pyoracle.build()
    assert pyoracle.have_reference(), "needs /root/reference"
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "analysis")
    shutil.rmtree(out, ignore_errors=True)
    cases = {}
    # 1. a committed golden case (nCTime 10: time domain only)
    d = os.path.join(out, "g2")
    os.makedirs(d)
    for f in ("in.lmp2sqt", "output.binary"):
        shutil.copy(os.path.join(here, "g2_final_print", f), d)
    cases["g2"] = ["t", "t_d2u05"]
    # 2. nCTime 9, made by the reference lmp2sqt itself
    d = os.path.join(out, "t9")
    os.makedirs(d)
    with tempfile.TemporaryDirectory() as td:
        t = synth.make_trajectory(n_cells=3, n_frames=60, seed=5)
        synth.write_text_dump(t, os.path.join(td, "dump.lammpstrj"))
        synth.write_deck(os.path.join(td, "in.lmp2sqt"), limitCorrAv=0, nCBuffer=3, nCSpatial=6, nCTime=9)
        r = pyoracle.run_reference_cli(td, "in.lmp2sqt", ["dump.lammpstrj"])
        assert r.returncode == 0
        shutil.copy(os.path.join(td, "in.lmp2sqt"), d)
        shutil.copy(os.path.join(td, "output.binary"), d)
    cases["t9"] = ["t", "f", "fw"]
    # 3. two sets in one file (a second run appends: lmp2sqt.c opens output.binary for append)
    d = os.path.join(out, "two_sets")
    os.makedirs(d)
    with tempfile.TemporaryDirectory() as td:
        synth.write_deck(os.path.join(td, "in.lmp2sqt"), limitCorrAv=0, nCBuffer=3, nCSpatial=5, nCTime=9)
        blobs = []
        for seed in (21, 22):
            t = synth.make_trajectory(n_cells=2, n_frames=50, seed=seed)
            synth.write_text_dump(t, os.path.join(td, "dump.lammpstrj"))
            if os.path.exists(os.path.join(td, "output.binary")):
                os.remove(os.path.join(td, "output.binary"))
            synth.write_deck(os.path.join(td, "in.lmp2sqt"), limitCorrAv=0, nCBuffer=3, nCSpatial=5, nCTime=9)
            r = pyoracle.run_reference_cli(td, "in.lmp2sqt", ["dump.lammpstrj"])
            assert r.returncode == 0
            blobs.append(open(os.path.join(td, "output.binary"), "rb").read())
        shutil.copy(os.path.join(td, "in.lmp2sqt"), d)
        open(os.path.join(d, "output.binary"), "wb").write(blobs[0] + blobs[1])
    cases["two_sets"] = ["t", "f"]
    for case, modes in cases.items():
        for m in modes:
            files = run_ref(os.path.join(out, case), MODES[m])
            with open(os.path.join(out, case, m + ".json"), "w") as f:
                json.dump(files, f, indent=0, sort_keys=True)
            print(case, m, len(files), "files")
    main_text(here)


# ---- the TEXT post-processor (analysis.c = AverageFqt2FqOrFt.c, built into oracle/_ref as analysis_ref.out) on output.txt ----
TEXT_MODES = {"t": [], "t_d2": ["-d", "2"], "f": ["-f"], "fw": ["-f", "-w"]}


def run_ref_text(indir, args):
    with tempfile.TemporaryDirectory() as td:
        shutil.copy(os.path.join(indir, "in.lmp2sqt"), td)
        shutil.copy(os.path.join(indir, "output.txt"), td)
        r = subprocess.run([pyoracle.REF_ANALYSIS_TEXT_BIN, *args, "output.txt"], cwd=td, capture_output=True, text=True)
        assert r.returncode == 0, (indir, args, r.returncode)
        d = {f: open(os.path.join(td, f)).read() for f in sorted(os.listdir(td)) if f.endswith(".info") or f.endswith(".out")}
        d["<stdout>"] = r.stdout
        return d


def main_text(here):
    """analysis_text/<case>/in.lmp2sqt, output.txt (written by the reference lmp2sqt) and <mode>.json = every file the
    reference text post-processor wrote, plus its stdout."""
    out = os.path.join(here, "analysis_text")
    shutil.rmtree(out, ignore_errors=True)
    cases = {}
    for case, src, modes in (("g5", "g5_ring30", ["t", "t_d2"]), ("g2", "g2_final_print", ["t"]), ("g4_f_only", "g4_collective_only", ["t"])):
        d = os.path.join(out, case)
        os.makedirs(d)
        shutil.copy(os.path.join(here, src, "in.lmp2sqt.after"), os.path.join(d, "in.lmp2sqt"))
        shutil.copy(os.path.join(here, src, "output.txt"), d)
        cases[case] = modes
    # nCTime 9 (2 (nCTime - 1) a power of two: the frequency domain), two sets concatenated
    d = os.path.join(out, "t9_two_sets")
    os.makedirs(d)
    with tempfile.TemporaryDirectory() as td:
        texts = []
        for seed in (31, 32):
            t = synth.make_trajectory(n_cells=2, n_frames=50, seed=seed)
            synth.write_text_dump(t, os.path.join(td, "dump.lammpstrj"))
            synth.write_deck(os.path.join(td, "in.lmp2sqt"), limitCorrAv=0, nCBuffer=3, nCSpatial=5, nCTime=9)
            r = pyoracle.run_reference_cli(td, "in.lmp2sqt", ["dump.lammpstrj"])
            assert r.returncode == 0
            texts.append(open(os.path.join(td, "output.txt")).read())
        shutil.copy(os.path.join(td, "in.lmp2sqt"), d)
        open(os.path.join(d, "output.txt"), "w").write(texts[0] + texts[1])
    cases["t9_two_sets"] = ["t", "f", "fw"]
    for case, modes in cases.items():
        for m in modes:
            files = run_ref_text(os.path.join(out, case), TEXT_MODES[m])
            with open(os.path.join(out, case, m + ".json"), "w") as f:
                json.dump(files, f, indent=0, sort_keys=True)
            print("text", case, m, len(files), "files")


if __name__ == "__main__":
    if sys.argv[1:] == ["text"]:
        pyoracle.build()
        main_text(os.path.dirname(os.path.abspath(__file__)))
    else:
        main()
