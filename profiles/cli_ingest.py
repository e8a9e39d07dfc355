#!/usr/bin/env python
"""Ingest rate of the command-line drop-in: the same 256 000-atom trajectory as LAMMPS text and as a raw binary
frame file (SQTRAW01), short ring so that the GPU work is negligible and the reader sets the wall clock.
usage: python profiles/cli_ingest.py [n_frames=40]"""
import json, os, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from outbin import parse_output_binary
from lmp2sqt_b200 import _lib, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 40
S, T, B = 30, 10, 10
with tempfile.TemporaryDirectory() as td:
    traj = synth.make_trajectory(n_cells=40, n_frames=frames, seed=3)
    res = {"atoms": traj.n_atoms, "frames": frames}
    out = {}
    for kind in ("text", "raw"):
        d = os.path.join(td, kind); os.makedirs(d)
        dump = os.path.join(d, "dump.lammpstrj" if kind == "text" else "dump.sqtraw")
        (synth.write_text_dump if kind == "text" else synth.write_raw_dump)(traj, dump)
        synth.write_deck(os.path.join(d, "in.lmp2sqt"), limitCorrAv=14825, nCBuffer=B, nCSpatial=S, nCTime=T)
        subprocess.run(["cat", dump], stdout=subprocess.DEVNULL)          # page cache warm for both
        t0 = time.time()
        r = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-i", "in.lmp2sqt", os.path.basename(dump)], cwd=d, capture_output=True, text=True)
        w = time.time() - t0
        assert r.returncode == 0, r.stderr[-500:]
        res[kind] = {"file_MB": os.path.getsize(dump) / 1e6, "wall_s": w, "atoms_per_s_two_passes": 2 * traj.n_atoms * frames / w}
        out[kind] = parse_output_binary(os.path.join(d, "output.binary"), S, T)
    res["outputs_equal"] = all(np.allclose(out["text"][k], out["raw"][k], rtol=1e-11, atol=1e-12) for k in ("full-density", "self-density", "self-vanHove"))
    print(json.dumps(res))
