#!/usr/bin/env python
"""The command-line drop-in on a config-2-shaped TEXT dump (32 000 atoms per frame, in.lmp2sqt.150.smallQ): frame loop with the
frame-parallel read-ahead (default) and without (SQT_PREFETCH=0).  usage: python profiles/cli_text_c2.py [n_frames=200]"""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from outbin import parse_output_binary
from lmp2sqt_b200 import _lib, synth
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 200
with tempfile.TemporaryDirectory() as td:
    t = synth.make_trajectory(n_cells=20, n_frames=frames, seed=2)
    dump = os.path.join(td, "d.lammpstrj")
    synth.write_text_dump(t, dump)
    res = {"atoms": t.n_atoms, "frames": frames, "file_MB": os.path.getsize(dump) / 1e6}
    outs = {}
    for name, env in (("sequential", {"SQT_PREFETCH": "0"}), ("read_ahead", {})):
        d = os.path.join(td, name); os.makedirs(d)
        synth.write_deck(os.path.join(d, "in.lmp2sqt"), limitCorrAv=14825, nCBuffer=150, nCSpatial=10, nCTime=150)
        r = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-i", "in.lmp2sqt", dump], cwd=d, capture_output=True, text=True, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr[-500:]
        loop_s = float(r.stderr.strip().split("\n")[-1].split("frame loop ")[1].split(" s")[0])
        res[name] = {"frame_loop_s": loop_s, "ms_per_frame": 1e3 * loop_s / frames}
        outs[name] = parse_output_binary(os.path.join(d, "output.binary"), 10, 150)
    res["same_output"] = all(np.allclose(outs["sequential"][k], outs["read_ahead"][k], rtol=1e-9, atol=1e-12) for k in ("full-density", "self-density", "self-vanHove"))
    print(json.dumps(res))
