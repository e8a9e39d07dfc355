#!/usr/bin/env python
"""Measurement script (like `bench.py --impl reference`: it times the reference as the CPU baseline).  Config-1 drop-in
comparison on a real text dump: the unmodified reference binary (oracle/_ref/lmp2sqt_ref.out,
all host threads) and our CLI (lmp2sqt_b200/bin/lmp2sqt_b200.out, one B200) run `-F -t -i in.lmp2sqt.150 dump` on
the same synthetic 4 000-atom LJ-liquid dump; wall clocks and the max deviation of output.binary are printed as one
JSON line.  usage: python profiles/cli_e2e_c1.py [n_frames=320]"""
import json, os, shutil, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from outbin import parse_output_binary
from lmp2sqt_b200 import _lib, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 320
S, T, B = 30, 150, 150
with tempfile.TemporaryDirectory() as td:
    traj = synth.make_trajectory(n_cells=10, n_frames=frames, seed=20261019)
    dump = os.path.join(td, "dump.lammpstrj")
    synth.write_text_dump(traj, dump)
    res = {"atoms": traj.n_atoms, "frames": frames, "nCSpatial": S, "ring": "%d/%d" % (T, B), "dump_MB": os.path.getsize(dump) / 1e6}
    out = {}
    for name, exe, env in (("ours", _lib.CLI_BIN, {}), ("reference", os.path.join(ROOT, "oracle", "_ref", "lmp2sqt_ref.out"), {"OMP_NUM_THREADS": str(len(os.sched_getaffinity(0)))})):
        d = os.path.join(td, name); os.makedirs(d)
        synth.write_deck(os.path.join(d, "in.lmp2sqt.150"), limitCorrAv=14825, nCBuffer=B, nCSpatial=S, nCTime=T)
        t0 = time.time()
        r = subprocess.run([exe, "-F", "-t", "-i", "in.lmp2sqt.150", dump], cwd=d, env=dict(os.environ, **env), capture_output=True, text=True)
        res[name + "_wall_s"] = time.time() - t0
        assert r.returncode == 0, r.stderr[-500:]
        out[name] = parse_output_binary(os.path.join(d, "output.binary"), S, T)
    res["host_threads_reference"] = len(os.sched_getaffinity(0))
    res["speedup_wall"] = res["reference_wall_s"] / res["ours_wall_s"]
    for key in ("full-density", "self-density", "cross-density", "self-vanHove"):
        a, b = out["ours"][key], out["reference"][key]
        tol = np.maximum(np.abs(b), 1e-3 * np.abs(b[:1, :])) + 1e-300
        res["max_rel_err_" + key] = float(np.max(np.abs(a - b) / tol))
    res["units_per_s_ours"] = traj.n_atoms * 3 * S * frames / res["ours_wall_s"]
    res["units_per_s_reference"] = traj.n_atoms * 3 * S * frames / res["reference_wall_s"]
    print(json.dumps(res))
