#!/usr/bin/env python
"""Ingest-included throughput of the command-line drop-in at config-3 size ("ingest-included and
kernel-only figures separately"): 1 000 188 atoms, nCSpatial 86, ring 150/150, frames from a raw binary frame file
(SQTRAW01, the high-rate frame source of section 8f.1) written by the on-device source, through
`lmp2sqt_b200.out -F -t [-g N]` -- one reader thread, one pinned frame, N chunk copies.
usage: python profiles/cli_ingest_c3.py [n_frames=240] [gpus=1,8] [nCSpatial=86 nCTime=150 nCBuffer=150 [n_cells=63]]   -> one JSON line
(config 4: `170 1,8 128 150 150 100` = 4 000 000 atoms, a 16 GB file)
(a small deck, e.g. `240 1 10 10 10`, makes the GPU work negligible: what is left is the reader's own pace)"""
import json, os, struct, subprocess, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from outbin import parse_output_binary
from lmp2sqt_b200 import SynthSource, _lib, synth

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 240
gpus = [int(g) for g in (sys.argv[2] if len(sys.argv) > 2 else "1").split(",")]
n_cells, S, T, B = 63, 86, 150, 150
if len(sys.argv) > 5:
    S, T, B = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
if len(sys.argv) > 6:
    n_cells = int(sys.argv[6])
n = 4 * n_cells ** 3
base = "/dev/shm" if os.path.isdir("/dev/shm") and os.statvfs("/dev/shm").f_bavail * os.statvfs("/dev/shm").f_frsize > 2 * frames * n * 24 else None
with tempfile.TemporaryDirectory(dir=base) as td:
    dump = os.path.join(td, "dump.sqtraw")
    src = SynthSource(n_cells, 0, n, device=0)
    L = src.L
    t0 = time.time()
    buf = np.empty((3, n))
    with open(dump, "wb") as f:
        f.write(b"SQTRAW01")
        for k in range(frames):
            src.next()
            src.to_host(buf)
            f.write(struct.pack("<qq6d", k * 100, n, 0.0, L, 0.0, L, 0.0, L))
            f.write(buf.tobytes())
    src.close()
    res = {"atoms": n, "frames": frames, "nCSpatial": S, "ring": "%d/%d" % (T, B), "file_GB": os.path.getsize(dump) / 1e9,
           "reader_threads": os.environ.get("SQT_READER_THREADS", "auto"),
           "file_written_s": time.time() - t0, "tmp": td.split("/")[1], "runs": {}}
    ref = None
    for g in gpus:
        d = os.path.join(td, "g%d" % g); os.makedirs(d)
        synth.write_deck(os.path.join(d, "in.lmp2sqt"), limitCorrAv=14825, nCBuffer=B, nCSpatial=S, nCTime=T)
        t0 = time.time()
        r = subprocess.run([_lib.CLI_BIN, "-F", "-t", "-g", str(g), "-i", "in.lmp2sqt", dump], cwd=d, capture_output=True, text=True)
        w = time.time() - t0
        assert r.returncode == 0, r.stderr[-800:]
        out = parse_output_binary(os.path.join(d, "output.binary"), S, T)
        if ref is None:
            ref = out
        same = all(np.allclose(out[k], ref[k], rtol=1e-9, atol=1e-12) for k in ("full-density", "self-density", "self-vanHove"))
        last = r.stderr.strip().split("\n")[-1]
        loop_s = float(last.split("frame loop ")[1].split(" s")[0])      # the CLI's own clock around the frame loop (second pass)
        res["runs"]["g%d" % g] = {"wall_s": w, "startup_s": w - loop_s, "frame_loop_s": loop_s, "frames_per_s": frames / loop_s,
                                  "ms_per_frame": 1e3 * loop_s / frames, "particle_k_frame_per_s": n * 3 * S * frames / loop_s,
                                  "same_output_as_first_run": bool(same), "engine": last[:120]}
    print(json.dumps(res))
