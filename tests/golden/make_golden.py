#!/usr/bin/env python
"""Regenerates tests/golden/* by running the UNMODIFIED reference here (needs /root/reference
built into oracle/_ref by `make -C oracle ref`).  The outputs are committed so that the CPU and
GPU test-suites can check the oracle restatement, the host writers and the CUDA path without the
reference being present (it does not exist on the GPU box).

Per case <name>/ :
  dump*.lammpstrj.gz   input text dump(s)                      (synthetic, lmp2sqt_b200.synth)
  in.lmp2sqt           deck as given to the reference
  in.lmp2sqt.after     deck as the reference rewrote it        (lmp2sqt.c:2112-2125)
  output.binary / output.txt / SSF000.info / Dt000.info / output.log / stdout.txt
                       what `lmp2sqt_ref.out <flags> -i in.lmp2sqt dump...` wrote (OMP_NUM_THREADS=1)
  arrays.npz           raw fp64 sums + normalised arrays read from the reference's globals when the
                       same frames are pushed through its own EvalSpacetimeCorr in-process
                       (oracle/pyoracle.py:ReferenceLib), with limitCorrAv large (no mid-run flush)
  case.json            parameters
"""
import glob
import gzip
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyoracle  # noqa: E402
from lmp2sqt_b200 import synth  # noqa: E402

CASES = {
    # mid-run print: deck limitCorrAv 0 + computed 15 is reached at frame 38
    "g1_midrun_print": dict(traj=[dict(n_cells=2, n_frames=40, seed=1)], S=4, T=10, B=5, skip=None, limit=0, flags="Ft"),
    # shipped-deck style: huge limitCorrAv, end-of-run print covers all buffers
    "g2_final_print": dict(traj=[dict(n_cells=2, n_frames=40, seed=2)], S=4, T=10, B=5, skip=None, limit=14825, flags="Ft"),
    # ragged N, B == T, nCSkip 1, fluctuating box, two input files
    "g3_two_files_skip": dict(traj=[dict(n_cells=3, n_atoms=50, n_frames=37, seed=3, box_jitter=0.01),
                                    dict(n_cells=3, n_atoms=50, n_frames=29, seed=4, box_jitter=0.01)],
                              S=6, T=8, B=8, skip=1, limit=14825, flags="Ft"),
    # collective only (-f -t): self/cross blocks absent from output.binary
    "g4_collective_only": dict(traj=[dict(n_cells=2, n_frames=33, seed=5)], S=5, T=6, B=3, skip=None, limit=14825, flags="ft"),
    # in.lmp2sqt.150.smallQ-like ring (T = B = 150 is too slow for a fixture; 30/30 keeps the shape)
    "g5_ring30": dict(traj=[dict(n_cells=3, n_frames=100, seed=6)], S=10, T=30, B=30, skip=None, limit=14825, flags="Ft"),
    # nCSpatial in the range the tensor-core kernels serve (>= 64); 128 = the second column layout, config 4's nCSpatial
    "g6_nCS70": dict(traj=[dict(n_cells=3, n_atoms=100, n_frames=36, seed=7, box_jitter=0.005)], S=70, T=8, B=8, skip=None, limit=14825, flags="Ft"),
    "g7_nCS128": dict(traj=[dict(n_cells=3, n_frames=30, seed=8)], S=128, T=6, B=3, skip=None, limit=14825, flags="Ft"),
    # flag combinations without -t / without -F: the van Hove block is all zero without -t (EvalTwoTimeEach as a whole is
    # `if (flag_t)`, lmp2sqt.c:1370), no Dt file; -t alone writes only the Dt file; -f -t at a tensor-core nCSpatial
    "g8_F_only": dict(traj=[dict(n_cells=2, n_frames=36, seed=10)], S=6, T=8, B=4, skip=None, limit=14825, flags="F"),
    "g9_t_only": dict(traj=[dict(n_cells=2, n_frames=36, seed=11)], S=6, T=8, B=4, skip=None, limit=14825, flags="t"),
    "g10_ft_nCS70": dict(traj=[dict(n_cells=2, n_frames=30, seed=12)], S=70, T=6, B=6, skip=None, limit=14825, flags="ft"),
}
# SURVEY.md section 8f.4 (next row, oracle first): -V, dumps with vx vy vz.  Kept apart from CASES: the product's host
# code and engine do not implement -V yet, only oracle/sqt_oracle.c does; tests/test_oracle.py pins it on this case.
CASES_V = {
    "v1_velocity": dict(traj=[dict(n_cells=2, n_frames=40, seed=9, box_jitter=0.005)], S=5, T=8, B=4, skip=None, limit=14825, flags="FtV"),
    # -M on a dump that carries mux muy muz (the reader normalises them, snapshot.c:440-459): C2_mu_mu / C2_mu_rho / C2_rho_mu
    # files and the (always zero) Cmm column of the Dt file
    "m1_magnet": dict(traj=[dict(n_cells=2, n_frames=40, seed=13, box_jitter=0.005)], S=5, T=8, B=4, skip=None, limit=14825, flags="FtM"),
    # -s on a dump that carries vx vy vz: Stress file, msdx msdy msdz columns; with -F (sum_n v (x) r zeroed per frame) and
    # alone (-s switches -t on; without -f the reference never zeroes sumVR_ct, so it keeps growing from frame to frame)
    "s1_stress": dict(traj=[dict(n_cells=2, n_frames=40, seed=14, box_jitter=0.005)], S=5, T=8, B=4, skip=None, limit=14825, flags="Fts"),
    # (zero_heap: without it the MSDCM column of this run's Dt file is uninitialised heap, rrMSDCM is never zeroed: lmp2sqt.c:1325-1333)
    "s2_stress_alone": dict(traj=[dict(n_cells=2, n_frames=36, seed=15)], S=6, T=6, B=3, skip=1, limit=14825, flags="s", zero_heap=True),
}


def main():
    pyoracle.build()
    assert pyoracle.have_reference(), "needs /root/reference"
    here = os.path.dirname(os.path.abspath(__file__))
    only = sys.argv[1:]                                   # optional: regenerate just the named cases
    for name, c in list(CASES.items()) + list(CASES_V.items()):
        if only and name not in only:
            continue
        vel = "V" in c["flags"] or "s" in c["flags"]
        dip = "M" in c["flags"]
        out = os.path.join(here, "velocity", name) if (vel or dip) else os.path.join(here, name)
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
        trajs = [synth.make_trajectory(**kw) for kw in c["traj"]]
        with tempfile.TemporaryDirectory() as td:
            dumps = []
            for i, t in enumerate(trajs):
                dn = "dump%d.lammpstrj" % i
                synth.write_text_dump(t, os.path.join(td, dn), velocities=vel, dipoles=dip)
                dumps.append(dn)
            synth.write_deck(os.path.join(td, "in.lmp2sqt"), limitCorrAv=c["limit"], nCBuffer=c["B"],
                             nCSpatial=c["S"], nCTime=c["T"], nCSkip=c["skip"])
            shutil.copy(os.path.join(td, "in.lmp2sqt"), os.path.join(out, "in.lmp2sqt"))
            cli_flags = tuple("-" + ch for ch in c["flags"])
            r = pyoracle.run_reference_cli(td, "in.lmp2sqt", dumps, flags=cli_flags, threads=1, zero_heap=c.get("zero_heap", False))
            assert r.returncode == 0, r.stderr
            with open(os.path.join(out, "stdout.txt"), "w") as f:
                f.write(r.stdout)
            shutil.copy(os.path.join(td, "in.lmp2sqt"), os.path.join(out, "in.lmp2sqt.after"))
            for fn in ("output.binary", "output.txt", "output.log"):
                if os.path.exists(os.path.join(td, fn)):
                    shutil.copy(os.path.join(td, fn), os.path.join(out, fn))
            vel_files = tuple((n + "000.info.*", n + "000.info") for n in ("C2_v_v", "C2_v_rho", "C2_rho_v", "C2_mu_mu", "C2_mu_rho", "C2_rho_mu",
                                                                            "C2_v_mu", "C2_mu_v", "Stress"))
            for pat, dst in (("SSF000.info.*", "SSF000.info"), ("Dt000.info.*", "Dt000.info")) + vel_files:
                g = glob.glob(os.path.join(td, pat))
                if g:
                    shutil.copy(g[0], os.path.join(out, dst))
            for dn in dumps:
                with open(os.path.join(td, dn), "rb") as fi, gzip.GzipFile(os.path.join(out, dn + ".gz"), "wb", mtime=0) as fo:
                    fo.write(fi.read())
        ref = pyoracle.run_reference(trajs, c["S"], c["T"], c["B"], nCSkip=c["skip"] or 0, flags=c["flags"], limit=10**9)
        arrs = {"raw_" + k: v for k, v in ref["raw"].items()}
        arrs.update({"printed_" + k: v for k, v in ref.get("printed", {}).items()})
        arrs["count_log"] = np.asarray(ref["log"])
        np.savez_compressed(os.path.join(out, "arrays.npz"), **arrs)
        with open(os.path.join(out, "case.json"), "w") as f:
            json.dump(c, f, indent=1)
        print(name, "countCorrAv", ref["raw"]["countCorrAv"], sorted(os.listdir(out)))


if __name__ == "__main__":
    main()
