"""Parse rate of the mmap text-dump reader (host only): 32 000-atom frames, 1 .. 16 threads.
usage: python profiles/probes/reader_bench.py"""
import ctypes as C, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
from lmp2sqt_b200 import _lib, synth
from lmp2sqt_b200._lib import SqtFrameHeader
h = _lib.host_lib()
with tempfile.TemporaryDirectory() as td:
    p = os.path.join(td, "d.lammpstrj")
    t = synth.make_trajectory(n_cells=20, n_frames=60, seed=1)
    synth.write_text_dump(t, p)
    sz = os.path.getsize(p)
    for thr in (1, 2, 4, 8, 16):
        best = 1e9
        for rep in range(3):
            r = h.sqt_reader_open(p.encode(), thr); hd = SqtFrameHeader(); n = 0
            xyz = np.empty((3, t.n_atoms))
            t0 = time.perf_counter()
            while h.sqt_reader_next_header(r, C.byref(hd)):
                if not h.sqt_reader_read_atoms(r, C.byref(hd), xyz[0].ctypes.data, xyz[1].ctypes.data, xyz[2].ctypes.data): break
                n += hd.n_atoms
            best = min(best, time.perf_counter() - t0); h.sqt_reader_close(r)
        print("threads %2d: %.1f M atom lines/s, %.2f GB/s" % (thr, n / best / 1e6, sz / best / 1e9))
