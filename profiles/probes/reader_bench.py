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
    # frame-parallel read-ahead (sqt_prefetch_*): whole small frames parsed ahead by several threads, taken in file order
    for thr in (1, 2, 4, 8):
        best = 1e9
        for rep in range(3):
            r = h.sqt_reader_open(p.encode(), 1); hd = SqtFrameHeader()
            while h.sqt_reader_next_header(r, C.byref(hd)) and hd.n_atoms > 0 and h.sqt_reader_skip_atoms(r, C.byref(hd)): pass
            h.sqt_reader_rewind(r)
            t0 = time.perf_counter()
            pf = h.sqt_prefetch_start(r, 0, 0, 2 * thr + 2, thr, 10**9); soa = C.c_void_p(); n = 0
            while h.sqt_prefetch_next(pf, C.byref(hd), C.byref(soa)): n += hd.n_atoms
            best = min(best, time.perf_counter() - t0); h.sqt_prefetch_stop(pf); h.sqt_reader_close(r)
        print("read-ahead, %d frame threads: %.1f M atom lines/s, %.2f GB/s, %.2f ms per 32 000-atom frame" % (thr, n / best / 1e6, sz / best / 1e9, best / 60 * 1e3))
