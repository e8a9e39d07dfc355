"""Per-warp phase clocks of self_ring_kernel (diagnostic build: make -C lmp2sqt_b200/csrc timing).
usage: python profiles/probes/phase_timing.py [n_cells]"""
import collections, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lmp2sqt_b200 import _lib
_lib.ENGINE_SO = os.path.join(ROOT, "lmp2sqt_b200", "libsqt_b200_timing.so")
from lmp2sqt_b200 import Deck, SpacetimeCorr, SynthSource
lib = _lib.engine_lib()
n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 40
S, T, B = (int(sys.argv[2]) if len(sys.argv) > 2 else 171), 150, 150
n = 4 * n_cells ** 3
eng = SpacetimeCorr(Deck(limitCorrAv=10**9, nCBuffer=B, nCSpatial=S, nCTime=T)); src = SynthSource(n_cells, 0, n)
eng.InitSpacetimeCorr()
for _ in range(T + 3): eng.push_device(src.next(), n, n, 0.0, src.L)
eng.sync()
tim = (C.c_ulonglong * 8)(); wb = (C.c_ulonglong * (16384 * 2))()
lib.sqt_debug_tim(tim, None)
eng.timer_start(); eng.push_device(src.next(), n, n, 0.0, src.L); ms = eng.timer_stop()
lib.sqt_debug_tim(tim, wb)
hot, seed, allc, ns = tim[0], tim[1], tim[2], tim[3]
print(eng.kernel_config)
print("frame %.3f ms; kernel span %.3f ms; per warp-super-iteration cycles: total %.0f hot %.0f seed %.0f other %.0f" % (ms, (tim[5] - tim[4]) / 1e6, allc / ns, hot / ns, seed / ns, (allc - hot - seed) / ns))
a = np.frombuffer(wb, dtype=np.uint64).reshape(16384, 2)
a = a[a[:, 0] > 0]
dur = a[:, 0].astype(float) / 1e6; smid = (a[:, 1] >> np.uint64(32)).astype(int); nsup = (a[:, 1] & np.uint64(0xffffffff)).astype(int)
print("warps %d: busy ms min/mean/max %.3f %.3f %.3f; super-iterations per warp min/mean/max %d %.0f %d" % (len(dur), dur.min(), dur.mean(), dur.max(), nsup.min(), nsup.mean(), nsup.max()))
per_sm = collections.defaultdict(int)
for s_, k in zip(smid, nsup): per_sm[s_] += k
v = np.array(list(per_sm.values()))
print("super-iterations per SM: min %d mean %.0f max %d over %d SMs" % (v.min(), v.mean(), v.max(), len(v)))
