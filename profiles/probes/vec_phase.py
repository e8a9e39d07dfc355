"""Per-warp phase clocks of self_ring_vec_kernel (diagnostic build: make -C lmp2sqt_b200/csrc timing).
usage: python profiles/probes/vec_phase.py [n_cells=10] [nCSpatial=30]"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lmp2sqt_b200 import _lib
_lib.ENGINE_SO = os.path.join(ROOT, "lmp2sqt_b200", "libsqt_b200_timing.so")
from lmp2sqt_b200 import Deck, SpacetimeCorr, SynthSource
lib = _lib.engine_lib()
n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 10
S, T, B = (int(sys.argv[2]) if len(sys.argv) > 2 else 30), 150, 150
n = 4 * n_cells ** 3
eng = SpacetimeCorr(Deck(limitCorrAv=10**9, nCBuffer=B, nCSpatial=S, nCTime=T)); src = SynthSource(n_cells, 0, n)
eng.InitSpacetimeCorr()
for _ in range(T + 3): eng.push_device(src.next(), n, n, 0.0, src.L)
eng.sync()
tim = (C.c_ulonglong * 16)()
lib.sqt_debug_vec_tim(tim)
K = 20
frames = [src.next() for _ in range(1)]
eng.timer_start()
for _ in range(K): eng.push_device(frames[0], n, n, 0.0, src.L)
ms = eng.timer_stop()
lib.sqt_debug_vec_tim(tim)
w, lw, steps = tim[6] / K, tim[7] / K, tim[8] / K
print(eng.kernel_config.split(";")[0])
print("frame %.2f us; warps %d, steps per warp %.1f; cycles per warp: prologue %.0f, loop %.0f (%.0f per step), butterfly+store %.0f, fence+counter %.0f, "
      "histogram flush + exit %.0f; group sum %.0f cycles on each of %d last warps; longest warp %d cycles = %.1f us"
      % (ms / K * 1e3, w, steps / w, tim[0] / K / w, tim[1] / K / w, tim[1] / max(1, tim[8]), tim[2] / K / w, tim[3] / K / w, tim[5] / K / w,
         tim[4] / K / max(1, lw), lw, tim[9], tim[9] / 1.965e3))
