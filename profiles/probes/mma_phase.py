"""Per-phase warp clocks of the FP64 tensor-core self kernel (diagnostic build with -DSQT_MMA_PHASE).
usage: python profiles/probes/mma_phase.py libsqt_b200_diag_PHASE.so [n_cells]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lmp2sqt_b200 import _lib
_lib.ENGINE_SO = os.path.join(ROOT, "lmp2sqt_b200", sys.argv[1])
from lmp2sqt_b200 import Deck, SpacetimeCorr, SynthSource
n_cells = int(sys.argv[2]) if len(sys.argv) > 2 else 63
S, T, B = 171, 150, 150
n = 4 * n_cells ** 3
eng = SpacetimeCorr(Deck(limitCorrAv=10**9, nCBuffer=B, nCSpatial=S, nCTime=T)); src = SynthSource(n_cells, 0, n)
eng.InitSpacetimeCorr()
raw = ctypes.CDLL(_lib.ENGINE_SO)
out = (ctypes.c_ulonglong * 8)()
for _ in range(T + 3): eng.push_device(src.next(), n, n, 0.0, src.L)
eng.sync(); raw.sqt_mma_phase_read(out)
eng.profile(True)
F = 5
for _ in range(F): eng.push_device(src.next(), n, n, 0.0, src.L)
k, ms = eng.profile_read()
raw.sqt_mma_phase_read(out)
w = out[5] / F
units = (n / 32) * 150 * 3 / w  # (slot, axis, 32-particle) units per warp per frame
names = ["slot loop head + load wait", "displacement + van Hove", "cos + Chebyshev rows -> smem", "fragment loads + DMMA"]
print("ring %.3f ms; %d warps; %.0f units/warp; warp lifetime %.0f clk" % (ms / k, w, units, out[4] / out[5]))
for i in range(4):
    per = out[i] / out[5] / (units if i >= 2 else units / 3)
    print("  %-32s %8.1f clk per %s  (%.1f%% of lifetime)" % (names[i], per, "unit" if i >= 2 else "slot", 100.0 * out[i] / out[4]))
