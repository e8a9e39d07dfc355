"""Time the ring kernel of diagnostic builds (libsqt_b200_diag_*.so: pieces of the seed phase compiled out, results
are wrong on purpose) to see how much each piece costs.  usage: python profiles/probes/diag_variants.py <so> [n_cells]"""
import os, sys
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
for _ in range(T + 3): eng.push_device(src.next(), n, n, 0.0, src.L)
eng.sync(); eng.profile(True)
for _ in range(5): eng.push_device(src.next(), n, n, 0.0, src.L)
k, ms = eng.profile_read()
print(sys.argv[1], "ring kernel %.3f ms" % (ms / k), "frac %.4f" % (n * 513 * 456 / (ms / k * 1e-3) / 36.05e12))
