"""Post-processing of the correlators (reference: analysis_binary.c): F(q,t) tables, dF/dt, Gamma(q,t) = d ln F/dt,
D(q,t) = -Gamma/q^2, D_c(q,t) = -ln(F/F0)/(q^2 t), H = D F0, optional cosine transform to omega, written in the
reference's file formats.  The arithmetic lives in csrc/host/sqt_analysis.c (include/sqt_host.h); this is the
ctypes mirror, usable on an ``output.binary`` file or directly on :class:`Results` fetched from the engine."""
from __future__ import annotations

import ctypes as C
import math
import os

import numpy as np

from . import _lib

BLOCKS = ("full-density", "self-density", "cross-density", "self-vanHove")
TABLES = ("mean", "sd", "err", "F", "F1", "Gamma", "D", "Dc", "H", "Hc")


class Analysis:
    def __init__(self):
        self.lib = _lib.host_lib()
        self.c = _lib.SqtAnalysis()
        self._live = False

    @classmethod
    def load(cls, deck_path: str, binary_path: str, n_set_skip: int = -1) -> "Analysis":
        a = cls()
        rc = a.lib.sqt_analysis_load(C.byref(a.c), os.fsencode(deck_path), os.fsencode(binary_path), n_set_skip)
        if rc:
            raise OSError("sqt_analysis_load(%s, %s) failed: %d" % (deck_path, binary_path, rc))
        a._live = True
        return a

    @classmethod
    def from_results(cls, results, deck, flags: str = "Ft") -> "Analysis":
        """``results``: a :class:`lmp2sqt_b200.engine.Results` that has been normalised (``results.normalise(deck, flags)``)."""
        a = cls()
        dk = deck.to_c()
        rc = a.lib.sqt_analysis_from_results(C.byref(a.c), C.byref(results.c), C.byref(dk), int("F" in flags))
        if rc:
            raise MemoryError("sqt_analysis_from_results failed: %d" % rc)
        a._live = True
        return a

    def time(self, n_val_diff: int = 4, under_limit: float = math.exp(-2.5), outdir: str | None = None) -> "Analysis":
        rc = self.lib.sqt_analysis_time(C.byref(self.c), n_val_diff, under_limit, os.fsencode(outdir) if outdir else None)
        if rc:
            raise RuntimeError("sqt_analysis_time failed: %d" % rc)
        return self

    def fourier(self, window: bool = False) -> "Analysis":
        rc = self.lib.sqt_analysis_fourier(C.byref(self.c), int(window))
        if rc:
            raise ValueError("sqt_analysis_fourier: 2 (nCTime - 1) must be a power of two (rc %d)" % rc)
        return self

    def print(self, outdir: str | None = None) -> None:
        if self.lib.sqt_analysis_print(C.byref(self.c), os.fsencode(outdir) if outdir else None):
            raise OSError("sqt_analysis_print failed")

    def table(self, name: str, block: int | str) -> np.ndarray:
        """Copy of one table as [nCSpatial][nCTime]."""
        j = BLOCKS.index(block) if isinstance(block, str) else block
        p = getattr(self.c, name)[j]
        n = self.c.nCSpatial * self.c.nCTime
        return np.ctypeslib.as_array(p, shape=(n,)).reshape(self.c.nCSpatial, self.c.nCTime).copy()

    def close(self):
        if self._live:
            self.lib.sqt_analysis_free(C.byref(self.c))
            self._live = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
