"""Reader of `output.binary` (layout written by lmp2sqt's PrintSpacetimeCorr_binary, lmp2sqt.c:768-864) for the
measurement scripts in this directory: {block name: array [nCTime][nCSpatial] or None}."""
import numpy as np


def parse_output_binary(path, nCSpatial, nCTime):
    data = open(path, "rb").read()
    off = 0

    def take(fmt, n=1):
        nonlocal off
        a = np.frombuffer(data, dtype=fmt, count=n, offset=off)
        off += a.nbytes
        return a

    n = int(take("<i4")[0])
    off += n                                   # "##space-time corr"
    out = {}
    for _ in range(int(take("<i4")[0])):
        ln = int(take("<i4")[0])
        name = data[off:off + ln].decode()
        off += ln
        take("<f8", 3)                         # kVal, deltaT, rVal
        out[name] = take("<f8", nCSpatial * nCTime).reshape(nCTime, nCSpatial).copy() if int(take("<i4")[0]) else None
    return out
