#!/usr/bin/env python
"""Compare the raw sums two north_star_job.py runs dumped (--dump), e.g. 1 GPU against 8 GPUs on the same frames:
python profiles/compare_dumps.py a.npz b.npz  ->  one JSON line (max deviations; F against the SURVEY 8a tolerance)."""
import json
import sys

import numpy as np

a, b = np.load(sys.argv[1]), np.load(sys.argv[2])
out = {"a": sys.argv[1], "b": sys.argv[2], "countCorrAv": [int(a["countCorrAv"]), int(b["countCorrAv"])]}
for k in a.files:
    x, y = a[k], b[k]
    if x.shape == ():
        continue
    s = float(np.abs(x).max())
    out[k] = {"max_abs_diff": float(np.abs(x - y).max()), "scale": s, "rel": float(np.abs(x - y).max() / s) if s else 0.0,
              "identical": bool(np.array_equal(x, y))}
for k in ("avF_qq2", "avF_s_qq2", "avF_d_qq2"):
    tol = 1e-6 * np.maximum(np.abs(a[k]), 1e-3 * np.abs(a[k][:, :1]))
    out[k]["worst_fraction_of_survey_tolerance"] = float((np.abs(a[k] - b[k]) / tol).max())
print(json.dumps(out))
