"""Latency-bound end of the spectrum: BASELINE configs[0] (bundled 1000 x 10000 PLINK file, k=9, Normal) and a
logistic fit on the same genotypes, wall time per fit_iht call through the public API.  Run twice
(IHTB_NO_BATCH=1 and unset) to see what the batched backtracking buys when the sweep is a few microseconds."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import mendeliht_jl_b200 as m

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
y = np.loadtxt(os.path.join(G, "normal_y.txt"))
x = m.B200SnpLinAlg.from_bed_file(os.path.join(G, "normal.bed"), 1000)
yb = (y > np.median(y)).astype(float)
rows = {}
for label, kw in (("normal_k9", dict(k=9)), ("logistic_k9", dict(k=9, d="Bernoulli", l="LogitLink"))):
    yy = y if label.startswith("normal") else yb
    for _ in range(5):
        res = m.fit_iht(yy, x, None, **kw)
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        res = m.fit_iht(yy, x, None, **kw)
    dt = (time.perf_counter() - t0) / reps
    rows[label] = {"ms_per_fit": dt * 1e3, "iterations": res.iter, "backtracks": res.n_backtracks,
                   "us_per_iteration": dt * 1e6 / res.iter}
print(json.dumps({"batched": os.environ.get("IHTB_NO_BATCH") != "1", "fits": rows}))
