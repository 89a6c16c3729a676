import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mendeliht_jl_b200 as m
from mendeliht_jl_b200 import synth
from oracle import glm, iht, snp
d, l, n, p, k, ncov, miss = "Bernoulli", "LogitLink", 1500, 3000, 6, 0, 0.0
seed = 100 + n + p
y, z, _, _, _ = synth.simulate_response(seed, n, p, k, d, n_cov=ncov, missing_rate=miss)
bed = synth.packed_columns(seed, n, np.arange(p), miss)
g = m.B200SnpLinAlg.from_bed_columns(bed, n)
o = snp.SnpLinAlgOracle(bed, n)
for mode in (0, 1):
    res = m.fit_iht(y, g, z, k=k + 2, d=d, l=l, nb_r=10.0, sweep_mode=mode)
    ref = iht.fit_iht(y, o, z, k=k + 2, d=d, l=l, nb_r=10.0)
    print("mode", mode, res.iter, ref.iter)
    for i, t in enumerate(res.trace):
        rel = abs(t[0] - ref.trace.logl[i]) / abs(ref.trace.logl[i])
        if rel > 1e-9 or i < 3:
            print(i + 1, t[0], ref.trace.logl[i], rel, "bt", t[1], ref.trace.backtracks[i], "eta", t[3], ref.trace.eta[i], "tol", t[2], ref.trace.tol[i], "ncand", t[4])
