import sys, json
sys.path.insert(0, '/root/repo')
import mendeliht_jl_b200 as m
for n, p in [(100000, 500000), (500000, 125000), (50000, 500000)]:
    g = m.B200SnpLinAlg.synthetic(n, p, 2025)
    for nc in (100, 300, 1000, 3000, 10000, 30000):
        ms, err = g.gather_bench(nc, 5)
        print(json.dumps({"n": n, "p": p, "ncols": nc, "ms": round(ms, 4), "us_per_col": round(ms * 1e3 / nc, 3), "err": err}), flush=True)
    g.close()
