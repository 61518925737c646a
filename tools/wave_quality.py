"""Dev diagnostic: recall of the GPU-built graph vs wave cap (build_batch) against the oracle-built graph, GMM data."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import embeddinghub_b200 as ehb
from oracle import oracle as orc
bench.DIST = "gmm"
N, d, Q, k = int(sys.argv[1]), 128, 1000, 10
base, q = bench.gen(N, d, bench.BASE_SEED), bench.gen(Q, d, bench.QUERY_SEED)
rec = lambda a, b: float(np.mean([len(set(x.tolist()) & set(y.tolist())) / k for x, y in zip(a, b)]))
gt = None
for bb in (16384, 4096, 1024):
    ix = ehb.NativeIndex(d, capacity=N, build_batch=bb)
    ix.add(base); t = time.time(); ix.build(); tb = time.time() - t
    if gt is None: gt, _, _ = ix.search_bruteforce(q, k)
    ix.set_search_width(1)
    print(f"wave cap {bb}: build {tb:.2f}s recall ef16/64/128 =", [round(rec(ix.search(q, k, ef=e)[0], gt), 4) for e in (16, 64, 128)], flush=True)
    del ix
o = orc.OracleHNSW(d, "l2", N); t = time.time(); o.add(base, threads=64); print(f"oracle build {time.time()-t:.1f}s", flush=True)
print("oracle graph recall ef16/64/128 =", [round(rec(o.search(q, k, ef=e, threads=32)[0], gt), 4) for e in (16, 64, 128)])
