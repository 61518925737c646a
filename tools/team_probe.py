"""Dev probe: search width (warps per query) sweep on one index."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import embeddinghub_b200 as ehb
N = int(sys.argv[1]); d = int(sys.argv[2]); Q = int(sys.argv[3]); efs = [int(x) for x in sys.argv[4].split(",")]
metric = sys.argv[5] if len(sys.argv) > 5 else "l2"
k = 10
rng = np.random.default_rng(1234)
base = np.empty((N, d), np.float32)
for i in range(0, N, 1 << 20):
    base[i:i + (1 << 20)] = rng.standard_normal((min(1 << 20, N - i), d), dtype=np.float32)
q = np.random.default_rng(4321).standard_normal((Q, d), dtype=np.float32)
ix = ehb.NativeIndex(d, metric=metric, capacity=N)
ix.add(base); t = time.time(); ix.build(); print(f"build {time.time()-t:.2f}s", flush=True)
gt, _, _ = ix.search_bruteforce(q, k)
ref = {}
for ef in efs:
    for T in (1, 2, 3, 4):
        ix.set_search_width(T)
        best = 1e9
        for rep in range(5):
            l, dd, c = ix.search(q, k, ef=ef)
            best = min(best, ix.last_kernel_ms())
        st = ix.stats()
        rec = np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(l, gt)])
        if T == 1: ref[ef] = l
        same = np.mean(l == ref[ef])
        print(f"ef={ef} T={T}: {best:.3f} ms ({Q/best*1e3:.0f} qps) recall {rec:.4f} ids==T1 {same:.4f} evals/q {st['dist_evals']/Q:.0f} "
              f"hops/q {st['hops_base']/Q:.0f} ovf {st['visited_overflow']} -> {st['algorithmic_bytes']/best/1e6:.0f} GB/s sorted={bool(np.all(np.diff(dd,axis=1)>=0))} cnt_ok={bool(np.all(c==k))}", flush=True)
