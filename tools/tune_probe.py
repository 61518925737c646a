"""Dev probe: sweep staging / hash tuning knobs on one index."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import embeddinghub_b200 as ehb
N = int(sys.argv[1]); d = int(sys.argv[2]); Q = int(sys.argv[3]); ef = int(sys.argv[4]); metric = sys.argv[5]
configs = [tuple(int(x) for x in c.split(":")) for c in sys.argv[6].split(",")]  # slots:groups:hash_bits:width
k = 10
rng = np.random.default_rng(1234)
base = np.empty((N, d), np.float32)
for i in range(0, N, 1 << 20):
    base[i:i + (1 << 20)] = rng.standard_normal((min(1 << 20, N - i), d), dtype=np.float32)
q = np.random.default_rng(4321).standard_normal((Q, d), dtype=np.float32)
ix = ehb.NativeIndex(d, metric=metric, capacity=N)
ix.add(base); t = time.time(); ix.build(); print(f"build {time.time()-t:.2f}s", flush=True)
gt, _, _ = ix.search_bruteforce(q[:1000], k)
for (slots, groups, hb, width) in configs:
    ix.set_tuning(slots, groups, hb, 0); ix.set_search_width(width)
    best = 1e9
    for rep in range(3):
        l, dd, c = ix.search(q, k, ef=ef); best = min(best, ix.last_kernel_ms())
    st = ix.stats()
    rec = np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(l[:1000], gt)])
    print(f"G={slots} NG={groups} hb={hb} T={width}: {best:.3f} ms ({Q/best*1e3:.0f} qps) recall {rec:.4f} evals/q {st['dist_evals']/Q:.0f} ovf {st['visited_overflow']} -> {st['algorithmic_bytes']/best/1e6:.0f} GB/s", flush=True)
