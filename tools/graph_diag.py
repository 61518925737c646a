"""Dev diagnostic: GPU-built graph vs oracle-built graph quality (degree stats, cross searches)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import embeddinghub_b200 as ehb
from oracle import oracle as orc

N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
k = 10
base = np.random.default_rng(1234).standard_normal((N, d), dtype=np.float32)
q = np.random.default_rng(4321).standard_normal((500, d), dtype=np.float32)
rec = lambda a, b: float(np.mean([len(set(x.tolist()) & set(y.tolist())) / k for x, y in zip(a, b)]))

ix = ehb.NativeIndex(d, capacity=N)
ix.add(base); t = time.time(); ix.build(); print("gpu build", time.time() - t)
gt, _, _ = ix.search_bruteforce(q, k)
gg = ix.export_graph()
o = orc.OracleHNSW(d, "l2", N)
t = time.time(); o.add(base, threads=8); print("oracle build", time.time() - t)
og = o.export_graph()

def deg(g, name):
    l0 = g["links0"]; dg = (l0 != 0xFFFFFFFF).sum(1)
    print(f"{name}: level0 degree mean {dg.mean():.2f} min {dg.min()} p10 {np.percentile(dg,10):.0f} median {np.median(dg):.0f} max {dg.max()}  full(32) {np.mean(dg==32):.3f}  levels>0: {np.mean(g['levels']>0):.4f} maxlevel {g['maxlevel']}")
    # in-degree
    ids = l0[l0 != 0xFFFFFFFF]; indeg = np.bincount(ids, minlength=N)
    print(f"   in-degree mean {indeg.mean():.2f} zero {np.mean(indeg==0):.4f} max {indeg.max()}")
    # mean edge length
    src = np.repeat(np.arange(N), 32).reshape(N, 32)[l0 != 0xFFFFFFFF]
    dist = ((base[src] - base[ids]) ** 2).sum(1)
    print(f"   mean edge dist {dist.mean():.3f}")
deg(gg, "gpu"); deg(og, "oracle")

# cross searches on CPU oracle walker
o2 = orc.OracleHNSW(d, "l2", N); o2.import_graph(gg)
ix2 = ehb.NativeIndex(d, capacity=N); ix2.import_graph(og)
for ef in (16, 64, 128):
    a = rec(o.search(q, k, ef=ef)[0], gt)       # oracle graph, oracle walk
    b = rec(o2.search(q, k, ef=ef)[0], gt)      # gpu graph, oracle walk
    c = rec(ix.search(q, k, ef=ef)[0], gt)      # gpu graph, gpu walk
    e = rec(ix2.search(q, k, ef=ef)[0], gt)     # oracle graph, gpu walk
    print(f"ef={ef}: oracleG/oracleW {a:.4f}  gpuG/oracleW {b:.4f}  gpuG/gpuW {c:.4f}  oracleG/gpuW {e:.4f}")
