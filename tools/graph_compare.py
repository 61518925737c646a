"""Dev diagnostic: with build_batch=1 the GPU build should reproduce the sequential oracle graph."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import embeddinghub_b200 as ehb
from oracle import oracle as orc
N = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 16
bb = int(sys.argv[3]) if len(sys.argv) > 3 else 1
base = np.random.default_rng(1234).standard_normal((N, d), dtype=np.float32)
ix = ehb.NativeIndex(d, capacity=N, build_batch=bb)
ix.add(base); ix.build()
gg = ix.export_graph()
o = orc.OracleHNSW(d, "l2", N); o.add(base, threads=1)
og = o.export_graph()
print("levels equal:", np.array_equal(gg["levels"], og["levels"]), "entry", gg["entry"], og["entry"], "maxlevel", gg["maxlevel"], og["maxlevel"])
def rowsets(l):
    return [frozenset(int(x) for x in r if x != 0xFFFFFFFF) for r in l]
a, b = rowsets(gg["links0"]), rowsets(og["links0"])
same = np.array([x == y for x, y in zip(a, b)])
print("level-0 rows identical:", same.mean(), " mean |sym diff|:", np.mean([len(x ^ y) for x, y in zip(a, b)]))
bad = np.nonzero(~same)[0]
print("first differing nodes:", bad[:10])
for i in bad[:3]:
    print(" node", i, "level", gg["levels"][i], "gpu-only", sorted(a[i] - b[i]), "oracle-only", sorted(b[i] - a[i]), "deg", len(a[i]), len(b[i]))
if len(gg["links_up"]):
    ua, ub = rowsets(gg["links_up"]), rowsets(og["links_up"])
    print("upper rows identical:", np.mean([x == y for x, y in zip(ua, ub)]), "rows", len(ua), len(ub), "up_off equal", np.array_equal(gg["up_off"], og["up_off"]))
