"""Dev probe: exact fp32 vs bf16 tensor-core brute force."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import embeddinghub_b200 as ehb
from embeddinghub_b200._native import BF16, FP32
N = int(sys.argv[1]); d = int(sys.argv[2]); Q = int(sys.argv[3]); k = int(sys.argv[4]); metric = sys.argv[5]
rng = np.random.default_rng(1234)
base = np.empty((N, d), np.float32)
for i in range(0, N, 1 << 20):
    base[i:i + (1 << 20)] = rng.standard_normal((min(1 << 20, N - i), d), dtype=np.float32)
q = np.random.default_rng(4321).standard_normal((Q, d), dtype=np.float32)
ix = ehb.NativeIndex(d, metric=metric, capacity=N, ef_construction=16)
ix.add(base)
res = {}
for name, prec in (("bf16", BF16), ("bf16", BF16), ("fp32", FP32)):
    t = time.time(); l, dd, c = ix.search_bruteforce(q, k, precision=prec); wall = time.time() - t
    ms = ix.last_kernel_ms()
    res[name] = l
    print(f"{name}: kernel {ms:.2f} ms wall {wall*1e3:.1f} ms  {Q/ms*1e3:.0f} qps  {2*Q*N*d/ms/1e9:.1f} TFLOP/s", flush=True)
rec = np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(res["bf16"], res["fp32"])])
print("recall bf16 vs exact", rec, "ids equal", np.mean(res["bf16"] == res["fp32"]))
