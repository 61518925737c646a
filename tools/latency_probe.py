"""Dev probe: small-batch latency vs warps per query."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import embeddinghub_b200 as ehb
N, d = 1000000, 128
base = np.random.default_rng(1234).standard_normal((N, d), dtype=np.float32)
ix = ehb.NativeIndex(d, capacity=N); ix.add(base); ix.build()
for Q in (1, 16, 100, 444):
    q = np.random.default_rng(4321).standard_normal((Q, d), dtype=np.float32)
    row = []
    for T in (1, 2, 4):
        ix.set_search_width(T)
        best = 1e9; wall = 1e9
        for _ in range(10):
            t = time.perf_counter(); ix.search(q, 10, ef=64); wall = min(wall, time.perf_counter() - t)
            best = min(best, ix.last_kernel_ms())
        row.append(f"T={T}: kernel {best*1e3:.0f} us, call {wall*1e6:.0f} us")
    print(f"Q={Q}: " + " | ".join(row), flush=True)
