"""Quick GPU probe: build + search timings, recall, achieved GB/s (dev tool)."""
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import embeddinghub_b200 as ehb


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    Q = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    metric = sys.argv[4] if len(sys.argv) > 4 else "l2"
    efs = [int(x) for x in (sys.argv[5].split(",") if len(sys.argv) > 5 else ["64"])]
    width = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    k = 10
    rng = np.random.default_rng(1234)
    base = np.empty((N, d), np.float32)
    for i in range(0, N, 1 << 20):
        base[i:i + (1 << 20)] = rng.standard_normal((min(1 << 20, N - i), d), dtype=np.float32)
    q = np.random.default_rng(4321).standard_normal((Q, d), dtype=np.float32)
    ix = ehb.NativeIndex(d, metric=metric, capacity=N)
    t = time.time(); ix.add(base); t_add = time.time() - t
    t = time.time(); ix.build(); t_build = time.time() - t
    print(f"N={N} d={d} add {t_add:.2f}s build {t_build:.2f}s ({N / t_build:.0f} ins/s)", flush=True)
    ix.set_search_width(width)
    t = time.time(); gt, _, _ = ix.search_bruteforce(q, k); t_bf = time.time() - t
    print(f"bruteforce Q={Q}: {t_bf:.3f}s wall, kernel {ix.last_kernel_ms():.2f} ms", flush=True)
    for ef in efs:
        for rep in range(3):
            t = time.time(); l, dd, c = ix.search(q, k, ef=ef); wall = time.time() - t
        ms = ix.last_kernel_ms()
        st = ix.stats()
        rec = np.mean([len(set(a.tolist()) & set(b.tolist())) / k for a, b in zip(l, gt)])
        gbs = st["algorithmic_bytes"] / (ms * 1e-3) / 1e9
        print(f"ef={ef}: kernel {ms:.3f} ms ({Q / ms * 1e3:.0f} qps) wall {wall * 1e3:.2f} ms recall@{k} {rec:.4f} "
              f"evals/q {st['dist_evals'] / Q:.0f} hops/q {st['hops_base'] / Q:.0f} up {st['hops_upper'] / Q:.1f} "
              f"ovf {st['visited_overflow']} alg {st['algorithmic_bytes'] / 1e6:.1f} MB -> {gbs:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
