#!/usr/bin/env python
"""Save / load throughput of the streaming persistence (io.cu): GB/s of file <-> HBM through two pinned
64 MB buffers, against the reference's cold start (one addPoint per row, version.cc:64-74).
  python tools/io_probe.py [N] [d] [metric] [dir]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import embeddinghub_b200 as ehb  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
    metric = sys.argv[3] if len(sys.argv) > 3 else "ip"
    where = sys.argv[4] if len(sys.argv) > 4 else "/tmp"
    rng = np.random.default_rng(1234)
    ix = ehb.NativeIndex(d, metric=metric, capacity=N)
    for i in range(0, N, 1 << 20):
        m = min(1 << 20, N - i)
        ix.add(rng.standard_normal((m, d), dtype=np.float32), np.arange(i, i + m, dtype=np.uint64))
    t = time.time()
    ix.build()
    t_build = time.time() - t
    q = np.random.default_rng(4321).standard_normal((100, d), dtype=np.float32)
    a = ix.search(q, 10, ef=64)
    path = os.path.join(where, "ehb200_io_probe.ehb")
    t = time.time()
    ix.save(path)
    t_save = time.time() - t
    size = os.path.getsize(path)
    del ix
    t = time.time()
    ix2 = ehb.NativeIndex.load(path)
    t_load = time.time() - t
    b = ix2.search(q, 10, ef=64)
    os.remove(path)
    print(json.dumps({"N": N, "d": d, "file_GB": size / 1e9, "gpu_build_s": round(t_build, 2),
                      "save_s": round(t_save, 2), "save_GBps": size / 1e9 / t_save,
                      "load_s": round(t_load, 2), "load_GBps": size / 1e9 / t_load,
                      "same_results_after_load": bool(np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])),
                      "dir": where, "note": "load includes the page-cache read of the file just written"}))


if __name__ == "__main__":
    main()
