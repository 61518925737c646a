#!/bin/bash
# A/B of library builds on the walk shapes (kernel ms from CUDA events, perf_probe.py)
echo "=== r01 tree (sanity), C5 shape"
(cd tools/ab/r01 && python tools/perf_probe.py 1000000 128 10000 cosine 256 2>&1 | grep -E "^ef=")
for lib in "" "$PWD/tools/ab/libehb200_v1nat.so" "$PWD/tools/ab/libehb200_mb5u2.so"; do
  echo "=== lib=${lib:-default}"
  EHB200_LIB=$lib python tools/perf_probe.py 1000000 128 1000 l2 64 2>&1 | grep -E "^ef="
  EHB200_LIB=$lib python tools/perf_probe.py 1000000 128 10000 cosine 256 2>&1 | grep -E "^ef="
done
echo "=== default, C3 shape"
python tools/perf_probe.py 1000000 768 10000 ip 128 2>&1 | grep -E "^ef="
