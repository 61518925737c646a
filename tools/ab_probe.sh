#!/bin/bash
# A/B of two library builds on the three walk shapes (kernel ms from CUDA events, perf_probe.py)
for lib in "" "$PWD/tools/ab/libehb200_v1.so"; do
  echo "=== lib=${lib:-default}"
  EHB200_LIB=$lib python tools/perf_probe.py 1000000 128 1000 l2 64 2>&1 | grep -E "^ef="
  EHB200_LIB=$lib python tools/perf_probe.py 1000000 128 10000 cosine 256 2>&1 | grep -E "^ef="
  EHB200_LIB=$lib python tools/perf_probe.py 1000000 768 10000 ip 128 2>&1 | grep -E "^ef="
done
