#!/bin/bash
# A/B of library builds on the C5 shape (kernel ms from CUDA events, perf_probe.py)
for lib in "" "$PWD/tools/ab/libehb200_u2.so" "$PWD/tools/ab/libehb200_mb5u2w20.so"; do
  echo "=== lib=${lib:-default}"
  EHB200_LIB=$lib python tools/perf_probe.py 1000000 128 10000 cosine 256 2>&1 | grep -E "^ef="
  EHB200_LIB=$lib python tools/perf_probe.py 1000000 128 10000 l2 64 2>&1 | grep -E "^ef="
done
