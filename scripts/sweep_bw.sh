#!/bin/bash
# granularity of the exhaustive work queue: default heuristic vs forced (b-window, x tiles per item)
for n in 40 90 150 300 600 1000; do
for k in "0 0" "32 1" "32 2" "32 4" "16 1"; do
  set -- $k
  echo "== n=$n BW=$1 XCH=$2"
  PIPSORT_EXH_BW=$1 PIPSORT_EXH_XCH=$2 python scripts/prof_one.py $n 5 | sed 's/total.*//'
done
done
