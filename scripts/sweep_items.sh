#!/bin/bash
# granularity sweep of the exhaustive work queue (items per resident warp) on the B150c3 / B1500c3 loci
for k in 1 2 3 4 6 8; do
  echo "== PIPSORT_EXH_ITEMS_PER_SLOT=$k"
  PIPSORT_EXH_ITEMS_PER_SLOT=$k python scripts/prof_one.py 150 8 | sed 's/total.*//'
done
for k in 2 6; do
  echo "== PIPSORT_EXH_ITEMS_PER_SLOT=$k (1500)"
  PIPSORT_EXH_ITEMS_PER_SLOT=$k python scripts/prof_one.py 1500 2 | sed 's/total.*//'
done
