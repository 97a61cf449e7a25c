#!/bin/bash
# Decompose the top-K scoring kernel's time (TTAM_TOPK_DEBUG bits: 1 no drain, 2 no appends, 4 no TMA item loads,
# 8 MMA does not wait for the drain, 16 MMA-thread polls back off).  Results are garbage under any debug bit; only
# the timings mean something.
Q=${1:-37888}
for d in ${2:-0 16 2 1 9 13}; do
  echo "debug=$d: $(TTAM_TOPK_DEBUG=$d python scripts/time_topk.py $Q)"
done
