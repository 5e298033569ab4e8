#!/bin/bash
# global-scratch tier: launch bounds (registers) x resident warps, on the long-expansion stress set
for b in 4 5 6; do for w in 16 20 24; do
  [ $b = 4 ] && [ $w != 16 ] && continue
  [ $b = 5 ] && [ $w = 24 ] && continue
  echo -n "minb=$b warps=$w: "; TDB_LIB=$PWD/build/sweep/lib_gminb$b.so TDB_GLOBAL_WARPS_PER_SM=$w python profiles/config4_stages.py 2>&1 | tail -2 | head -1 | cut -c1-130
done; done
