#!/bin/bash
# config 4 (long-expansion stress) stage times under different resident-warp settings of the global tier
for w in 16 24; do echo "== TDB_GLOBAL_WARPS_PER_SM=$w"; TDB_GLOBAL_WARPS_PER_SM=$w python profiles/config4_stages.py 2>&1 | grep "call ms" | tail -1 | cut -c1-200; done
