#!/bin/bash
# Resident kernel timing experiment: in-list lengths per warp (sorted by in-degree: first warp 2.5 x the average) cut to <= 7 steps
# (MGA_RES_X=4) or set to 6 for every warp (8: the balanced total 60 ~ the real 65).  Wrong results; times only.
out=gpurun_out
for v in "" rx4 rx8; do
  echo "== ${v:-default}"
  lib=$PWD/mixed_graph_admm_b200/_lib/$v/libmga.so
  MGA_SCHED_VERBOSE=1 MGA_LIB=$lib python profiles/profile_step.py --mode resident --batch 1024 --steps 6 | tail -12
done > $out/r02v_variants.log 2>&1
cat $out/r02v_variants.log
