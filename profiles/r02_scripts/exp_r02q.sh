#!/bin/bash
# Resident kernel timing experiments: reductions compiled out (MGA_RES_X bit 1: <r,r>, bit 2: <p,Ap>) -> how much the
# sync points of a CG iteration cost.  Results are WRONG by construction; only the times count.
out=gpurun_out
for v in "" rx1 rx2 rx3; do
  echo "== ${v:-default}"
  lib=$PWD/mixed_graph_admm_b200/_lib/$v/libmga.so
  MGA_LIB=$lib python profiles/profile_step.py --mode resident --batch 1024 --steps 6 | tail -2
  MGA_LIB=$lib python profiles/profile_step.py --mode resident --batch 8192 --steps 4 | tail -1
done > $out/r02q_variants.log 2>&1
cat $out/r02q_variants.log
