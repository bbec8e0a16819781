#!/bin/bash
# Resident kernel: two-barrier CG iteration (-DMGA_RES_CG2=1: all systems, 2: z_u only, 3: x / z_d only) against the
# three-barrier one (0)
out=gpurun_out
for v in cg0 cg2 cg3; do
  echo "== ${v:-default}"
  lib=$PWD/mixed_graph_admm_b200/_lib/$v/libmga.so
  MGA_LIB=$lib python profiles/profile_step.py --mode resident --batch 1024 --steps 6 | tail -2
  MGA_LIB=$lib python profiles/profile_step.py --mode resident --batch 8192 --steps 4 | tail -1
done > $out/r02t2_variants.log 2>&1
cat $out/r02t2_variants.log
