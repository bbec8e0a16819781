#!/bin/bash
# A/B of the resident kernel variants (chunk hand-shake / ld.cg compiled out) + the end-to-end call
out=gpurun_out
for v in v_cur v_nopipe v_noldcg v_none; do
  echo "== $v"
  MGA_LIB=$PWD/mixed_graph_admm_b200/_lib/$v/libmga.so python profiles/profile_step.py --mode resident --batch 1024 --steps 6 | tail -3
done > $out/r02b_variants.log 2>&1
cat $out/r02b_variants.log
MGA_LIB=$PWD/mixed_graph_admm_b200/_lib/v_cur/libmga.so python profiles/e2e_profile.py > $out/r02b_e2e.log 2>&1
tail -40 $out/r02b_e2e.log
