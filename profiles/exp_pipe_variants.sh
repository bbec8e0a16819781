#!/bin/bash
# (needs the experiment builds --tag=ldcg0 -DMGA_RES_LDCG=0 / --tag=nofence and the MGA_HOST_COEF_CHUNK patch; none is in the shipped library - DESIGN 4.6)
# what the PIPE instantiation of the resident kernel pays for (timing only; the variants are not correct builds):
# ldcg0 = y through L1 (plain loads), nofence = hand-over without fence + barrier
for v in main ldcg0 nofence; do
  if [ $v = main ]; then unset MGA_LIB; else export MGA_LIB=$PWD/mixed_graph_admm_b200/_lib/$v/libmga.so; fi
  for coef in 1 0; do
    echo "== $v MGA_HOST_COEF_CHUNK=$coef"
    MGA_HOST_COEF_CHUNK=$coef python profiles/e2e_trace.py 2>&1 | grep -E "host entry|python call" | head -12 | tail -6
  done
done
