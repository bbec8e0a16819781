#!/bin/bash
# A/B of experiment builds of libmga (mixed_graph_admm_b200/build.py --tag=...): same driver, same inputs.
#   profiles/exp_variants.sh <out.log> <tag> [<tag> ...]     ("" / default = _lib/libmga.so)
out=$1; shift
: > "$out"
for tag in "$@"; do
  if [ "$tag" = default ]; then lib=mixed_graph_admm_b200/_lib/libmga.so; else lib=mixed_graph_admm_b200/_lib/$tag/libmga.so; fi
  for B in 1024 8192; do
    echo "== $tag B=$B" >> "$out"
    MGA_RES_VERBOSE=1 MGA_LIB=$PWD/$lib timeout 300 python profiles/profile_step.py --mode resident --batch $B --steps 6 2>&1 \
      | awk '/resident</ {if (!seen) print; seen=1; next} {print}' | tail -8 >> "$out"
  done
done
