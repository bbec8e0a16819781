#!/bin/bash
tag=${1:-r02s}
out=gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_admm_cluster" --launch-skip 2 -c 1 \
  -o $out/${tag}_cluster -f python profiles/notebook_call.py > $out/${tag}_cluster_ncu.log 2>&1
tail -3 $out/${tag}_cluster_ncu.log
