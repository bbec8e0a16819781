#!/bin/bash
out=gpurun_out
{ python profiles/k4_time.py
for x in 1 2 4 8 16 32 63; do MGA_LIB=$PWD/mixed_graph_admm_b200/_lib/x$x/libmga.so python profiles/k4_time.py; done
MGA_S4=0 python profiles/k4_time.py; } 2>&1 | grep -v Warning | tee $out/r02j_k4_variants.log
