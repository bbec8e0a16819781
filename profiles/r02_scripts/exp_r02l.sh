#!/bin/bash
out=gpurun_out
{ python profiles/k4_time.py
for x in 16 64 128 208; do MGA_LIB=$PWD/mixed_graph_admm_b200/_lib/x$x/libmga.so python profiles/k4_time.py; done; } 2>&1 | grep -v Warning | tee $out/r02l_k4_variants.log
