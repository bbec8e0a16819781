#!/bin/bash
out=gpurun_out
{ for cfg in "768 1" "512 1"; do set -- $cfg; echo "NC=$1 G=$2"; MGA_S4_NC=$1 MGA_S4_G=$2 python profiles/k4_time.py; done
MGA_S4=0 python profiles/k4_time.py; } 2>&1 | grep -v Warning | tee $out/r02k_k4.log
timeout 600 python -m pytest tests -m gpu -q -x -k "long or tiled or t288 or streaming" 2>&1 | tail -3
