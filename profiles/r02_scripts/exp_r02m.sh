#!/bin/bash
out=gpurun_out
{ for nc in 768 512; do echo "NC=$nc"; MGA_S4_NC=$nc python profiles/k4_time.py; done
MGA_S4=0 python profiles/k4_time.py; } 2>&1 | grep -v Warning | tee $out/r02m_k4.log
timeout 600 python -m pytest tests -m gpu -q -x -k "long or tiled or t288 or streaming" 2>&1 | tail -3
