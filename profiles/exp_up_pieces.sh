#!/bin/bash
# (needs the experiment patch that reads MGA_HOST_UP_PIECES in solve_host_pipelined; the shipped library keeps 8 pieces - DESIGN 4.6)
# upload granularity of the pipelined host entry (MGA_HOST_UP_PIECES): timeline of combined_loop(y_pinned) at B = 1024
for p in 8 4 6 12 16; do
  echo "== MGA_HOST_UP_PIECES=$p"
  MGA_HOST_UP_PIECES=$p python profiles/e2e_trace.py 2>&1 | grep -E "host entry|python call" | tail -7
done
