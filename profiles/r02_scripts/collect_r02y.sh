#!/bin/bash
tag=${1:-r02y}
out=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "long or tiled or t288 or streaming or host_entry_streaming" > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log | cut -c1-300
for s5 in 1 0; do MGA_S5=$s5 timeout 300 python profiles/bench_configs.py t288 --mode streaming --steps 3; done 2>/dev/null | cut -c1-220
MGA_S5=1 timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file $out/${tag}_launches_t288.csv python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_launches_t288.log 2>&1
python profiles/launch_bw.py $out/${tag}_launches_t288.csv 2>/dev/null | head -12
