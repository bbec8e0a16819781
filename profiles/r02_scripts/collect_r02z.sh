#!/bin/bash
tag=${1:-r02z}
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log | cut -c1-300
timeout 300 python profiles/bench_configs.py t288 large20k pems07_t288 --mode streaming --steps 3 2>/dev/null | cut -c1-230
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file $out/${tag}_launches_t288.csv python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_launches_t288.log 2>&1
python profiles/launch_bw.py $out/${tag}_launches_t288.csv 2>/dev/null | grep -E "kernel|k2_init|k5_tail|k2_rhs"
