#!/bin/bash
tag=${1:-r02h}
out=gpurun_out
for cfg in "768 2" "768 1" "512 1"; do set -- $cfg; MGA_S4_NC=$1 MGA_S4_G=$2 timeout 300 python profiles/bench_configs.py t288 --mode streaming --steps 3; done > $out/${tag}_t288.jsonl 2> $out/${tag}_t288.err
cut -c1-200 $out/${tag}_t288.jsonl; tail -3 $out/${tag}_t288.err
timeout 600 python -m pytest tests -m gpu -q -x -k "long or tiled or t288 or streaming" > $out/${tag}_pytest.log 2>&1; tail -3 $out/${tag}_pytest.log
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file $out/${tag}_launches_t288.csv python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_launches_t288.log 2>&1
python profiles/launch_bw.py $out/${tag}_launches_t288.csv 2>/dev/null | head -8
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k4_cg" --launch-skip 8 -c 1 \
  -o $out/${tag}_k4_t288 -f python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_k4_ncu.log 2>&1
