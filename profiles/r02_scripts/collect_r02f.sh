#!/bin/bash
tag=${1:-r02f}
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -6 $out/${tag}_pytest.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k4_cg" --launch-skip 8 -c 2 \
  -o $out/${tag}_k4_t288 -f python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_k4_ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
  --log-file $out/${tag}_launches_t288.csv python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_launches_t288.log 2>&1
python profiles/launch_summary.py $out/${tag}_launches_t288.csv 2>/dev/null | head -30
