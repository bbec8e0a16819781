#!/bin/bash
tag=${1:-r02e}
out=gpurun_out
python profiles/e2e_trace.py > $out/${tag}_trace.log 2>&1; tail -12 $out/${tag}_trace.log
for s4 in 1 0; do MGA_S4=$s4 timeout 300 python profiles/bench_configs.py t288 pems07_t288 --mode streaming --steps 3; done > $out/${tag}_t288.jsonl 2> $out/${tag}_t288.err
cut -c1-330 $out/${tag}_t288.jsonl; tail -3 $out/${tag}_t288.err
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -4 $out/${tag}_pytest.log
