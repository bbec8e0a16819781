#!/bin/bash
tag=${1:-r02c}
out=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
python profiles/e2e_profile.py > $out/${tag}_e2e.log 2>&1; head -8 $out/${tag}_e2e.log
timeout 600 python bench.py > $out/${tag}_bench_n1.log 2> $out/${tag}_bench_n1.err; echo "bench exit $?"
tail -1 $out/${tag}_bench_n1.log | cut -c1-400
