#!/bin/bash
# Round 2, GPU call A: the whole GPU test-suite, the bench line (N = 1) with its probes, the reference arm, and a fresh
# ncu --set full capture of the resident kernel (-> counters JSON for bench.py's shared-memory roofline).
tag=${1:-r02a}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
tail -3 $out/${tag}_pytest.log
timeout 600 python bench.py > $out/${tag}_bench_n1.log 2> $out/${tag}_bench_n1.err; echo "bench exit $?"
tail -1 $out/${tag}_bench_n1.log | cut -c1-600
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.log 2>&1
timeout 300 python profiles/profile_step.py --mode resident --batch 1024 --steps 3 > $out/${tag}_profile_step.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_admm_resident --launch-skip 1 -c 1 \
  -o $out/${tag}_resident -f python profiles/profile_step.py --mode resident --batch 1024 --steps 2 > $out/${tag}_resident_ncu.log 2>&1 && \
python profiles/refresh_counters.py $out/${tag}_resident.ncu-rep 1024 $out/${tag}_resident_counters.json > $out/${tag}_counters.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_step.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cg-probe --no-probes --min-timed-s 0 > $out/${tag}_launches_stdout.log 2>&1
echo finished
