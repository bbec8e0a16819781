#!/bin/bash
# One GPU call that refreshes the evidence of a round (run under gpurun from the repo root):
#   bash profiles/collect_round.sh r02
# GPU test-suite, bench line (+ reference arm), other configs, variants, ncu --set full of the dominant kernels (resident,
# k4_cg, k5_tail, cluster), ncu launch lists (bench step, T = 288, 20 000 nodes).  Each profiler pass runs only after the
# plain command exited 0.
tag=${1:-r02}
out=gpurun_out
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $out/${tag}_gpu.txt
timeout 900 python -m pytest tests -m gpu -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench_n1.log 2> $out/${tag}_bench_n1.err || exit 1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.log 2>&1
{ python profiles/bench_configs.py pems08 pems04 pems04_t24 t288 large20k;
  python profiles/bench_configs.py pems04 pems04_t24 n600_t96 pems07_t288 --mode streaming; } > $out/${tag}_configs.jsonl 2> $out/${tag}_configs.err
python profiles/bench_variants.py > $out/${tag}_variants.jsonl 2> $out/${tag}_variants.err
python profiles/notebook_call.py > $out/${tag}_notebook.json 2>/dev/null
# ---- ncu --set full, one kernel per capture
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_admm_resident --launch-skip 1 -c 1 \
  -o $out/${tag}_resident -f python profiles/profile_step.py --mode resident --batch 1024 --steps 2 > $out/${tag}_resident_ncu.log 2>&1 && \
python profiles/refresh_counters.py $out/${tag}_resident.ncu-rep 1024 $out/${tag}_resident_counters.json > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k4_cg --launch-skip 8 -c 1 \
  -o $out/${tag}_k4_t288 -f python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_k4_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k5_tail" --launch-skip 2 -c 1 \
  -o $out/${tag}_k5_t288 -f python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_k5_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_admm_cluster" --launch-skip 2 -c 1 \
  -o $out/${tag}_cluster -f python profiles/notebook_call.py > $out/${tag}_cluster_ncu.log 2>&1
# ---- summaries on the box (the reports together exceed what a call may bring back): keep the text, drop the big reports
for k in resident k4_t288 k5_t288 cluster; do
  rep=$out/${tag}_$k.ncu-rep
  [ -f $rep ] || continue
  { python profiles/ncu_summary.py $rep; echo; echo "---- stall samples by source line (ncu source page, -lineinfo)"; python profiles/ncu_lines.py $rep; } \
    > $out/${tag}_${k}_ncu_summary.txt 2> $out/${tag}_${k}_ncu_summary.err
  [ $k = resident ] || rm -f $rep
done
# ---- launch lists (time + DRAM bytes per launch)
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/${tag}_launches_step.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cg-probe --no-probes --min-timed-s 0 > $out/${tag}_launches_stdout.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
  --log-file $out/${tag}_launches_t288.csv python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_launches_t288.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
  --log-file $out/${tag}_launches_large20k.csv python profiles/bench_configs.py large20k --mode streaming --steps 1 > $out/${tag}_launches_large20k.log 2>&1
tail -1 $out/${tag}_bench_n1.log | cut -c1-300
