#!/bin/bash
# One GPU call that refreshes the evidence of a round (run under gpurun from the repo root):
#   bash profiles/collect_round.sh r01j
# bench line, other configs (auto + streaming), variants, ncu --set full of the time-tiled streaming kernels,
# ncu launch list of the bench step.  Each profiler pass runs only after the plain command exited 0.
tag=${1:-r01}
out=gpurun_out
set -x
python bench.py > $out/${tag}_bench_n1.log 2> $out/${tag}_bench_n1.err || exit 1
{ python profiles/bench_configs.py pems08 pems04 pems04_t24 t288 large20k;
  python profiles/bench_configs.py pems04 pems04_t24 n600_t96 pems07_t288 --mode streaming; } > $out/${tag}_configs.jsonl 2> $out/${tag}_configs.err || exit 1
python profiles/bench_variants.py > $out/${tag}_variants.jsonl 2> $out/${tag}_variants.err || exit 1
# (-k matches the base name: skip the initial-residual launches of the first solve to land on CG iterations)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k3_p_ldr|k3_ldrt_lhs" --launch-skip 6 -c 2 \
  -o $out/${tag}_k3_t288 -f python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_k3_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k3_lu" --launch-skip 3 -c 1 \
  -o $out/${tag}_k3lu_t288 -f python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_k3lu_ncu.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_step.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cg-probe > $out/${tag}_launches_stdout.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
  --log-file $out/${tag}_launches_t288.csv python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_launches_t288.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv \
  --log-file $out/${tag}_launches_large20k.csv python profiles/bench_configs.py large20k --mode streaming --steps 1 > $out/${tag}_launches_large20k.log 2>&1
tail -1 $out/${tag}_bench_n1.log | cut -c1-300
