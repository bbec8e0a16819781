# final refresh of the round-2 evidence after the deferred x update: bench line, k4 ncu capture, launch lists (T = 288, 20 000 nodes)
tag=r02z; out=gpurun_out
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err || exit 1
python profiles/bench_configs.py t288 --mode streaming --steps 1 > /dev/null 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k4_cg --launch-skip 8 -c 1 -o $out/${tag}_k4_t288 -f python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_k4_ncu.log 2>&1
{ python profiles/ncu_summary.py $out/${tag}_k4_t288.ncu-rep; echo; echo "---- stall samples by source line (ncu source page, -lineinfo)"; python profiles/ncu_lines.py $out/${tag}_k4_t288.ncu-rep; } > $out/${tag}_k4_t288_ncu_summary.txt 2> $out/${tag}_k4_err.txt
rm -f $out/${tag}_k4_t288.ncu-rep
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches_t288.csv python profiles/bench_configs.py t288 --mode streaming --steps 1 > $out/${tag}_launches_t288.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file $out/${tag}_launches_large20k.csv python profiles/bench_configs.py large20k --mode streaming --steps 1 > $out/${tag}_launches_large20k.log 2>&1
{ python profiles/bench_configs.py pems08 pems04 pems04_t24 t288 large20k; python profiles/bench_configs.py pems04 pems04_t24 n600_t96 pems07_t288 --mode streaming; } > $out/${tag}_configs.jsonl 2> $out/${tag}_configs.err
tail -c 300 $out/${tag}_bench_n1.json
