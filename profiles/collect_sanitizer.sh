#!/bin/bash
# compute-sanitizer over the shared-memory kernels: racecheck (hazards between the phases that reuse pbuf / qbuf / the
# TMA-staged tiles) and memcheck, on small batches (the tools run the kernels 10-100x slower)
tag=${1:-r02q}
out=gpurun_out
run() { name=$1; shift; timeout 600 compute-sanitizer "$@" > $out/${tag}_${name}.log 2>&1; echo "$name exit $?"; grep -E "RACECHECK SUMMARY|ERROR SUMMARY|hazard|Error" $out/${tag}_${name}.log | head -5; }
run race_resident --tool racecheck --racecheck-report all python profiles/profile_step.py --mode resident --batch 6 --steps 1
run race_resident_host --tool racecheck --racecheck-report all python profiles/e2e_profile.py 70
run race_k4 --tool racecheck --racecheck-report all python profiles/bench_configs.py t288 --mode streaming --steps 0 --batch 2
run race_nodetile --tool racecheck --racecheck-report all python profiles/bench_configs.py pems07_t288 --mode streaming --steps 0 --batch 1
run mem_k4 --tool memcheck python profiles/bench_configs.py t288 --mode streaming --steps 0 --batch 2
run mem_resident --tool memcheck python profiles/profile_step.py --mode resident --batch 6 --steps 1
run sync_k4 --tool synccheck python profiles/bench_configs.py t288 --mode streaming --steps 0 --batch 2
