#!/bin/bash
# Seeded fuzz search split over processes of 25 cases (a device fault poisons only its own block).
# usage: profiles/fuzz_blocks.sh FIRST LAST OUT
first=${1:-0}; last=${2:-400}; out=${3:-gpurun_out/fuzz_blocks.log}
: > "$out"
run_block() { MGA_FUZZ_FIRST=$1 MGA_FUZZ_CASES=${FUZZ_SHORT:-25} MGA_FUZZ_LONG_CASES=${FUZZ_LONG:-0} MGA_FUZZ_API_CASES=${FUZZ_API:-0} timeout 600 python -m pytest tests/test_gpu_fuzz.py -q -m gpu --timeout 120 -p no:cacheprovider 2>&1 | grep -E "^(FAILED|ERROR|[0-9]+ (passed|failed)|E  +(mixed|Assert|torch\.Acc|[A-Za-z]+Error))" | sed "s/^/[$1] /"; }
for ((s = first; s < last; s += 100)); do
  for ((k = s; k < s + 100 && k < last; k += 25)); do run_block $k >> "$out.$k" & done
  wait
done
cat "$out".* > "$out"; rm -f "$out".*
grep -c passed "$out"; grep -E "FAILED|failed" "$out" | head -60
