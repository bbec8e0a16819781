#!/bin/bash
# Build and run the host-side AddressSanitizer check of plan creation (see asan_plan_harness.cpp).  CPU only.
# usage: profiles/asan_plan.sh [N_SHORT] [N_LONG]      (needs _lib/*.o from a normal build)
set -e
cd "$(dirname "$0")/.."
W=${ASAN_WORK:-/tmp/asan}; mkdir -p $W
CS=mixed_graph_admm_b200/csrc; LIBD=mixed_graph_admm_b200/_lib
FL="-gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC,-fsanitize=address,-fno-omit-frame-pointer -I include -I $CS"
nvcc $FL -c $CS/mga_plan.cu -o $W/mga_plan.o &
nvcc -O1 -g -std=c++17 -Xcompiler -fsanitize=address,-fno-omit-frame-pointer -I include -I $CS -x c++ -c $CS/mga_schedule.cpp -o $W/mga_schedule.o &
nvcc -O1 -g -std=c++17 -Xcompiler -fsanitize=address,-fno-omit-frame-pointer -I include -I $CS -x c++ -c $CS/mga_knn.cpp -o $W/mga_knn.o &
nvcc -O1 -g -std=c++17 -Xcompiler -fsanitize=address,-fno-omit-frame-pointer -I include -c profiles/asan_plan_harness.cpp -o $W/harness.o &
wait
OTHER=$(ls $LIBD/*.o | grep -v -E "mga_plan.o|mga_schedule.o|mga_knn.o")
nvcc -gencode arch=compute_100a,code=sm_100a -cudart shared -Xcompiler -fsanitize=address $W/harness.o $W/mga_plan.o $W/mga_schedule.o $W/mga_knn.o $OTHER -o $W/asan_plan
python profiles/asan_plan_dump.py $W/descs ${1:-300} ${2:-200}
ASAN_OPTIONS=detect_leaks=1 $W/asan_plan $W/descs/*.bin
