#!/bin/bash
# 2-GPU call: NCCL-sharded parity test + the N = 2 bench line (ONE batch of 65536 sharded)
tag=${1:-r02p}
out=gpurun_out
nvidia-smi -L > $out/${tag}_gpus.txt
timeout 600 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -k nccl > $out/${tag}_pytest_nccl.log 2>&1; tail -3 $out/${tag}_pytest_nccl.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $out/${tag}_bench_n2.log 2> $out/${tag}_bench_n2.err; echo "bench exit $?"
tail -1 $out/${tag}_bench_n2.log | cut -c1-1500
tail -3 $out/${tag}_bench_n2.err
