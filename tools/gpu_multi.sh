#!/bin/bash
# N-GPU check of the driver's launch line: strong scaling of the 10 GiB corpus + the sharded entry point's test.
# usage (under gpurun --gpus N): bash tools/gpu_multi.sh <tag> <N>
TAG=${1:-s}; N=${2:-2}; OUT=gpurun_out; mkdir -p $OUT
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 5 --warmup 3 > $OUT/bench_${N}gpu_$TAG.json 2> $OUT/bench_${N}gpu_$TAG.err; echo "bench N=$N rc=$?"
tail -n 3 $OUT/bench_${N}gpu_$TAG.err; head -c 1500 $OUT/bench_${N}gpu_$TAG.json; echo
timeout -k 10 300 python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 tests/test_gpu_configs.py -k "sharded" > $OUT/t_sharded_${N}gpu_$TAG.log 2>&1; echo "sharded tests: $(tail -n 1 $OUT/t_sharded_${N}gpu_$TAG.log)"
