#!/bin/bash
# usage: bash tools/gpu_session.sh <tag>   — parity suites, ncu capture, variant A/B
TAG=${1:-s}; OUT=gpurun_out; mkdir -p $OUT
PYT="python -m pytest -m gpu -q -p no:cacheprovider --timeout 300 --timeout-method thread"
timeout -k 10 900 $PYT -x tests/test_gpu_parity.py > $OUT/t_parity_$TAG.log 2>&1; echo "parity rc=$?"; tail -3 $OUT/t_parity_$TAG.log
timeout -k 10 1200 $PYT tests/test_gpu_configs.py > $OUT/t_configs_$TAG.log 2>&1; echo "configs-tests rc=$?"; tail -6 $OUT/t_configs_$TAG.log
bash tools/ncu_full.sh $TAG en
bash tools/ab.sh $TAG
