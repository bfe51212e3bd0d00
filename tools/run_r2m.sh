#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2m}
L=gpurun_out/${TAG}.log
: > $L
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
python tools/encode_bench.py 20000 6,0 --check >> $L 2>&1
python tools/dev_encode_bench.py 20000 6 >> $L 2>&1
cat $L
python tools/encode_bench.py 6000 6 > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"k_svb_encode1" -s 3 -c 1 -o gpurun_out/${TAG}_enc1 python tools/encode_bench.py 6000 6 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
