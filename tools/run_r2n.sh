#!/bin/bash
# round-2 baseline of the current tree: GPU tests, decode / encode timing, ncu --set full of the decode step (20k-read C5)
set -u
mkdir -p gpurun_out
TAG=${1:-r2n}
L=gpurun_out/${TAG}.log
: > $L
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
python tools/decode_bench.py 20000 6 --model R10.4.1 --phases --cache /tmp/c5.npz >> $L 2>&1
python tools/decode_bench.py 20000 0 --model R9.4.1 --cache /tmp/vbz.npz >> $L 2>&1
python tools/dev_encode_bench.py 20000 6 >> $L 2>&1
python tools/dev_encode_bench.py 20000 0 >> $L 2>&1
cat $L
ncu --set full --clock-control none --import-source on -k regex:"k_decode_fast|k_plan" -s 12 -c 4 -o gpurun_out/${TAG}_dec python tools/decode_bench.py 20000 6 --model R10.4.1 --cache /tmp/c5.npz > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
