#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-ncuplan}
ncu --set full --clock-control none --import-source on -k regex:"k_plan" -s 3 -c 1 -o gpurun_out/${TAG}_plan python tools/decode_bench.py 20000 6 --model R10.4.1 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
