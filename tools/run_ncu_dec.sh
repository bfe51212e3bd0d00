#!/bin/bash
# ncu --set full of one decode launch group (20k-read C5 R10.4.1): tools/run_ncu_dec.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-ncu}
ncu --set full --clock-control none --import-source on -k regex:"k_decode_fast|k_plan" -s 12 -c 4 -o gpurun_out/${TAG}_dec python tools/decode_bench.py 20000 6 --model R10.4.1 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
