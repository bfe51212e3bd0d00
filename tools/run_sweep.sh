#!/bin/bash
# phase times as seen by different warps of the CTA (20k-read C5 set): tools/run_sweep.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-sweep}
L=gpurun_out/${TAG}.log
: > $L
for w in 0 1 2 4 8 12 16 20 24 28; do
  RNSC_PROF_WARP=$w python tools/decode_bench.py 20000 6 --model R10.4.1 --phases --cache /tmp/c5.npz --tag "warp-$w" >> $L 2>&1
done
grep -E "B_1024|A_512" $L
