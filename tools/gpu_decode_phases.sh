#!/bin/bash
# decode timing of the current tree: tools/gpu_decode_phases.sh TAG [ncu]  (20k-read C5 R10.4.1 + VBZ R9.4.1; optional ncu --set full of the decode kernels)
set -u
mkdir -p gpurun_out
TAG=${1:-dec}
L=gpurun_out/${TAG}.log
: > $L
python tools/decode_bench.py 20000 6 --model R10.4.1 --phases --cache /tmp/c5.npz >> $L 2>&1
python tools/decode_bench.py 20000 0 --model R9.4.1 --phases --cache /tmp/vbz.npz >> $L 2>&1
python tools/decode_bench.py 20000 6 --model R9.4.1 >> $L 2>&1
cat $L
if [ "${2:-}" = "ncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:"k_decode_fast|k_plan" -s 12 -c 4 -o gpurun_out/${TAG}_dec python tools/decode_bench.py 20000 6 --model R10.4.1 --cache /tmp/c5.npz > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
fi
