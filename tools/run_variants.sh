#!/bin/bash
# GPU box: GPU parity tests, then the same cached workload decoded by library variants / planner settings (tuning aid).
set -u
mkdir -p gpurun_out
TAG=${1:-r2a}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -3 gpurun_out/${TAG}_tests.log
C=/tmp/wl_c5.npz
L=gpurun_out/${TAG}_variants.log
: > $L
python tools/decode_bench.py 20000 6 --model R10.4.1 --cache $C --lib tools/variants/librnsc_r1.so --phases --tag r1 >> $L 2>&1
for s in 0 6 12 24; do
  RNSC_COST_SYM=$s python tools/decode_bench.py 20000 6 --model R10.4.1 --cache $C --phases --tag new_sym$s >> $L 2>&1
done
V=/tmp/wl_vbz.npz
python tools/decode_bench.py 20000 0 --model R9.4.1 --cache $V --lib tools/variants/librnsc_r1.so --tag r1 >> $L 2>&1
python tools/decode_bench.py 20000 0 --model R9.4.1 --cache $V --tag new >> $L 2>&1
cat $L
