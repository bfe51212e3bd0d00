#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2g}
L=gpurun_out/${TAG}_variants.log
: > $L
python -m pytest tests -m gpu -x -q -k "entropy" > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log; tail -3 gpurun_out/${TAG}_tests.log
python tools/dev_encode_bench.py 20000 6 >> $L 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_zenc|k_svb|k_exclusive" -s 21 -c 9 --csv --log-file gpurun_out/${TAG}_enc_launches.csv python tools/dev_encode_bench.py 20000 6 --steps 3 > gpurun_out/${TAG}_ncu_enc.log 2>&1
python tools/dev_encode_bench.py 20000 0 --model R9.4.1 >> $L 2>&1
cat $L
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/${TAG}_enc_launches.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
for r in rows[1:]: print(r[ki][:50], r[vi])
PY
