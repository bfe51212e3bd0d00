#!/bin/bash
# planner timing: launch list of one decode step (20k reads) + decode bench
set -u
mkdir -p gpurun_out
TAG=${1:-plan}
python tools/decode_bench.py 20000 6 --model R10.4.1 --cache /tmp/c5.npz 2>&1 | tail -1
python tools/decode_bench.py 20000 0 --model R9.4.1 --cache /tmp/vbz.npz 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_plan|k_decode_fast" -s 8 -c 4 --csv --log-file gpurun_out/${TAG}_launch.csv python tools/decode_bench.py 20000 6 --model R10.4.1 --cache /tmp/c5.npz > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${TAG}_launch.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]; ix={k:i for i,k in enumerate(hdr)}
for r in rows[h+1:]:
    if len(r)>=len(hdr): print(r[ix['Kernel Name']][:36], r[ix['Metric Name']], r[ix['Metric Value']])
PY
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
