#!/bin/bash
# device encode (GPU entropy stage): timing + launch list of one step (20k reads): tools/gpu_encode_launches.sh TAG [codec]
set -u
mkdir -p gpurun_out
TAG=${1:-enc}; CODEC=${2:-6}
python tools/dev_encode_bench.py 20000 $CODEC 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"^k_|k_svb|k_zenc" -s 14 -c 7 --csv --log-file gpurun_out/${TAG}_launch.csv python tools/dev_encode_bench.py 20000 $CODEC > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${TAG}_launch.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
hdr=rows[h]; ix={k:i for i,k in enumerate(hdr)}
cur={}
for r in rows[h+1:]:
    if len(r)>=len(hdr):
        k=(r[ix['ID']], r[ix['Kernel Name']][:40]); cur.setdefault(k,{})[r[ix['Metric Name']]]=r[ix['Metric Value']]+' '+r[ix['Metric Unit']]
for k,v in cur.items(): print(k[1], v)
PY
