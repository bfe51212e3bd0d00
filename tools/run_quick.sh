#!/bin/bash
# GPU tests + a reduced bench line (20k reads) -- sanity of the current tree
set -u
mkdir -p gpurun_out
TAG=${1:-quick}
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --reads 20000 --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench20k.json 2> gpurun_out/${TAG}_bench20k.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_bench20k.json'))
print(d['value'], d['e2e']['value'])
print(d['other_legs']['C5_splitter_encode'])
print(d['other_legs']['config4_VBZ_splitter_encode_R9.4.1'])
print(d['other_legs']['config3_encode_gpu_entropy']['value'])
PY
