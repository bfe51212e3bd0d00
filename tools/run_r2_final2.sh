#!/bin/bash
# Refresh of the round-2 bench lines after the e2e pipeline fixes (tests, bench, reference arm, launch list)
set -u
mkdir -p gpurun_out
TAG=${1:-r2_final2}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
SECONDS=0
python bench.py > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "bench rc=$? ${SECONDS}s"
SECONDS=0
python bench.py --impl reference > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$? ${SECONDS}s"
SECONDS=0
timeout 900 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launch list rc=$? ${SECONDS}s"
python - <<PY
import json
d=json.load(open('gpurun_out/${TAG}_n1.json')); r=json.load(open('gpurun_out/${TAG}_ref.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['pcie_link']['rank0'],'enc e2e',d['e2e'].get('encode_gpu_entropy'),'ref',r['value'])
PY
