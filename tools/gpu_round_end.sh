#!/bin/bash
# Round-2 final measurement on one B200: GPU tests, smoke, default bench line + reference arm, ncu launch list of the bench command
# (first 400 launches of the bench process itself), ncu --set full of one decode launch group on the bench model.
set -u
mkdir -p gpurun_out
TAG=${1:-r2_final}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python tools/foreign_routing.py > gpurun_out/${TAG}_foreign_routing.txt 2>&1; echo "foreign routing rc=$?"
SECONDS=0
python bench.py > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "bench rc=$? ${SECONDS}s"
head -c 400 gpurun_out/${TAG}_n1.json; echo
SECONDS=0
python bench.py --impl reference > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$? ${SECONDS}s"
head -c 300 gpurun_out/${TAG}_ref.json; echo
SECONDS=0
timeout 900 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu launch list rc=$? ${SECONDS}s"
ncu --set full --clock-control none --import-source on -k regex:"k_decode_fast|k_plan" -s 12 -c 4 -o gpurun_out/${TAG}_dec python tools/decode_bench.py 20000 6 --model R10.4.1 > gpurun_out/${TAG}_ncu.log 2>&1; tail -1 gpurun_out/${TAG}_ncu.log
ncu --set full --clock-control none --import-source on -k regex:"k_zenc|k_svb_pack_slab" -s 8 -c 6 -o gpurun_out/${TAG}_enc python tools/dev_encode_bench.py 20000 6 --steps 2 > gpurun_out/${TAG}_ncu_enc.log 2>&1; tail -1 gpurun_out/${TAG}_ncu_enc.log
