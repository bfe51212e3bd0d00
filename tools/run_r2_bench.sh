#!/bin/bash
# Round-2 measurement on one B200: default bench line + reference arm, ncu launch list of the same command (first 600 launches
# of the bench process itself).
set -u
mkdir -p gpurun_out
TAG=${1:-r2_bench}
SECONDS=0
python bench.py > gpurun_out/${TAG}_n1.json 2> gpurun_out/${TAG}_n1.err; echo "bench rc=$? ${SECONDS}s"
head -c 1200 gpurun_out/${TAG}_n1.json; echo
SECONDS=0
python bench.py --impl reference > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$? ${SECONDS}s"
head -c 600 gpurun_out/${TAG}_ref.json; echo
if [ "${2:-}" = "ncu" ]; then
SECONDS=0
timeout 600 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --no-cpu-baseline > gpurun_out/${TAG}_ncu_launch.log 2>&1; echo "ncu rc=$? ${SECONDS}s"
fi
