#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2j}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log; tail -5 gpurun_out/${TAG}_tests.log
df -h /tmp | tail -1; free -g | head -2
python tools/stream_repack_bench.py ${2:-20000} /tmp/rnsc_stream > gpurun_out/${TAG}_stream.json 2> gpurun_out/${TAG}_stream.err; echo "stream rc=$?"; tail -3 gpurun_out/${TAG}_stream.err; cat gpurun_out/${TAG}_stream.json
