#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2k}
python tools/stream_repack_bench.py ${2:-20000} /tmp/rnsc_stream > gpurun_out/${TAG}_stream.json 2> gpurun_out/${TAG}_stream.err; echo "stream rc=$?"; tail -3 gpurun_out/${TAG}_stream.err; cat gpurun_out/${TAG}_stream.json
