#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2d}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -12 gpurun_out/${TAG}_tests.log
python bench.py --reads 20000 --steps 3 --cpu-reads 300 --cpu-seconds 4 > gpurun_out/${TAG}_bench20k.json 2> gpurun_out/${TAG}_bench20k.err; echo "bench rc=$?"
grep -E "config 3|per-chunk|config 5|config 4" gpurun_out/${TAG}_bench20k.err | cut -c1-1500
