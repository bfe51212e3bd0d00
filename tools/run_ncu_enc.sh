#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-ncuenc}
ncu --set full --clock-control none --import-source on -k regex:"k_zenc_emit|k_zenc_build|k_zenc_histogram|k_svb_pack_slab" -s 8 -c 4 -o gpurun_out/${TAG}_enc python tools/dev_encode_bench.py 20000 6 --steps 2 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
