#!/bin/bash
set -u
mkdir -p gpurun_out
python tools/dev_encode_bench.py 6000 6 > gpurun_out/r2h_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_zenc_build" -s 2 -c 1 -o gpurun_out/r2h_build python tools/dev_encode_bench.py 6000 6 --steps 2 > gpurun_out/r2h_ncu.log 2>&1
tail -3 gpurun_out/r2h_plain.log; tail -3 gpurun_out/r2h_ncu.log; ls -la gpurun_out/r2h_build.ncu-rep
