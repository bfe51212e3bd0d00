#!/bin/bash
# Round-end measurement on one B200 (run through gpurun): default bench line, the ncu launch list of the same command,
# one `ncu --set full` capture of a whole decode step and of the splitter-encode kernels. Outputs under gpurun_out/.
TAG=${1:-r1_final}
python bench.py > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
tail -c 600 gpurun_out/bench_${TAG}.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ --csv --log-file gpurun_out/launches_${TAG}.csv python bench.py > gpurun_out/ncu_launch_${TAG}.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:'k_plan|k_decode' -s 15 -c 5 -f -o gpurun_out/prof_step_${TAG} python tools/decode_bench.py 6000 6 > gpurun_out/ncu_step_${TAG}.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:k_svb -s 6 -c 2 -f -o gpurun_out/prof_enc_${TAG} python tools/encode_bench.py 6000 6 > gpurun_out/ncu_enc_${TAG}.log 2>&1
python tools/encode_bench.py 20000 6,0 2>&1 | grep GB
python tools/decode_bench.py 20000 6 --phases 2>&1 | tail -4
