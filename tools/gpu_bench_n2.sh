#!/bin/bash
# N-rank validation of bench.py (reduced reads): NP=4 gpurun --gpus 4 -- bash tools/gpu_bench_n2.sh (default 2 ranks)
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NP:-2} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${NP:-2} --steps 3 --warmup 3 --reads 8000 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"
tail -5 gpurun_out/bench_n2.err | cut -c1-600
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
print(d['value'], d['n_gpus'], d['e2e']['value'], d['e2e']['pcie_link'].get('limiter'), d['e2e']['pcie_link'].get('box_aggregate'))
c5=d['other_legs']['config5_strong_scaling']; print(c5['decode'], c5['bit_exact'], c5['table_checksum_decoded'], c5['table_checksum_source'], c5['ranks'])
print(d['other_legs']['config3_encode_gpu_entropy'])
PY
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 --cpu-reads 200 --cpu-seconds 3 > gpurun_out/bench_ref.json 2>/dev/null; head -c 600 gpurun_out/bench_ref.json
