#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2c}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_tests.log
tail -15 gpurun_out/${TAG}_tests.log
python - <<'PY' > gpurun_out/${TAG}_seam.json 2> gpurun_out/${TAG}_seam.err
import sys, os, subprocess, json, tempfile
sys.path.insert(0, '.')
from rawnanoporesignalcompression_b200 import synth
ls = synth.read_lengths(200, 20261018)
sig = synth.generate_reads(ls, 20261018, synth.R10_4_1).numpy()
so, sn, _ = synth.chunk_table(ls)
with tempfile.TemporaryDirectory() as td:
    p = os.path.join(td, 'c.bin'); synth.write_chunk_file(p, sig, so, sn)
    r = subprocess.run(['integration/_build/seam_bench', p, '4'], capture_output=True, text=True)
print(r.stdout); sys.stderr.write(r.stderr)
PY
cat gpurun_out/${TAG}_seam.json; tail -5 gpurun_out/${TAG}_seam.err
