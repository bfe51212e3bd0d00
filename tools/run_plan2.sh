#!/bin/bash
# planner variants at the bench size (100k reads): bench.py decode leg only, default library vs a variant copied over it (scratch copy of the repo on the box)
set -u
mkdir -p gpurun_out
python bench.py --no-extra --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('default', d['value'], d['ms_per_step'])"
for v in "$@"; do
  cp tools/variants/$v rawnanoporesignalcompression_b200/librnsc.so
  python bench.py --no-extra --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['ms_per_step'])"
done
