#!/bin/bash
# decode variants (libraries under tools/variants) against the default library on the 20k-read C5 set, with phases
python tools/decode_bench.py 20000 6 --model R10.4.1 --cache /tmp/c5.npz --phases 2>&1 | grep -E "S_256|A_512|GB/s"
for v in tools/variants/*.so; do echo "$v: "; python tools/decode_bench.py 20000 6 --model R10.4.1 --cache /tmp/c5.npz --lib $v --phases 2>&1 | grep -E "S_256|A_512|GB/s"; done
