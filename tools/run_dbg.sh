#!/bin/bash
python tools/decode_bench.py 20000 6 --model R10.4.1 --phases --lib tools/variants/librnsc_dbg.so 2>&1 | tail -4
