#!/bin/bash
# e2e decode (rnsc_decompress_batch) and encode (rnsc_compress_batch, GPU entropy stage) with pinned host buffers, the box PCIe link beside them, and the slice size
python tools/pcie_probe.py 2>&1 | tail -3 | head -1
python tools/e2e_bench.py 40000 2>&1 | tail -1
python tools/e2e_encode_bench.py 40000 6 2>&1 | tail -1
RNSC_SLICE_SAMPLES=16777216 python tools/e2e_bench.py 40000 2>&1 | tail -1
RNSC_SLICE_SAMPLES=33554432 python tools/e2e_bench.py 40000 2>&1 | tail -1
