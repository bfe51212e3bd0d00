#!/bin/bash
# e2e (rnsc_decompress_batch, pinned host buffers) and the box's PCIe link beside it
python tools/pcie_probe.py 2>&1 | tail -3 | head -1
python tools/e2e_bench.py 40000 2>&1 | tail -1
RNSC_SLICE_SAMPLES=16777216 python tools/e2e_bench.py 40000 2>&1 | tail -1
RNSC_SLICE_SAMPLES=33554432 python tools/e2e_bench.py 40000 2>&1 | tail -1
