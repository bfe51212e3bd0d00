#!/bin/bash
# e2e (rnsc_decompress_batch, pinned host buffers) against the slice size, and the box's PCIe link beside it
python tools/pcie_probe.py 2>&1 | tail -4
for s in 8388608 16777216 33554432 67108864 134217728 268435456; do RNSC_SLICE_SAMPLES=$s python tools/e2e_bench.py 40000 2>&1 | tail -1; done
