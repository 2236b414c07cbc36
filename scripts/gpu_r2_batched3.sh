#!/bin/bash
# evidence for DESIGN 4b: timeline of the launch chain, kernel-boundary floor, graph vs host launches
mkdir -p gpurun_out
{ timeout 100 python scripts/chain_trace.py 16 2>&1 | sed -n 1,70p; timeout 100 python scripts/chain_trace.py 64 2>&1 | sed -n 1,70p; } > gpurun_out/r02_chain_timeline.txt
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/pdl_gap scripts/ubench/pdl_gap.cu && timeout 120 /tmp/pdl_gap > gpurun_out/r02_pdl_gap.txt 2>&1
timeout 120 python scripts/graph_probe.py 2>&1 | grep "us/step" | tee gpurun_out/r02_graph_probe.txt
