#!/bin/bash
# refill-threshold sweep of the traversal kernels on the C2 scene (under gpurun)
for v in 4 8 12 16 20 24 28; do
  echo "== HC_TRACE_REFILL=$v"
  HC_TRACE_REFILL=$v python scripts/gpu_trace_perf.py refill 2>&1 | tail -1
done
