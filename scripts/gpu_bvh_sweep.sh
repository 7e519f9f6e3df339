#!/bin/bash
# builder-parameter sweep on the C2 scene: traversal throughput per variant (under gpurun)
mkdir -p gpurun_out
for v in "" "HC_BVH_BINS=32" "HC_BVH_PICK=1" "HC_BVH_MAXLEAF=2 HC_BVH_BLOCK=2" "HC_BVH_BLOCK=2" "HC_BVH_SWEEP=64" "HC_BVH_SWEEP=64 HC_BVH_BINS=32 HC_BVH_PICK=1" "HC_BVH_MAXLEAF=6 HC_BVH_BLOCK=2" "HC_BVH_MAXLEAF=8"; do
  echo "== $v"
  env $v python scripts/gpu_trace_perf.py sweep 2>&1 | tail -1
done
