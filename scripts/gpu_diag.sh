#!/bin/bash
# development aid: epilogue timelines from the lib_trace build (DF_EPI_TRACE=1, DF_DIAG=1) at several knock-out levels
export DFB200_LIB_DIR=$PWD/deep-fusion_b200/lib_trace
out=gpurun_out/${1:-diag}; shift
for v in "$@"; do
  DF_DEBUG_NO_MMA=$v timeout 120 python scripts/trace_conv.py cfg3 1024 400 > ${out}_trace_k$v.log 2>&1
done
