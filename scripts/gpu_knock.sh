#!/bin/bash
# development aid: knock-out timings (DF_DIAG builds); usage: gpu_knock.sh <tag> <libdir> <flag values...>
tag=$1; export DFB200_LIB_DIR=$PWD/deep-fusion_b200/$2; shift; shift
out=gpurun_out/$tag.log
for v in "$@"; do
  echo "== DF_DEBUG_NO_MMA=$v" >> $out
  DF_DEBUG_NO_MMA=$v timeout 120 python scripts/sweep_batch.py cfg3 2>&1 | grep -E "N= +(1024) " >> $out
  DF_DEBUG_NO_MMA=$v timeout 120 python scripts/sweep_batch.py cfg1 2>&1 | grep -E "N= +(1024) " >> $out
done
cat $out
