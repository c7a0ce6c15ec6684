#!/bin/bash
# development aid: time the cfg3 / cfg1 sweeps with several library builds (lib, lib_<tag> ...)
# usage: gpu_variants.sh <out-tag> <libdir> [<libdir> ...]
out=gpurun_out/$1; shift
for L in "$@"; do
  echo "== $L" >> ${out}_variants.log
  DFB200_LIB_DIR=$PWD/deep-fusion_b200/$L timeout 200 python scripts/sweep_batch.py cfg3 2>&1 | grep -E "N= +(64|256|1024) " >> ${out}_variants.log
  DFB200_LIB_DIR=$PWD/deep-fusion_b200/$L timeout 200 python scripts/sweep_batch.py cfg1 2>&1 | grep -E "N= +(64|1024) " >> ${out}_variants.log
done
cat ${out}_variants.log
