#!/bin/bash
# development aid: parity of the BASELINE shapes + cfg3 / cfg4 timings with several library builds
# usage: gpu_pipe.sh <out-tag> <libdir> [<libdir> ...]
out=gpurun_out/$1; shift
for L in "$@"; do
  echo "== $L" >> ${out}_variants.log
  DFB200_LIB_DIR=$PWD/deep-fusion_b200/$L timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "baseline_configs_full_size or batch_properties or back_to_back" 2>&1 | tail -3 >> ${out}_variants.log
  for s in cfg3 cfg4; do
    SWEEP_NS=64,256,1024 DFB200_LIB_DIR=$PWD/deep-fusion_b200/$L timeout 200 python scripts/sweep_batch.py $s 2>&1 | grep -E "N= " >> ${out}_variants.log
  done
done
cat ${out}_variants.log
