#!/bin/bash
# round-end evidence run (one gpurun call): tests, both bench arms, ncu launch list, ncu --set full captures, sweeps
tag=${1:-v3}; o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > $o/r02_pytest_gpu_$tag.log
python bench.py --steps 100 --warmup 10 > $o/r02_bench_$tag.json 2> $o/r02_bench_$tag.err
python bench.py --impl reference --steps 20 --warmup 3 > $o/r02_bench_${tag}_reference.json 2>> $o/r02_bench_$tag.err
B="python bench.py --steps 20 --warmup 5 --no-cpu --no-graph --quick"
$B > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r02_launches_$tag.csv $B > /dev/null 2>&1
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 10 -c 3 -f -o $o/r02_conv_cfg3_$tag $B > /dev/null 2>&1
$B > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:concat -s 5 -c 3 -f -o $o/r02_concat_$tag $B > /dev/null 2>&1
C="python scripts/bench_one.py cfg4 256 20"
$C > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_pair -s 5 -c 1 -f -o $o/r02_conv_cfg4_pair_$tag $C > /dev/null 2>&1
for s in cfg3 cfg1 cfg4; do timeout 200 python scripts/sweep_batch.py $s > $o/r02_sweep_${s}_$tag.log 2>&1; done
cat $o/r02_pytest_gpu_$tag.log; tail -3 $o/r02_sweep_cfg3_$tag.log
