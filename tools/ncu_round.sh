#!/bin/bash
# ncu evidence for one round (run under gpurun): launch list of the bench command + full captures of the hot kernels.
# usage: tools/ncu_round.sh <tag>
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'persistent_greedy_kernel|conv_pool_kernel|conv1_kernel|fc_splitk' -s 15 -c 5 -o gpurun_out/prof_enc_greedy_$TAG -f $CMD > gpurun_out/ncu_full1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'persistent_beam_kernel' -s 1 -c 1 -o gpurun_out/prof_beam_$TAG -f $CMD > gpurun_out/ncu_full2_$TAG.log 2>&1
ls -la gpurun_out/ | tail -8
