#!/bin/bash
# ncu evidence for one round (run under gpurun): launch list of the bench command + full captures of the hot kernels.
# usage: tools/ncu_round.sh <tag>
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
# headline kernels: skip the warm-up launches, take one of each
ncu --set full --clock-control none --import-source on -k regex:'persistent_greedy_kernel|conv_pool_kernel|conv1_u8_kernel|conv1_kernel|fc_splitk' -s 15 -c 5 -o gpurun_out/prof_enc_greedy_$TAG -f $CMD > gpurun_out/ncu_full1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'persistent_beam_kernel' -s 1 -c 1 -o gpurun_out/prof_beam_$TAG -f $CMD > gpurun_out/ncu_full2_$TAG.log 2>&1
# ResNet trunk: the generic conv kernel of one resnet18 forward (stem + the 3x3 / 1x1 convs)
ncu --set full --clock-control none --import-source on -k regex:'conv_igemm_kernel' -s 63 -c 21 -o gpurun_out/prof_resnet_$TAG -f python tools/bench_resnet.py resnet18 1024 320 > gpurun_out/ncu_full3_$TAG.log 2>&1
ls -la gpurun_out/ | tail -8
# sampling mode of the persistent kernel (decode_persistent.cu, MODE = 1)
ncu --set full --clock-control none --import-source on -k regex:'persistent_greedy_kernel' -s 2 -c 1 -o gpurun_out/prof_sample_$TAG -f python tools/time_sample.py 256 30 > gpurun_out/ncu_full4_$TAG.log 2>&1
# image preparation and evaluation metrics
ncu --set full --clock-control none --import-source on -k regex:'resize_rows_kernel|resize_cols_kernel' -s 6 -c 2 -o gpurun_out/prof_pre_$TAG -f python tools/time_preprocess.py > gpurun_out/ncu_full5_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'sequence_metrics_kernel' -s 3 -c 1 -o gpurun_out/prof_metrics_$TAG -f python tools/time_metrics.py > gpurun_out/ncu_full6_$TAG.log 2>&1
ls -la gpurun_out/ | tail -8
# round 2: the persistent whole-GPU loop of the reference's shipped decoder (decode_wide.cu), greedy and sampling
ncu --set full --clock-control none --import-source on -k regex:'wide_loop_kernel' -s 2 -c 1 -o gpurun_out/prof_wide_$TAG -f python tools/time_wide.py 1024 512 512 2 150 nosample > gpurun_out/ncu_full7_$TAG.log 2>&1
ls -la gpurun_out/ | tail -4
