#!/bin/bash
# Round-2 ncu evidence (run under gpurun): launch list of the bench command + full captures of the kernels that changed
# in round 2, summarised to CSV ON THE BOX (gpurun copies back at most 64 MiB: the large reports are deleted after the
# extraction; the kernels round 2 did not touch keep their r1s4 captures).
TAG=${1:-r2}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'persistent_greedy_kernel|conv_pool_kernel|conv1_u8_kernel|fc_splitk' -s 12 -c 5 -o gpurun_out/prof_enc_greedy_$TAG -f $CMD > gpurun_out/ncu_full1_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'persistent_beam_kernel' -s 1 -c 1 -o gpurun_out/prof_beam_$TAG -f $CMD > gpurun_out/ncu_full2_$TAG.log 2>&1
python tools/time_wide.py 1024 512 512 2 150 > gpurun_out/time_wide_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'wide_loop_kernel' -s 2 -c 1 -o gpurun_out/prof_wide_$TAG -f python tools/time_wide.py 1024 512 512 2 150 nosample > gpurun_out/ncu_full7_$TAG.log 2>&1
python tools/ncu_extract.py gpurun_out/${TAG}_ncu_full_selected_metrics.csv gpurun_out/prof_enc_greedy_$TAG.ncu-rep gpurun_out/prof_beam_$TAG.ncu-rep gpurun_out/prof_wide_$TAG.ncu-rep > gpurun_out/extract_$TAG.log 2>&1
rm -f gpurun_out/prof_beam_$TAG.ncu-rep
du -sh gpurun_out
