#!/bin/bash
# Evidence pass for the end of round 2 (run under gpurun, one GPU): ncu launch lists of one UNet step and one MoVQ decode
# (time + DRAM bytes per launch) and --set full captures of the kernels written last (attention with P in tensor memory,
# the fused MoVQ attention, SpatialNorm apply).  Outputs under gpurun_out/; summarised into profiles/ by ncu_summary.py /
# traffic_summary.py on the build container.
set -x
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r2b.csv \
    python profiles/ncu_step.py unet > gpurun_out/ncu_step_unet.txt 2>&1
ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_movq_r2b.csv \
    python profiles/ncu_step.py movq > gpurun_out/ncu_step_movq.txt 2>&1
FULL="--set full --clock-control none --import-source on --profile-from-start off"
ncu $FULL -k regex:attention_d64 --launch-count 1 -o gpurun_out/ncu_r2b_attention_l1 -f python profiles/ncu_step.py unet > /dev/null 2>&1
ncu $FULL -k regex:attention_d512 --launch-skip 3 --launch-count 1 -o gpurun_out/ncu_r2b_attention_d512 -f python profiles/ncu_step.py movq > /dev/null 2>&1
ncu $FULL -k regex:sn_apply --launch-skip 28 --launch-count 1 -o gpurun_out/ncu_r2b_sn_apply -f python profiles/ncu_step.py movq > /dev/null 2>&1
ls -la gpurun_out
