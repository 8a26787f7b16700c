#!/bin/bash
# Run on the GPU box (under gpurun): writes everything under gpurun_out/prof/
mkdir -p gpurun_out/prof
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prof/step_launches.csv \
    python tools/step_launches.py > /dev/null 2>&1
for op in voxelize devoxelize fps ball_query three_nn grouping; do
  ncu --set full --clock-control none --import-source on \
      -k regex:"vox_fill|vox_sort|devox_|fps_register|ball_query_kernel|three_nn_kernel|three_interp|grouping_" \
      -c 6 -o gpurun_out/prof/ncu_$op -f python tools/run_op.py $op --reps 2 > /dev/null 2>&1
done
ls -la gpurun_out/prof
