#!/bin/bash
# Run on the GPU box (under gpurun): writes everything under gpurun_out/prof/
mkdir -p gpurun_out/prof
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prof/step_launches.csv \
    python tools/step_launches.py > /dev/null 2>&1
python tools/op_bench.py --iters 10 > gpurun_out/prof/op_bench.log 2>&1; cp gpurun_out/op_bench.json gpurun_out/prof/op_bench.json
python tools/sparse_conv_bench.py > gpurun_out/prof/sparse_conv_bench.log 2>&1; cp gpurun_out/sparse_conv_bench.json gpurun_out/prof/sparse_conv_bench.json
for op in voxelize devoxelize fps ball_query three_nn grouping sparse_conv attention groupnorm_cl groupnorm_small devox_cl; do
  ncu --set full --clock-control none --import-source on \
      -k regex:"vox_fill|vox_sort|devox_|fps_register|ball_query_kernel|three_nn_kernel|three_interp|grouping_|gn_|sparse_conv3|attention_hd64" \
      -c 6 -o /tmp/ncu_$op -f python tools/run_op.py $op --reps 2 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/ncu_$op.ncu-rep > gpurun_out/prof/ncu_$op.md 2>&1   # the reports themselves are too big to bring back
done
ncu --set full --clock-control none --import-source on -k regex:"gn_stats|gn_apply" -c 2 -o /tmp/ncu_groupnorm -f python -c "
import torch, sys
sys.path.insert(0, '.')
from bdm_b200 import backend as B
x = torch.randn(16, 64, 32768, device='cuda'); w = torch.randn(64, device='cuda'); b = torch.randn(64, device='cuda')
for _ in range(2): B.groupnorm_act(x, 8, w, b, 1e-5, True, conv_bias=b)
torch.cuda.synchronize()" > /dev/null 2>&1
python tools/ncu_summary.py /tmp/ncu_groupnorm.ncu-rep > gpurun_out/prof/ncu_groupnorm.md 2>&1
python tools/launch_share.py gpurun_out/prof/step_launches.csv > gpurun_out/prof/step_launch_share.md
python bench.py --steps 20 --warmup 5 > gpurun_out/prof/bench_n1.json 2> gpurun_out/prof/bench_n1.err
ls -la gpurun_out/prof
