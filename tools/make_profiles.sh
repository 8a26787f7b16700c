#!/bin/bash
# Run on the GPU box (under gpurun, ONE GPU): writes everything under gpurun_out/prof/ ; copy the summaries into profiles/.
#   gpurun --timeout 2400 -- 'bash tools/make_profiles.sh'
mkdir -p gpurun_out/prof
export BDM_BATCH=32
# 1. every launch of one eager PC^2 sampler iteration (B = 32) with its device time
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/prof/step_launches.csv \
    python tools/step_launches.py > /dev/null 2>&1
python tools/launch_share.py gpurun_out/prof/step_launches.csv > gpurun_out/prof/step_launch_share.md
# 2. a window of ~2 sampler iterations (graph replays) out of the middle of the bench command's second warm-up job
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 600000 -c 900 --csv \
    --log-file gpurun_out/prof/bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-breakdown > gpurun_out/prof/bench_under_ncu.log 2>&1
python tools/launch_share.py gpurun_out/prof/bench_launches.csv all > gpurun_out/prof/bench_launch_share.md 2>&1
# 3. one ncu --set full capture per top kernel
for spec in "attention:attention_tc05_kernel" "sparse_conv:sparse_conv3_gather" "groupnorm_cl:gn_cl_apply" "devox_cl:devox_cl" \
            "groupnorm_mid:gn_cluster_kernel" "voxelize:vox_fill"; do
  op=${spec%%:*}; kern=${spec##*:}
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:"$kern" -s 1 -c 1 -o /tmp/ncu_$op -f \
      python tools/run_op.py $op --reps 3 --batch 32 > /dev/null 2>&1
  cp /tmp/ncu_$op.ncu-rep gpurun_out/prof/ 2>/dev/null
done
python tools/ncu_summary.py /tmp/ncu_*.ncu-rep > gpurun_out/prof/ncu_kernels.md 2>&1
# 3b. the tcgen05 convolution, one capture per kernel instance of the step (B R C)
for c in "32 32 64" "32 16 128" "32 32 32"; do
  n=$(echo $c | tr " " "_")
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3_tc05_kernel -s 2 -c 1 -o gpurun_out/prof/conv3_$n -f \
      python tools/conv3_check.py --only $c > /dev/null 2>&1
done
python tools/ncu_summary.py gpurun_out/prof/conv3_*.ncu-rep > gpurun_out/prof/ncu_conv3_tc05.md 2>&1
timeout 300 python tools/conv3_check.py > gpurun_out/prof/conv3_check.log 2>&1
timeout 300 python tools/conv3_sparse_time.py > gpurun_out/prof/conv3_sparse_time.log 2>&1
# 4. op-level timings and the stand-alone probes
timeout 300 python tools/gn_bench.py > gpurun_out/prof/gn_bench.log 2>&1
timeout 300 python tools/attn_bench.py > gpurun_out/prof/attn_bench.log 2>&1
BDM_ATTENTION=mma timeout 300 python tools/attn_bench.py > gpurun_out/prof/attn_bench_mma.log 2>&1
for v in 0 2 4 6 8 10; do timeout 30 tools/probe/tc05_probe $v; done > gpurun_out/prof/tc05_probe.log 2>&1
# 5. the bench line itself
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/prof/bench_n1.json 2> gpurun_out/prof/bench_n1.err
ls -la gpurun_out/prof
