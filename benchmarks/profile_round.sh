#!/bin/bash
# One gpurun call's worth of ncu evidence (B200_PROFILING.md recipe): every program first runs WITHOUT ncu and must exit 0.
#   gpurun --timeout 1500 -- 'bash benchmarks/profile_round.sh r02'
# Outputs (gpurun_out/, summarised into profiles/ by benchmarks/ncu_summary.py on the CPU box):
#   <tag>_launches_eager_steps3.csv     every launch of 3 eager steps of the bench (device time per launch)
#   <tag>_prof_step.ncu-rep             --set full of the step's heaviest kernels
#   <tag>_prof_{sim_ts,seg_strip,nce_bwd,adapter}.ncu-rep   --set full of the kernels the 16-query step does not exercise
TAG=${1:-r02}
OUT=gpurun_out
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-graph --no-secondary --no-u8-variant"
$B > $OUT/${TAG}_plain_step.log 2>&1 || { echo "plain bench failed"; exit 1; }
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches_eager_steps3.csv $B > $OUT/${TAG}_ncu_launches.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"mask_prep_staged|pool_umma_kernel|gemm_umma_kernel|seg_loss_tile|infonce_bwd_kernel|infonce_tail|fgbg_reduce" -c 7 \
    -o $OUT/${TAG}_prof_step -f $B > $OUT/${TAG}_ncu_step.log 2>&1
[ "$2" = "step-only" ] && exit 0
python benchmarks/one_sim.py 1024 102400 lse > $OUT/${TAG}_plain_sim.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:sim_umma_ts -s 2 -c 1 -o $OUT/${TAG}_prof_sim_ts -f python benchmarks/one_sim.py 1024 102400 lse > $OUT/${TAG}_ncu_sim.log 2>&1
python benchmarks/one_seg.py 128 > $OUT/${TAG}_plain_seg.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:seg_loss_strip -s 2 -c 1 -o $OUT/${TAG}_prof_seg_strip -f python benchmarks/one_seg.py 128 > $OUT/${TAG}_ncu_seg.log 2>&1
python benchmarks/nce_bench.py --iters 1 --shapes 1024x102400 > $OUT/${TAG}_plain_nce.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:nce_bwd_umma -s 2 -c 2 -o $OUT/${TAG}_prof_nce_bwd -f python benchmarks/nce_bench.py --iters 1 --shapes 1024x102400 > $OUT/${TAG}_ncu_nce.log 2>&1
python benchmarks/one_adapter.py 16 16 > $OUT/${TAG}_plain_adapter.log 2>&1 && \
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"gemm_umma_kernel|dwconv_cl|ln_rows" -c 8 -o $OUT/${TAG}_prof_adapter -f python benchmarks/one_adapter.py 16 16 > $OUT/${TAG}_ncu_adapter.log 2>&1
ls -la $OUT/${TAG}_*
