#!/bin/bash
# Round-2 evidence pass on one GPU box (under gpurun): instruction / DRAM counters of every kernel of the step, the serialised
# launch list, and one `ncu --set full` capture of the kernels named.  Everything lands in gpurun_out/<tag>_*.
# Usage:  bash profiles/tools/gpu_round2.sh <tag> "<kernels for --set full>"
tag=${1:-r02z}
kern=${2:-"k_observe k_rs_check k_advance k_rs_enumerate k_rs_walk k_pack_lidar"}
out=gpurun_out
mkdir -p $out
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum
ncu --metrics $M --clock-control none -k regex:"k_advance|k_observe|k_rs_|k_pack" -c 60 --csv --log-file $out/${tag}_counts.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-only > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-only > /dev/null 2>&1
python profiles/tools/launch_shares.py $out/${tag}_launches.csv > $out/${tag}_launch_shares.txt 2>&1
cat $out/${tag}_launch_shares.txt | head -12
for k in $kern; do
  if [ "$k" = "k_policy_forward" ] || [ "$k" = "k_norm_partial" ]; then
    ncu --set full --clock-control none --import-source on -k regex:"$k" -s 5 -c 1 -f -o $out/${tag}_full_$k \
        python bench.py --config rollout --steps 4 --warmup 4 > /dev/null 2>&1
  elif [ "$k" = "k_pack_lidar" ]; then
    HOPE_B200_HOST_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:"$k" -s 8 -c 1 -f -o $out/${tag}_full_$k \
        python profiles/tools/e2e_sweep.py --child --steps 6 --warmup 6 > /dev/null 2>&1
  else
    ncu --set full --clock-control none --import-source on -k regex:"^(void )?(hope::)?$k" -s 5 -c 1 -f -o $out/${tag}_full_$k \
        python bench.py --steps 3 --warmup 3 --no-cpu-baseline --device-only > /dev/null 2>&1
  fi
done
ls -la $out/${tag}_* | tail -20
