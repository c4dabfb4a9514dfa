#!/bin/bash
# Run ON THE GPU BOX (gpurun): ncu launch list + --set full captures of the hot kernels of one bench step
# (tools/ncu_step.py: two steps of the cfg2 workload).  Reports land in gpurun_out/<tag>_*.ncu-rep / .csv; summarise them
# here with tools/ncu_summary.py / tools/launch_summary.py and commit the summaries under profiles/.
#   bash tools/capture_profiles.sh <tag>
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
T="timeout 240"
$T $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $OUT/${TAG}_step_launches.csv python tools/ncu_step.py > /dev/null 2>&1
FULL="$NCU --set full --import-source on"
$T $FULL -k regex:attn_pv_kernel -s 12 -c 1 -f -o $OUT/${TAG}_attnpv python tools/ncu_step.py > /dev/null 2>&1
$T $FULL -k regex:attn_tc_kernel -s 17 -c 3 -f -o $OUT/${TAG}_attntc python tools/ncu_step.py > /dev/null 2>&1
$T $FULL -k regex:gemm_tc_kernel -s 66 -c 5 -f -o $OUT/${TAG}_gemm python tools/ncu_step.py > /dev/null 2>&1
$T $FULL -k regex:par_affinity -s 3 -c 1 -f -o $OUT/${TAG}_paraff python tools/ncu_step.py > /dev/null 2>&1
for s in 60 80 100; do
  $T $FULL -k regex:par_iterate -s $s -c 1 -f -o $OUT/${TAG}_par_s$s python tools/ncu_step.py > /dev/null 2>&1
done
$T $FULL -k 'regex:cam_|amax_kernel|split_scaled' -s 6 -c 6 -f -o $OUT/${TAG}_cam python tools/ncu_step.py > /dev/null 2>&1
ls -la $OUT | grep ${TAG}_
