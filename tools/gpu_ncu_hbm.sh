#!/bin/bash
# ncu --set full of the non-conv kernels of one decode step and of the loss block; only the CSV pages come back
TAG=${1:-r01j}
mkdir -p gpurun_out
K1='regex:canvas_conv|paste_back|extract_rows|extract_cols|build_filters|controller_cluster|score_kernel'
timeout 600 ncu --set full --clock-control none -k "$K1" -s 70 -c 7 -o /tmp/${TAG}_step -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench1.log 2>&1
echo "ncu step exit $?"
ncu -i /tmp/${TAG}_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_full_raw.csv 2>&1
K2='regex:pairwise_iou_kernel|gt_box_kernel|hungarian_kernel'
timeout 600 ncu --set full --clock-control none -k "$K2" -c 5 -o /tmp/${TAG}_loss -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench2.log 2>&1
echo "ncu loss exit $?"
ncu -i /tmp/${TAG}_loss.ncu-rep --page raw --csv > gpurun_out/${TAG}_loss_full_raw.csv 2>&1
ls -la gpurun_out/ /tmp/${TAG}_*.ncu-rep
exit 0
