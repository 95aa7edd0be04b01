#!/bin/bash
# Round-2 GPU-box visit: parity tests, bench + reference arm, ncu launch lists (eval step, training step) and
# --set full captures of the dominant kernels.  Usage: gpurun --timeout 2400 -- 'bash tools/gpu_round2.sh [tag] [stages]'
# stages: t tests, b bench, l launch lists, f ncu --set full.  Only CSV / JSON / logs come back (the .ncu-rep stay in /tmp).
TAG=${1:-r02}
STAGES=${2:-tblf}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [[ $STAGES == *t* ]]; then
  timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -3 gpurun_out/${TAG}_pytest.log
fi
if [[ $STAGES == *b* ]]; then
  timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench exit $?"; head -c 300 gpurun_out/${TAG}_bench.json; echo
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
  echo "reference arm exit $?"
fi
if [[ $STAGES == *l* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv \
    --log-file gpurun_out/${TAG}_launches_eval.csv python tools/ncu_target.py --reps 2 > gpurun_out/${TAG}_launches_eval.log 2>&1
  echo "launch list (eval) exit $?"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/${TAG}_launches_train.csv python tools/ncu_target.py --train --reps 1 > gpurun_out/${TAG}_launches_train.log 2>&1
  echo "launch list (train) exit $?"
fi
if [[ $STAGES == *f* ]]; then
  # second eager forward: steady state (weights packed); 20 conv launches of one decode step
  timeout 900 ncu --set full --clock-control none -k regex:conv3x3_umma -s 420 -c 20 -o /tmp/${TAG}_conv -f \
    python tools/ncu_target.py --reps 2 > gpurun_out/${TAG}_ncu_conv.log 2>&1
  echo "ncu conv exit $?"
  ncu -i /tmp/${TAG}_conv.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_umma_full_raw.csv 2>&1
  K1='regex:canvas_conv|paste_back_kernel|extract_rows|extract_cols|build_filters|controller_cluster|score_kernel|iou_umma_kernel|iou_umma_finalize|pairwise_iou_kernel|gt_box_kernel|hungarian_kernel|pp_label'
  timeout 900 ncu --set full --clock-control none -k "$K1" -s 130 -c 14 -o /tmp/${TAG}_step -f \
    python tools/ncu_target.py --reps 2 > gpurun_out/${TAG}_ncu_step.log 2>&1
  echo "ncu step exit $?"
  ncu -i /tmp/${TAG}_step.ncu-rep --page raw --csv > gpurun_out/${TAG}_step_full_raw.csv 2>&1
  K2='regex:conv_bwd_weight_kernel|bn_bwd_apply|bn_bwd_reduce|ex_T_kernel|ex_dfy_kernel|ex_dfx_kernel|pb_row_kernel|pb_col_kernel|controller_bwd_kernel|outer_sum_kernel'
  timeout 900 ncu --set full --clock-control none -k "$K2" -s 0 -c 70 -o /tmp/${TAG}_train -f \
    python tools/ncu_target.py --train --reps 1 --batch 8 > gpurun_out/${TAG}_ncu_train.log 2>&1
  echo "ncu train exit $?"
  ncu -i /tmp/${TAG}_train.ncu-rep --page raw --csv > gpurun_out/${TAG}_train_full_raw.csv 2>&1
fi
ls -la gpurun_out | head -40
exit 0
