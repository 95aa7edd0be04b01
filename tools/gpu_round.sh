#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one --set full capture of the
# dominant kernel.  Usage (from the repo root):  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag] [stages]'
# stages: any of t (tests) b (bench) l (launch list) f (ncu --set full).  Default: tblf
TAG=${1:-r01}
STAGES=${2:-tblf}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [[ $STAGES == *t* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
if [[ $STAGES == *b* ]]; then
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench exit $?"
  head -c 1500 gpurun_out/${TAG}_bench.json
fi
if [[ $STAGES == *l* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_launches_bench.log 2>&1
  echo "launch list exit $?"
fi
if [[ $STAGES == *f* ]]; then
  # skip the first forward's lazy packing launches: capture 24 conv launches of a steady-state decode step
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 60 -c 24 \
    -o gpurun_out/${TAG}_conv_umma_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_full_bench.log 2>&1
  echo "ncu full exit $?"
  ncu -i gpurun_out/${TAG}_conv_umma_full.ncu-rep --page raw --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,sm__issue_active.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
    > gpurun_out/${TAG}_conv_umma_full_summary.csv 2>&1
fi
exit 0
