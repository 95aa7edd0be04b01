#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one --set full capture of the
# dominant kernel.  Usage (from the repo root):  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag] [stages]'
# stages: any of t (tests) b (bench) l (launch list) f (ncu --set full).  Default: tblf
# Only CSV / JSON / logs are left under gpurun_out/ (the .ncu-rep files stay in /tmp: gpurun returns at most 64 MiB).
TAG=${1:-r01}
STAGES=${2:-tblf}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
if [[ $STAGES == *t* ]]; then
  timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -5 gpurun_out/${TAG}_pytest.log
fi
if [[ $STAGES == *b* ]]; then
  timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench exit $?"
  head -c 400 gpurun_out/${TAG}_bench.json; echo
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
  echo "reference arm exit $?"; head -c 300 gpurun_out/${TAG}_bench_reference.json; echo
fi
if [[ $STAGES == *l* ]]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_launches_bench.log 2>&1
  echo "launch list exit $?"
fi
if [[ $STAGES == *f* ]]; then
  # skip the first forward's lazy packing launches: capture the 20 conv launches of one steady-state decode step
  timeout 900 ncu --set full --clock-control none -k regex:conv3x3_umma -s 60 -c 20 \
    -o /tmp/${TAG}_conv_umma_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_full_bench.log 2>&1
  echo "ncu full exit $?"
  ncu -i /tmp/${TAG}_conv_umma_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_conv_umma_full_raw.csv 2>&1
fi
ls -la gpurun_out | head -30
exit 0
