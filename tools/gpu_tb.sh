#!/bin/bash
# tests + bench (no CPU baseline).  usage: bash tools/gpu_tb.sh TAG
TAG=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; head -c 330 gpurun_out/${TAG}_bench.json; echo; tail -3 gpurun_out/${TAG}_bench.err
exit 0
