#!/bin/bash
TAG=${1:-r01e}
mkdir -p gpurun_out
timeout 120 tools/launch_cost > gpurun_out/${TAG}_launch_cost.txt 2>&1
cat gpurun_out/${TAG}_launch_cost.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench exit $?"; head -c 330 gpurun_out/${TAG}_bench.json; echo
RA_CONV_CARVEOUT=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_carveout.json 2> gpurun_out/${TAG}_bench_carveout.err
echo "bench carveout exit $?"; head -c 330 gpurun_out/${TAG}_bench_carveout.json; echo
exit 0
