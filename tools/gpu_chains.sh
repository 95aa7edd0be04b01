#!/bin/bash
# tests + bench at 1 / 2 / 4 sub-batch chains
TAG=${1:-r01d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
for n in 1 2 4; do
  RA_CHAINS=$n timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_chains$n.json 2> gpurun_out/${TAG}_bench_chains$n.err
  echo "chains $n exit $?"; head -c 330 gpurun_out/${TAG}_bench_chains$n.json; echo
done
exit 0
